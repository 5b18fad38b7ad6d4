#!/usr/bin/env python
"""Golden fixtures for softmasking (the reference's default mode on real genomes), from the UNMODIFIED reference.

BASELINE.json configs[2] runs  augustus --species=fly  with its defaults on chr2L: UTR states on, softmasking on (every lower-case
run of the input becomes a nonexonpart hint, extrinsicinfo.cc:1696-1724).  Writes
  fly_softmask_utr.params.xz   parameter blob of  --species=fly --UTR=on --softmasking=1
  fly_softmask_window.fa       40 kb of examples/chr2L (32 % lower case), case preserved
  ref_paths_softmask.json      Viterbi path + score + GC stairs of the reference for that window, and for the same window
                               upper-cased (= softmasking without any masked base)
  ref_samples_softmask.json.gz the 99 sampled paths of --sample=100 --alternatives-from-sampling=true
"""
import gzip
import json
import lzma
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from augustus_b200 import synth  # noqa: E402
from make_golden import AUGDUMP, ENV, REF, condense  # noqa: E402

ARGS = ["--species=fly", "--UTR=on", "--softmasking=1"]
START, LEN = 22150000, 40000


def run_one(name, dna, extra=()):
    with tempfile.TemporaryDirectory() as td:
        fa, pf = os.path.join(td, "w.fa"), os.path.join(td, "p")
        synth.write_fasta(fa, [dna], [name])
        subprocess.run([AUGDUMP] + ARGS + list(extra) + [fa], env=dict(ENV, AUGDUMP_PATH=pf), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        cur, samples = None, []
        for line in open(pf):
            t = line.split()
            if t[0] == "seq":
                cur = {"name": t[1], "length": int(t[2]), "log_prob": float(t[4]), "gc": [], "states": []}
            elif t[0] == "gc":
                cur["gc"].append([int(t[1]), int(t[2])])
            elif t[0] == "state":
                cur["states"].append([int(v) for v in t[1:5]])
            elif t[0] == "sample":
                samples.append({"log_prob": float(t[3]), "states": []})
            elif t[0] == "sstate":
                samples[-1]["states"].append([int(v) for v in t[1:5]])
        cur["states"] = condense(cur["states"])
        for sm in samples:
            sm["states"] = condense(sm["states"])
        return cur, samples


def main():
    seq = "".join(l.strip() for l in gzip.open(REF + "/examples/chr2L/chr2L.sm.fa.gz", "rt") if l[0] != ">")
    dna = seq[START:START + LEN]
    name = "chr2L_%d_%d" % (START, LEN)
    fa = os.path.join(HERE, "fly_softmask_window.fa")
    synth.write_fasta(fa, [dna], [name])
    with tempfile.TemporaryDirectory() as td:
        blob = os.path.join(td, "b")
        subprocess.run([AUGDUMP] + ARGS + [fa], env=dict(ENV, AUGDUMP_PARAMS=blob), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with lzma.open(os.path.join(HERE, "fly_softmask_utr.params.xz"), "wb", preset=9) as f:
            f.write(open(blob, "rb").read())
    out = {"masked": run_one(name, dna)[0], "unmasked": run_one(name, dna.upper())[0]}
    json.dump(out, open(os.path.join(HERE, "ref_paths_softmask.json"), "w"), separators=(",", ":"))
    vit, samples = run_one(name, dna, extra=("--sample=100", "--alternatives-from-sampling=true"))
    with gzip.open(os.path.join(HERE, "ref_samples_softmask.json.gz"), "wt") as f:
        json.dump({name: {"viterbi": vit, "samples": samples}}, f, separators=(",", ":"))
    print("softmasking golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

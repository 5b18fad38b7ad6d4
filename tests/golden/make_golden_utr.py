#!/usr/bin/env python
"""Golden fixtures for the 71-state model (--UTR=on), from the UNMODIFIED reference (build container only).

Writes
  human_utr.params.xz    parameter blob of  augustus --species=human --UTR=on --softmasking=0
  aug_utr_on.gff         the reference's OWN expected output of that command on examples/example.fa
                         (tests/short/examples/expected_results/test_utr_on/aug_utr_on.gff), a golden vector of its test suite
  ref_paths_utr.json     Viterbi paths (condensed) + scores + GC stairs from the reference for example.fa, synthetic windows
                         (50 kb x 4, one 200 kb = BASELINE.json configs[3] size, 4 short ones) and real_windows.fa
  ref_samples_utr.json.gz  the 99 paths of --sample=100 --alternatives-from-sampling=true on the first sequence of example.fa
One reference process per window: UtrModel keeps its TSS memo across sequences of equal length (utrmodel.cc:749-752).
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

ARGS = ["--species=human", "--UTR=on", "--softmasking=0"]


def read_fasta(fn):
    out, name, buf = [], None, []
    for line in open(fn):
        if line.startswith(">"):
            if name is not None:
                out.append((name, "".join(buf)))
            name, buf = line[1:].split()[0], []
        else:
            buf.append(line.strip())
    out.append((name, "".join(buf)))
    return out


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
    with tempfile.TemporaryDirectory() as td:
        blob = os.path.join(td, "b")
        subprocess.run([AUGDUMP] + ARGS + [REF + "/examples/example.fa"], env=dict(ENV, AUGDUMP_PARAMS=blob), check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with lzma.open(os.path.join(HERE, "human_utr.params.xz"), "wb", preset=9) as f:
            f.write(open(blob, "rb").read())
    open(os.path.join(HERE, "aug_utr_on.gff"), "w").write(open(REF + "/tests/short/examples/expected_results/test_utr_on/aug_utr_on.gff").read())
    out = {}
    out["example"] = [run_one(n, s)[0] for n, s in read_fasta(os.path.join(HERE, "example.fa"))]
    out["real"] = [run_one(n, s)[0] for n, s in read_fasta(os.path.join(HERE, "real_windows.fa"))]
    out["synthetic50k"] = [run_one("w%d" % i, synth.window(i, 50000))[0] for i in range(4)]
    out["synthetic200k"] = [run_one("w200k_0", synth.window(0, 200000))[0]]
    shorts = [(100, 3000), (101, 1500), (102, 700), (103, 120)]
    out["synthetic_short"] = [run_one("short%d_%d" % s, synth.window(*s))[0] for s in shorts]
    out["synthetic_short_spec"] = shorts
    json.dump(out, open(os.path.join(HERE, "ref_paths_utr.json"), "w"), separators=(",", ":"))
    n, s = read_fasta(os.path.join(HERE, "example.fa"))[0]
    vit, samples = run_one(n, s, extra=("--sample=100", "--alternatives-from-sampling=true"))
    with gzip.open(os.path.join(HERE, "ref_samples_utr.json.gz"), "wt") as f:
        json.dump({n: {"viterbi": vit, "samples": samples}}, f, separators=(",", ":"))
    print("UTR golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

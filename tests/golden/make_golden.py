#!/usr/bin/env python
"""Regenerate the golden fixtures from the UNMODIFIED reference (run in the build container only).

Needs /root/reference and oracle/_ref/augdump (``make -C oracle ref tools``).  Writes
  human.params.xz        parameter blob exported after StateModel::readAllParameters()
  example.fa             the reference's examples/example.fa (test input, 11.8 kb)
  real_windows.fa        three windows of examples/chr2L (two GC classes under human parameters, one with N)
  ref_paths.json         the reference's own Viterbi paths (condensed) + scores + GC stairs for
                         example.fa, synthetic windows 0..15 (50 kb), 4 short synthetic windows and real_windows.fa
The reference command per input is the one BASELINE.md lists: augustus --species=human --softmasking=0.
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
from augustus_b200 import synth  # noqa: E402

REF = "/root/reference"
AUGDUMP = os.path.join(ROOT, "oracle", "_ref", "augdump")
ENV = dict(os.environ, AUGUSTUS_CONFIG_PATH=REF + "/config")


def condense(states):
    out = []
    for t, b, e, tr in states:
        coding = (1 <= t <= 8) or (36 <= t <= 43)
        if out and out[-1][0] == t and not coding:
            out[-1][2] = e
            out[-1][3] |= tr
        else:
            out.append([t, b, e, tr])
    return out


def run_ref(fasta, species="human", extra=()):
    with tempfile.TemporaryDirectory() as td:
        pf = os.path.join(td, "p")
        env = dict(ENV, AUGDUMP_PATH=pf)
        subprocess.run([AUGDUMP, "--species=" + species, "--softmasking=0"] + list(extra) + [fasta], env=env, check=True,
                       stdout=subprocess.DEVNULL)
        res, cur = [], None
        for line in open(pf):
            t = line.split()
            if t[0] == "seq":
                cur = {"name": t[1], "length": int(t[2]), "log_prob": float(t[4]), "gc": [], "states": []}
                res.append(cur)
            elif t[0] == "gc":
                cur["gc"].append([int(t[1]), int(t[2])])
            elif t[0] == "state":
                cur["states"].append([int(v) for v in t[1:5]])
        for r in res:
            r["states"] = condense(r["states"])
        return res


def main():
    with tempfile.TemporaryDirectory() as td:
        blob = os.path.join(td, "human.blob")
        subprocess.run([AUGDUMP, "--species=human", "--softmasking=0", REF + "/examples/example.fa"],
                       env=dict(ENV, AUGDUMP_PARAMS=blob), check=True, stdout=subprocess.DEVNULL)
        with lzma.open(os.path.join(HERE, "human.params.xz"), "wb", preset=9) as f:
            f.write(open(blob, "rb").read())
    ex = open(REF + "/examples/example.fa").read()
    open(os.path.join(HERE, "example.fa"), "w").write(ex)
    seq = "".join(l.strip() for l in gzip.open(REF + "/examples/chr2L/chr2L.sm.fa.gz", "rt") if l[0] != ">")
    wins = [(5005000, 30000), (12010000, 30000), (21470000, 20000)]
    synth.write_fasta(os.path.join(HERE, "real_windows.fa"), [seq[a:a + n] for a, n in wins],
                      ["chr2L_%d_%d" % w for w in wins])
    out = {"example": run_ref(os.path.join(HERE, "example.fa")), "real": run_ref(os.path.join(HERE, "real_windows.fa"))}
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "s.fa")
        synth.write_fasta(fa, synth.windows(16, 50000))
        out["synthetic50k"] = run_ref(fa)
        shorts = [(100, 3000), (101, 1500), (102, 700), (103, 120)]
        synth.write_fasta(fa, [synth.window(i, n) for i, n in shorts], ["short%d_%d" % s for s in shorts])
        out["synthetic_short"] = run_ref(fa)
        out["synthetic_short_spec"] = shorts
    # second parameter set with a different signal geometry (ass_start 1 / ass_end 4, no exon-terminal content, d = 929, one
    # GC class): --species=fly --UTR=off --softmasking=0, the 47-state variant of BASELINE.json configs[2]
    with tempfile.TemporaryDirectory() as td:
        blob = os.path.join(td, "fly.blob")
        fa = os.path.join(HERE, "fly_window.fa")
        synth.write_fasta(fa, [seq[5000000:5060000]], ["chr2L_5000000_60000"])
        subprocess.run([AUGDUMP, "--species=fly", "--UTR=off", "--softmasking=0", fa], env=dict(ENV, AUGDUMP_PARAMS=blob), check=True, stdout=subprocess.DEVNULL)
        with lzma.open(os.path.join(HERE, "fly_noutr.params.xz"), "wb", preset=9) as f:
            f.write(open(blob, "rb").read())
    out["fly"] = run_ref(os.path.join(HERE, "fly_window.fa"), species="fly", extra=("--UTR=off",))
    json.dump(out, open(os.path.join(HERE, "ref_paths.json"), "w"), separators=(",", ":"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

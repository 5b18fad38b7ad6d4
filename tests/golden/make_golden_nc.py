#!/usr/bin/env python
"""Golden fixtures for the 83-state model with NcModel states (--UTR=on --nc=on), from the UNMODIFIED reference.

Writes
  human_nc.params.xz     parameter blob of  augustus --species=human --UTR=on --nc=on --softmasking=0
  ref_paths_nc.json      Viterbi paths (condensed) + scores of the reference for example.fa, synthetic windows and real_windows.fa, and
                         for every window the number of non-zero cells per nc state of the reference's Viterbi matrix
  ref_samples_nc.json.gz the 99 paths of --sample=100 --alternatives-from-sampling=true on the second sequence of example.fa
One reference process per window.
"""
import gzip
import json
import lzma
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from augustus_b200 import synth  # noqa: E402
from make_golden import AUGDUMP, ENV, REF, condense  # noqa: E402
from make_golden_utr import read_fasta  # noqa: E402

ARGS = ["--species=human", "--UTR=on", "--nc=on", "--softmasking=0"]
S = 83


def run_one(name, dna, extra=(), matrix=True):
    with tempfile.TemporaryDirectory() as td:
        fa, pf, mf = os.path.join(td, "w.fa"), os.path.join(td, "p"), os.path.join(td, "m")
        synth.write_fasta(fa, [dna], [name])
        env = dict(ENV, AUGDUMP_PATH=pf)
        if matrix:
            env["AUGDUMP_MATRIX"] = mf
        subprocess.run([AUGDUMP] + ARGS + list(extra) + [fa], env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
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
        if matrix:
            V = np.fromfile(mf + ".1.vit").reshape(-1, S)
            cur["nc_cells"] = [int(np.isfinite(V[:, s]).sum()) for s in range(71, S)]
            cur["last_column"] = [None if not np.isfinite(v) else float(v) for v in V[-1]]
        return cur, samples


def main():
    with tempfile.TemporaryDirectory() as td:
        blob = os.path.join(td, "b")
        subprocess.run([AUGDUMP] + ARGS + [REF + "/examples/example.fa"], env=dict(ENV, AUGDUMP_PARAMS=blob), check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with lzma.open(os.path.join(HERE, "human_nc.params.xz"), "wb", preset=9) as f:
            f.write(open(blob, "rb").read())
    out = {}
    out["example"] = [run_one(n, s)[0] for n, s in read_fasta(os.path.join(HERE, "example.fa"))]
    out["real"] = [run_one(n, s)[0] for n, s in read_fasta(os.path.join(HERE, "real_windows.fa"))]
    out["synthetic50k"] = [run_one("w%d" % i, synth.window(i, 50000))[0] for i in range(2)]
    shorts = [(100, 3000), (101, 1500), (102, 700), (103, 120)]
    out["synthetic_short"] = [run_one("short%d_%d" % s, synth.window(*s))[0] for s in shorts]
    out["synthetic_short_spec"] = shorts
    json.dump(out, open(os.path.join(HERE, "ref_paths_nc.json"), "w"), separators=(",", ":"))
    n, s = read_fasta(os.path.join(HERE, "example.fa"))[1]
    vit, samples = run_one(n, s, extra=("--sample=100", "--alternatives-from-sampling=true"), matrix=False)
    with gzip.open(os.path.join(HERE, "ref_samples_nc.json.gz"), "wt") as f:
        json.dump({n: {"viterbi": vit, "samples": samples}}, f, separators=(",", ":"))
    print("nc golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

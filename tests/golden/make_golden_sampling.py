#!/usr/bin/env python
"""Golden vectors for the forward + sampling path, generated with the UNMODIFIED reference (build container only).

One reference process per window (fresh glibc rand() stream, seed 1), command of BASELINE.md config 5:
    augustus --species=human --softmasking=0 --sample=100 --alternatives-from-sampling=true
Writes ref_samples.json.gz: per window the 99 sampled state paths (condensed) with ln pathemiProb, and a few
forward-matrix cells (column, state, ln value) for spot checks.
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from augustus_b200 import synth  # noqa: E402
from tests import util  # noqa: E402

REF = "/root/reference"
AUGDUMP = os.path.join(ROOT, "oracle", "_ref", "augdump")
ENV = dict(os.environ, AUGUSTUS_CONFIG_PATH=REF + "/config")
S = 47


def run(dna, species="human", extra=()):
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "w.fa"); synth.write_fasta(fa, [dna])
        subprocess.run([AUGDUMP, "--species=" + species, "--softmasking=0", "--sample=100", "--alternatives-from-sampling=true"] + list(extra) + [fa],
                       env=dict(ENV, AUGDUMP_PATH=os.path.join(td, "p"), AUGDUMP_MATRIX=os.path.join(td, "m")), check=True, stdout=subprocess.DEVNULL)
        samples, cur = [], None
        for line in open(os.path.join(td, "p")):
            t = line.split()
            if t[0] == "sample":
                cur = {"log_prob": float(t[3]), "states": []}; samples.append(cur)
            elif t[0] == "sstate":
                cur["states"].append([int(v) for v in t[1:5]])
        for s in samples:
            s["states"] = [list(x) for x in util.condense([tuple(x) for x in s["states"]])]
        F = np.fromfile(os.path.join(td, "m.1.fwd"), dtype=np.float64).reshape(len(dna), S)
        cols = sorted(set(np.linspace(1, len(dna) - 1, 40).astype(int).tolist()))
        cells = [[int(j), int(s), float(F[j, s])] for j in cols for s in range(S) if np.isfinite(F[j, s])]
        return {"length": len(dna), "samples": samples, "forward_cells": cells}


def main():
    wins = {"example_HS08198": util.read_fasta(os.path.join(HERE, "example.fa"))[1][1],
            "synthetic_301_20000": synth.window(301, 20000),
            "real_chr2L_5005000": util.read_fasta(os.path.join(HERE, "real_windows.fa"))[0][1]}
    out = {k: run(v) for k, v in wins.items()}
    out["fly_chr2L_5000000"] = run(util.read_fasta(os.path.join(HERE, "fly_window.fa"))[0][1], "fly", ("--UTR=off",))
    with gzip.open(os.path.join(HERE, "ref_samples.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print({k: len(v["samples"]) for k, v in out.items()})


if __name__ == "__main__":
    main()

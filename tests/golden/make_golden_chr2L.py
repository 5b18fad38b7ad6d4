#!/usr/bin/env python
"""Digests of the reference's own result on EVERY window of BASELINE.json configs[2]: examples/chr2L cut into 200 kb windows
stepping 150 kb (157 windows), ``augustus --species=fly`` with its defaults (UTR on = 71 states, softmasking on, sample=100), one
process per window.  For each window: ln of the Viterbi path probability, SHA-1 of the condensed Viterbi path and SHA-1 of the 99
condensed sampled paths, as produced by oracle/_ref/augdump (our driver linked with the UNMODIFIED reference objects).

Writes tests/golden/ref_chr2L_digests.json (small: the 31 Mbp of input stay in oracle/_ref/data/, copied there by oracle/Makefile).
Usage: make_golden_chr2L.py [first_window [last_window [workers]]]   (results are merged into the existing file)
"""
import concurrent.futures as cf
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from augustus_b200 import synth  # noqa: E402
from make_golden import AUGDUMP, ENV, condense  # noqa: E402
import bench  # noqa: E402

OUT = os.path.join(HERE, "ref_chr2L_digests.json")


def digest_states(states):
    """SHA-1 of the path as little-endian int32 rows (type, begin, end, truncated)."""
    import numpy as np
    return hashlib.sha1(np.asarray(states, dtype="<i4").reshape(-1, 4).tobytes()).hexdigest()


def digest_samples(samples):
    """SHA-1 over all sampled paths in order: per path an int32 state count, then its rows."""
    import numpy as np
    h = hashlib.sha1()
    for st in samples:
        h.update(np.asarray([len(st)], dtype="<i4").tobytes())
        h.update(np.asarray(st, dtype="<i4").reshape(-1, 4).tobytes())
    return h.hexdigest()


def run_window(args):
    idx, dna = args
    with tempfile.TemporaryDirectory() as td:
        fa, pf = os.path.join(td, "w.fa"), os.path.join(td, "p")
        synth.write_fasta(fa, [dna], ["chr2L_w%d" % idx])
        subprocess.run([AUGDUMP, "--species=fly", fa], env=dict(ENV, AUGDUMP_PATH=pf), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        vit, samples, logp = [], [], None
        for line in open(pf):
            t = line.split()
            if t[0] == "seq":
                logp = float(t[4])
            elif t[0] == "state":
                vit.append([int(v) for v in t[1:5]])
            elif t[0] == "sample":
                samples.append([])
            elif t[0] == "sstate":
                samples[-1].append([int(v) for v in t[1:5]])
    vit = condense(vit)
    samples = [condense(s) for s in samples]
    return idx, {"length": len(dna), "log_prob": logp, "n_states": len(vit), "viterbi_sha1": digest_states(vit), "n_samples": len(samples),
                 "sample_states": sum(len(s) for s in samples), "samples_sha1": digest_samples(samples)}


def main():
    wins = bench.chr2l_windows()
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    last = int(sys.argv[2]) if len(sys.argv) > 2 else len(wins) - 1
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, (os.cpu_count() or 2) - 1)
    have = json.load(open(OUT)) if os.path.exists(OUT) else {"window": 200000, "step": 150000, "args": "--species=fly", "windows": {}}
    todo = [(i, wins[i]) for i in range(first, last + 1) if str(i) not in have["windows"]]
    with cf.ThreadPoolExecutor(workers) as ex:
        for idx, rec in ex.map(run_window, todo):
            have["windows"][str(idx)] = rec
            json.dump(have, open(OUT, "w"), indent=0, sort_keys=True)
            print("window", idx, rec["n_states"], "states", rec["log_prob"], flush=True)


if __name__ == "__main__":
    main()

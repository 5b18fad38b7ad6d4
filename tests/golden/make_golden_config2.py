#!/usr/bin/env python
"""Reference digests of BASELINE.json configs[1] (run in the build container only; needs /root/reference and oracle/_ref/augdump).

Writes ref_config2_digests.json: for the synthetic 50 kb windows with global index 0..511 (augustus_b200.synth.window) the unmodified
reference's ln-probability and the SHA-1 of its condensed Viterbi path (`augustus --species=human --softmasking=0`, one record per
window, 8 windows per process).  bench.py checks the paths produced in its timed region against these (64 windows per rank at any
world size <= 8: rank r decodes the global windows r, r + world, ...), tests/test_gpu.py the first 64.
"""
import concurrent.futures as cf
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from augustus_b200 import synth  # noqa: E402
from make_golden import run_ref  # noqa: E402

N, PER = 512, 8


def digest(states):
    return hashlib.sha1(json.dumps([list(map(int, s)) for s in states], separators=(",", ":")).encode()).hexdigest()


def chunk(a):
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "s.fa")
        synth.write_fasta(fa, [synth.window(i, 50000) for i in range(a, a + PER)], ["w%d" % i for i in range(a, a + PER)])
        return [(a + k, r["log_prob"], digest(r["states"]), len(r["states"])) for k, r in enumerate(run_ref(fa))]


def main():
    out = {}
    with cf.ThreadPoolExecutor(max(1, (os.cpu_count() or 2) - 1)) as ex:
        for rows in ex.map(chunk, range(0, N, PER)):
            for i, lp, d, n in rows:
                out[str(i)] = {"log_prob": lp, "sha1": d, "n": n}
    json.dump({"window_len": 50000, "command": "augustus --species=human --softmasking=0", "windows": out},
              open(os.path.join(HERE, "ref_config2_digests.json"), "w"), indent=0, sort_keys=True)
    print("wrote", len(out), "digests")


if __name__ == "__main__":
    main()

"""Differential fuzz (CPU only): the device flavour of the kernel source on the 32-lane executor against the oracle on random windows —
composition shifts (GC-class boundaries), N runs, soft-masked runs, tiny windows; cells, paths and scores must be identical.
usage: fuzz_simt32.py [n_cases=100] [seed=1] [blob=human|human_utr|human_nc|fly_noutr|fly_softmask_utr]
FUZZ_SAMPLING=1 also compares 11 sampled paths per window (windows up to 4 kb); FUZZ_ONE_CLASS=1 keeps the composition fixed (no GC-class boundaries: isolates everything but the class-history memos of DESIGN.md 3.6)"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
blobname = sys.argv[3] if len(sys.argv) > 3 else "human"
blob = util.blob_bytes(blobname)
orc, emu = util.Oracle(blob), util.HostEmu(blob, simt32=True)
rng = random.Random(seed)


def piece(length, gc):
    at = (1 - gc) / 2
    return "".join(rng.choices("ACGT", weights=(at, gc / 2, gc / 2, at), k=length))


def make():
    L = rng.choice([2, 3, 5, 17, 60, 61, 62, 63, 64, 97, 300, 301, 1500, 4000, 9000, 14000])
    parts = []
    while sum(map(len, parts)) < L:
        parts.append(piece(rng.randint(1, max(1, L // 2)), 0.41 if os.environ.get("FUZZ_ONE_CLASS") else rng.choice([0.3, 0.41, 0.41, 0.5, 0.62, 0.7])))
    s = "".join(parts)[:L]
    s = list(s)
    for _ in range(rng.randint(0, 3)):                     # N runs
        a = rng.randrange(L); b = min(L, a + rng.randint(1, 40))
        for i in range(a, b): s[i] = "N"
    if "softmask" in blobname:
        for _ in range(rng.randint(0, 6)):                 # soft-masked runs
            a = rng.randrange(L); b = min(L, a + rng.randint(1, 600))
            for i in range(a, b): s[i] = s[i].lower()
    return "".join(s)


bad = 0; t0 = time.time()
for k in range(n):
    dna = make()
    r, e = orc.viterbi(dna, want_matrix=True), emu.decode(dna, want_cells=True)
    ok = e["status"] == r.get("status", 0) if "status" in r else True
    if e["status"] == 0:
        V, E = r["V"], e["cells"]
        ok = ok and e["states"] == r["condensed"] and e["log_prob"] == r["log_prob"]
        if set(dna.upper()) - {"N"}:                       # (a window without a single nucleotide is all intergenic: the dumps of column 0 differ by convention)
            ok = ok and ((V <= util.NEGT) == (E <= util.NEGT)).all() and (V[V > util.NEGT] == E[V > util.NEGT]).all()
    else:
        ok = ok and r["n"] < 0 if "n" in r else False
    if ok and os.environ.get("FUZZ_SAMPLING") and len(dna) <= 4000 and set(dna.upper()) - {"N"}:      # also the forward fill + sampling walks
        ns = 12
        so = orc.sample(dna, ns)["samples"]; se = emu.sample(dna, ns - 1)
        ok = se["status"] == 0 and len(se["samples"]) == ns - 1 and all(a["states"] == b["states"] for a, b in zip(se["samples"], so))
        if not ok: print("  (sampling differs)")
    if not ok:
        bad += 1
        print("MISMATCH case", k, "len", len(dna), "status", e["status"]); open("/tmp/fuzz_bad_%s_%d_%d.fa" % (blobname, seed, k), "w").write(">x\n" + dna + "\n")
print(blobname, "cases", n, "mismatches", bad, "seconds", round(time.time() - t0, 1))

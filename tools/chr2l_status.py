"""Decode all chr2L windows (fly defaults, sample=100) and print per-window status; usage: chr2l_status.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from augustus_b200 import Decoder
from tests import util
w = bench.chr2l_windows()
dec = Decoder(util.blob_bytes("fly_softmask_utr"), 0)
t0 = time.perf_counter()
vit, samp = dec.decode_batch_sampling_raw([x.encode() for x in w], 100)
dt = time.perf_counter() - t0
print("windows", len(w), "seconds", dt, "Mbp/s", sum(map(len, w)) / 1e6 / dt, "sweep ms", dec.last_sweep_ms)
bad = [(i, int(vit[1][i])) for i in range(len(w)) if vit[1][i]]
print("viterbi status != 0:", bad)
ss = samp[1].reshape(len(w), 99)
print("sample status != 0:", [(i, sorted(set(int(x) for x in ss[i] if x))) for i in range(len(w)) if ss[i].any()])
for i, _ in bad[:3]:
    x = w[i]; print(i, len(x), "N count", x.upper().count("N"), "lower", sum(c.islower() for c in x))

"""Sampling kernel timing: usage prof_sampling.py [n_windows=1000] [nsample=100] [blob=human] [wlen=50000]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augustus_b200 import Decoder, synth
from tests import util
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 100
blob = sys.argv[3] if len(sys.argv) > 3 else "human"
wlen = int(sys.argv[4]) if len(sys.argv) > 4 else 50000
wins = [w.encode() for w in synth.windows_parallel(n, wlen)]
dec = Decoder(util.blob_bytes(blob), 0)
dec.decode_batch_sampling_raw(wins[:8], ns)
t0 = time.perf_counter()
vit, samp = dec.decode_batch_sampling_raw(wins, ns)
dt = time.perf_counter() - t0
print("blob", blob, "windows", n, "x", wlen, "nsample", ns, "e2e s", round(dt, 3), "sweep+sampling kernel ms", round(dec.last_sweep_ms, 1), "Mbp/s", round(n * wlen / 1e6 / dt, 2),
      "bad", int(vit[1].any()) + int(samp[1].any()))

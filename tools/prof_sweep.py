"""Profiling driver: stage N synthetic windows and run the kernels a few times (used under ncu).
usage: prof_sweep.py [n_windows=592] [reps=2] [blob=human|human_utr|fly_noutr] [window_len=50000]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augustus_b200 import Decoder, synth
from tests import util

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
blob = sys.argv[3] if len(sys.argv) > 3 else "human"
wlen = int(sys.argv[4]) if len(sys.argv) > 4 else 50000
wins = [w.encode() for w in synth.windows_parallel(n, wlen)]
dec = Decoder(util.blob_bytes(blob), 0)
dec.stage(wins)
for i in range(reps):
    dec.run_staged()
ps = dec.fetch_staged()
print("blob", blob, "windows", n, "x", wlen, "sweep ms", dec.last_sweep_ms, "Mbp/s (sweep)", n * wlen / 1e3 / dec.last_sweep_ms, "ok", all(p.status == 0 for p in ps))

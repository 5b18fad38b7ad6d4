"""Profiling driver: stage N synthetic 50 kb windows and run the kernels a few times (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from augustus_b200 import Decoder, synth
from tests import util

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dec = Decoder(util.blob_bytes(), 0)
wins = [w.encode() for w in synth.windows_parallel(n, 50000)]
dec.stage(wins)
for i in range(reps):
    dec.run_staged()
ps = dec.fetch_staged()
print("windows", n, "sweep ms", dec.last_sweep_ms, "ok", all(p.status == 0 for p in ps))

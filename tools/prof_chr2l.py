"""config-3 timing split: chr2L windows (fly defaults) with nsample = 2 (forward fill + one walk) and 100 (99 walks).
usage: prof_chr2l.py [n_windows=157] [nsample ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from augustus_b200 import Decoder
from tests import util
n = int(sys.argv[1]) if len(sys.argv) > 1 else 157
nss = [int(x) for x in sys.argv[2:]] or [2, 100]
w = [x.encode() for x in bench.chr2l_windows()[:n]]
dec = Decoder(util.blob_bytes("fly_softmask_utr"), 0)
dec.decode_batch_sampling_raw(w[:2], 2)
t0 = time.perf_counter(); out = dec.decode_batch_raw(w); dt = time.perf_counter() - t0
print("windows", len(w), "viterbi only: e2e s", round(dt, 2), "sweep ms", round(dec.last_sweep_ms, 1), flush=True)
for ns in nss:
    t0 = time.perf_counter()
    vit, samp = dec.decode_batch_sampling_raw(w, ns)
    dt = time.perf_counter() - t0
    print("windows", len(w), "nsample", ns, "e2e s", round(dt, 2), "sweep+walk kernel ms", round(dec.last_sweep_ms, 1), "bad", int(vit[1].any()) + int(samp[1].any()), flush=True)

"""CPU test of the multi-GPU host logic with world_size 2 over gloo: block-cyclic sharding of the window list,
one variable-length gather of the results to rank 0, merge back into window order.  The per-window 'decode' is the
oracle (test infrastructure), standing in for the CUDA decode that needs a GPU."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from augustus_b200 import shard, synth
from tests import util

LENS = [3000, 1200, 2500, 700, 1800]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _raw_from_oracle(orc, wins):
    n, st, lp, off, b, e, t, tr = [], [], [], [], [], [], [], []
    for w in wins:
        o = orc.viterbi(w)
        off.append(len(b)); n.append(len(o["condensed"])); st.append(0); lp.append(o["log_prob"])
        for x in o["condensed"]:
            t.append(x[0]); b.append(x[1]); e.append(x[2]); tr.append(x[3])
    return (np.array(n, np.int32), np.array(st, np.int32), np.array(lp), np.array(off, np.int64), np.array(b, np.int32),
            np.array(e, np.int32), np.array(t, np.uint8), np.array(tr, np.uint8))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = util.Oracle(util.blob_bytes())
    idx = shard.shard_indices(len(LENS), rank, world)
    wins = [synth.window(500 + i, LENS[i]) for i in idx]
    vec = shard.pack_paths(*_raw_from_oracle(orc, wins))
    got = shard.gather_to_rank0(vec)
    if rank == 0:
        per_rank = [shard.unpack_paths(v) for v in got]
        merged = shard.merge_in_window_order(per_rank, len(LENS), world)
        q.put([(s, lp, st.tolist()) for s, lp, st in merged])
    dist.destroy_process_group()


def test_sharding_and_gather_world2():
    assert shard.shard_indices(5, 0, 2) == [0, 2, 4] and shard.shard_indices(5, 1, 2) == [1, 3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc = util.Oracle(util.blob_bytes())
    for i, (status, lp, states) in enumerate(merged):
        o = orc.viterbi(synth.window(500 + i, LENS[i]))
        assert status == 0 and lp == o["log_prob"]
        assert [(t, b, e, tr) for b, e, t, tr in states] == o["condensed"]


def test_pack_roundtrip():
    raw = (np.array([2, 0, 1], np.int32), np.array([0, 6, 0], np.int32), np.array([-1.5, 0.0, -3.25]), np.array([0, 2, 2], np.int64),
           np.array([1, 5, 1], np.int32), np.array([4, 9, 7], np.int32), np.array([0, 2, 0], np.uint8), np.array([0, 2, 1], np.uint8))
    out = shard.unpack_paths(shard.pack_paths(*raw))
    assert [o[0] for o in out] == [0, 6, 0] and [o[1] for o in out] == [-1.5, 0.0, -3.25]
    assert out[0][2].tolist() == [[1, 4, 0, 0], [5, 9, 2, 2]] and out[2][2].tolist() == [[1, 7, 0, 1]]


def test_deal_windows_by_measured_rate():
    """bench.py re-deals the world x M windows in proportion to each rank's measured sweep rate (the same kernel ran 17 % slower on
    three GPUs of the round-1 node): every window exactly once, counts proportional, block-cyclic base kept where possible, and the
    same answer on every rank (pure function of the gathered rates)."""
    import bench
    world, M = 8, 1000
    rates = [3.34, 3.39, 2.91, 3.42, 2.91, 3.38, 2.86, 3.36]
    deal = bench.deal_windows(world, M, rates)
    flat = sorted(g for d in deal for g in d)
    assert flat == list(range(world * M))
    want = [world * M * r / sum(rates) for r in rates]
    assert all(abs(len(d) - w) <= 1.0 for d, w in zip(deal, want))
    for r, d in enumerate(deal):                       # a slow rank keeps a prefix of its block-cyclic share
        own = [g for g in d if g % world == r]
        assert own == list(range(r, r + world * len(own), world))
    # equal rates (within 2 %): nothing moves
    assert bench.deal_windows(4, 100, [1.0, 1.01, 0.995, 1.0]) == [list(range(r, 400, 4)) for r in range(4)]
    assert bench.deal_windows(1, 10, [5.0]) == [list(range(10))]
    # the time of the slowest rank with the new deal is what a perfect split would give (within one window)
    t = max(len(d) / r for d, r in zip(deal, rates))
    assert t <= world * M / sum(rates) + 1.0 / min(rates)

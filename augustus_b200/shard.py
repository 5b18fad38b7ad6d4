"""Multi-GPU sharding of independent windows.

Windows share only read-only parameters, so the path shards with no data-path collective: window i of the
global list goes to rank i mod G (block-cyclic, as the reference's own scale-out splits a genome into chunks and
runs one process per chunk, scripts/splitMfasta.pl + join_aug_pred.pl).  The only communication is ONE gather of the
final path arrays to rank 0 (variable length -> sizes first, then padded payloads), over NCCL on GPUs or gloo on CPU.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_indices(n_windows: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_windows, world))


def pack_paths(n, status, logp, offset, begin, end, typ, trunc) -> np.ndarray:
    """Flatten one rank's raw decode result (Decoder.decode_batch_raw) into a single int32 vector:
    [n_windows, (n_i, status_i, logp_bits_lo, logp_bits_hi) per window, (begin, end, type, trunc) per state]."""
    nw = len(n)
    head = np.empty((nw, 4), dtype=np.int32)
    head[:, 0] = n
    head[:, 1] = status
    head[:, 2:4] = np.asarray(logp, dtype=np.float64).view(np.int32).reshape(nw, 2)
    # states in window order
    order = np.concatenate([np.arange(o, o + k) for o, k in zip(offset, n)]) if nw and int(np.sum(n)) else np.zeros(0, dtype=np.int64)
    body = np.stack([begin[order], end[order], typ[order].astype(np.int32), trunc[order].astype(np.int32)], axis=1).astype(np.int32)
    return np.concatenate([np.array([nw], dtype=np.int32), head.ravel(), body.ravel()])


def unpack_paths(vec: np.ndarray) -> List[Tuple[int, float, np.ndarray]]:
    """Inverse of pack_paths: list of (status, log_prob, states[n,4])."""
    nw = int(vec[0])
    head = vec[1:1 + 4 * nw].reshape(nw, 4)
    body = vec[1 + 4 * nw:].reshape(-1, 4)
    out, pos = [], 0
    for i in range(nw):
        k = int(head[i, 0])
        lp = float(np.ascontiguousarray(head[i, 2:4]).view(np.float64)[0])
        out.append((int(head[i, 1]), lp, body[pos:pos + k].copy()))
        pos += k
    return out


def gather_to_rank0(vec: np.ndarray, device: Optional[str] = None):
    """One variable-length gather of per-rank result vectors to rank 0.  Returns the list of vectors (rank order) on
    rank 0 and None elsewhere.  Works with any initialised torch.distributed backend (nccl: device='cuda')."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device(device) if device else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.int32)).to(dev)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros(mx, dtype=torch.int32, device=dev)
    pad[: t.numel()] = t
    out = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, out, dst=0)
    if rank != 0:
        return None
    return [o[: int(s)].cpu().numpy() for o, s in zip(out, sizes)]


def merge_in_window_order(per_rank: Sequence[List], n_windows: int, world: int) -> List:
    """Interleave per-rank result lists back into global window order (inverse of shard_indices)."""
    merged = [None] * n_windows
    for r, items in enumerate(per_rank):
        for k, item in enumerate(items):
            merged[r + k * world] = item
    return merged

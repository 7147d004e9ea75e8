"""Multi-GPU plumbing: one process per GPU, bubbles sharded by rank, KMC index replicated.

The path partitions by independent superbubbles: every rank takes a contiguous range of the
host's bubble stream and runs the whole hot path on it; there is NO data-path collective.  torch.distributed is
used only to agree on timings / totals (max over ranks, sum of work) and, when a caller wants one result on
rank 0, to gather per-rank result sizes so that shards concatenate in range order.  Works over NCCL (GPU
tensors) and gloo (CPU tensors; the world_size-2 tests).
"""
from __future__ import annotations

import numpy as np


def bubble_range(n_bubbles: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [b0, b1) of rank `rank`; ranges are in rank order and cover [0, n) exactly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_bubbles), world)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def shard_batch(bases: np.ndarray, seq_off: np.ndarray, bubble_off: np.ndarray, rank: int, world: int):
    """The rank's slice of a flat bubble batch (include/pf_types.h layout), offsets rebased to 0."""
    n = len(bubble_off) - 1
    b0, b1 = bubble_range(n, rank, world)
    s0, s1 = int(bubble_off[b0]), int(bubble_off[b1])
    c0, c1 = int(seq_off[s0]), int(seq_off[s1])
    return (bases[c0:c1], (seq_off[s0:s1 + 1] - seq_off[s0]).astype(np.uint64),
            (bubble_off[b0:b1 + 1] - bubble_off[b0]).astype(np.uint32))


def reduce_step(times_ms, work, device=None):
    """(max over ranks of each time, sum over ranks of each work counter); identity when not initialised."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [float(x) for x in w.tolist()]


def gather_sizes(local_sizes, device=None) -> np.ndarray:
    """[world, len(local_sizes)] table of every rank's counters (e.g. bubbles, rows bytes, variable sites)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(local_sizes), dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.cpu().numpy()[None, :]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def concat_msa(shards: list) -> dict:
    """Concatenates per-rank pf_msa_batch_t dumps (dicts of numpy arrays, in rank order) into one batch."""
    out = {"n_bubbles": int(sum(s["n_bubbles"] for s in shards))}
    for key in ("status", "n_rows", "aln_len", "rows", "var_col", "var_kind", "cls", "ilen"):
        out[key] = np.concatenate([s[key] for s in shards])
    for key in ("rows_off", "var_off", "cls_off", "ilen_off"):
        parts, base = [], 0
        for i, s in enumerate(shards):
            o = s[key].astype(np.uint64)
            parts.append((o[:-1] if i + 1 < len(shards) else o) + np.uint64(base))
            base += int(o[-1])
        out[key] = np.concatenate(parts)
    return out

"""Multi-GPU driver: one process per GPU, the proposal batch sharded by contiguous row blocks.

Samples are independent under inference-mode BN, so the K-step refinement needs NO collective.
NCCL is used only around the accept-reject stage:
  1. all_gather of the per-rank scores (4 B/sample) so that every rank evaluates the GLOBAL DRS threshold /
     MH chain redundantly and bit-identically to a 1-GPU run,
  2. all_gather of accepted-row counts + a padded all_gather of the accepted rows each rank owns,
  3. all_reduce of the acceptance / score statistics.
The functions below work on CPU tensors with the gloo backend as well (that is how the host logic is tested).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_global, rank, world_size):
    """Contiguous block [lo, hi) of rows owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(n_global, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_scores(local_scores, group=None):
    """[n_local] -> [n_global] in rank order (equal shard sizes use all_gather_into_tensor, ragged ones pad)."""
    rank, ws = world()
    if ws == 1:
        return local_scores
    n = torch.tensor([local_scores.numel()], dtype=torch.int64, device=local_scores.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    if min(sizes) == m:
        out = torch.empty(ws * m, dtype=local_scores.dtype, device=local_scores.device)
        dist.all_gather_into_tensor(out, local_scores.contiguous(), group=group)
        return out
    pad = torch.zeros(m, dtype=local_scores.dtype, device=local_scores.device)
    pad[:local_scores.numel()] = local_scores
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def owned_slice(src_global, lo, hi):
    """Positions of an ascending global source-row list that fall in [lo, hi) -> (start, stop) of the run."""
    if src_global.numel() == 0:
        return 0, 0
    start = int(torch.searchsorted(src_global, torch.tensor(lo, device=src_global.device, dtype=src_global.dtype)))
    stop = int(torch.searchsorted(src_global, torch.tensor(hi, device=src_global.device, dtype=src_global.dtype)))
    return start, stop


def gather_accepted(local_rows, src_global, bounds, group=None):
    """All-gather(v) of the accepted rows in global order.

    ``local_rows`` [n_local, ...] are this rank's samples; ``bounds[r] = (lo, hi)`` are the global row blocks of every
    rank (``shard_bounds``); ``src_global`` is the ascending list of accepted / emitted GLOBAL row ids, identical on
    every rank because every rank evaluated the same global chain.  Ownership is a pure function of that list, so
    all counts are known everywhere without a collective; one padded all_gather moves the rows.
    Returns [len(src_global), ...] on every rank.
    """
    rank, ws = world()
    lo, hi = bounds[rank]
    start, stop = owned_slice(src_global, lo, hi)
    mine = local_rows[(src_global[start:stop] - lo).long()] if stop > start else local_rows[:0]
    if ws == 1:
        return mine
    counts = []
    for rlo, rhi in bounds:
        a, b = owned_slice(src_global, rlo, rhi)
        counts.append(b - a)
    m = max(counts) if counts else 0
    if m == 0:
        return local_rows[:0]
    pad = torch.zeros((m,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[:mine.shape[0]] = mine
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def reduce_stats(n_accepted, score_sum, score_max, group=None):
    """(sum, sum, max) all-reduce of the acceptance / score statistics; returns python floats."""
    rank, ws = world()
    dev = score_sum.device if isinstance(score_sum, torch.Tensor) else "cpu"
    s = torch.tensor([float(n_accepted), float(score_sum)], dtype=torch.float64, device=dev)
    m = torch.tensor([float(score_max)], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return float(s[0]), float(s[1]), float(m[0])

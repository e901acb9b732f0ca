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


def gather_scores(local_scores, group=None, sizes=None):
    """[n_local] -> [n_global] in rank order (equal shard sizes use all_gather_into_tensor, ragged ones pad).

    ``sizes`` = per-rank shard sizes when the caller knows them (``shard_bounds``): skips the size exchange and its
    host synchronisation."""
    rank, ws = world()
    if ws == 1:
        return local_scores
    if sizes is None:
        n = torch.tensor([local_scores.numel()], dtype=torch.int64, device=local_scores.device)
        got = torch.empty(ws, dtype=torch.int64, device=local_scores.device)
        dist.all_gather_into_tensor(got, n, group=group)
        sizes = got.tolist()
    sizes = [int(v) for v in sizes]
    m = max(sizes)
    if min(sizes) == m:
        out = torch.empty(ws * m, dtype=local_scores.dtype, device=local_scores.device)
        dist.all_gather_into_tensor(out, local_scores.contiguous(), group=group)
        return out
    pad = torch.zeros(m, dtype=local_scores.dtype, device=local_scores.device)
    pad[:local_scores.numel()] = local_scores
    out = torch.empty(ws * m, dtype=local_scores.dtype, device=local_scores.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m:r * m + c] for r, c in enumerate(sizes)])


def owned_runs(src_global, bounds):
    """For an ascending global source-row list and contiguous ascending row blocks ``bounds[r] = (lo, hi)``: the
    positions [start_r, stop_r) of the rows each rank owns.  One searchsorted, one host read-back."""
    if src_global.numel() == 0:
        return [(0, 0) for _ in bounds]
    edges = torch.tensor([lo for lo, _ in bounds] + [bounds[-1][1]], device=src_global.device, dtype=src_global.dtype)
    pos = torch.searchsorted(src_global, edges).tolist()
    return [(pos[r], pos[r + 1]) for r in range(len(bounds))]


def owned_slice(src_global, lo, hi):
    """Positions of an ascending global source-row list that fall in [lo, hi) -> (start, stop) of the run."""
    return owned_runs(src_global, [(lo, hi)])[0]


def gather_accepted(local_rows, src_global, bounds, group=None):
    """All-gather(v) of the accepted rows in global order.

    ``local_rows`` [n_local, ...] are this rank's samples; ``bounds[r] = (lo, hi)`` are the global row blocks of every
    rank (``shard_bounds``: contiguous, ascending); ``src_global`` is the ascending list of accepted / emitted GLOBAL
    row ids, identical on every rank because every rank evaluated the same global chain.  Ownership is a pure function
    of that list, so all counts are known everywhere without a collective; one padded all_gather moves the rows.
    Returns [len(src_global), ...] on every rank.
    """
    rank, ws = world()
    runs = owned_runs(src_global, bounds)
    lo = bounds[rank][0]
    start, stop = runs[rank]
    mine = local_rows[(src_global[start:stop] - lo).long()] if stop > start else local_rows[:0]
    if ws == 1:
        return mine
    counts = [b - a for a, b in runs]
    m = max(counts) if counts else 0
    if m == 0:
        return local_rows[:0]
    pad = torch.zeros((m,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[:mine.shape[0]] = mine
    out = torch.empty((ws * m,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if min(counts) == m:
        return out
    return torch.cat([out[r * m:r * m + c] for r, c in enumerate(counts)])


def gather_accepted_async(local_rows, emit_global, count, lo, hi, group=None):
    """``gather_accepted`` with NO host synchronisation (device-resident counts).

    ``emit_global`` [cap] int32 is the (padded) ascending list of emitted GLOBAL row ids and ``count`` [1] the number
    of valid entries -- identical on every rank, because every rank evaluated the same global chain
    (``IndependenceSampler.select_async`` on the gathered scores).  This rank owns global rows [lo, hi).
    Every rank fills the positions whose source row it owns into a zero [cap, ...] buffer; ONE all-reduce(SUM) on the
    int32 view of that buffer merges the disjoint contributions (integer addition of zero bits is exact, also for
    -0.0 and NaN payloads), so the result is bit-identical to the single-GPU gather.  The payload is cap x row bytes
    whatever the world size.  Returns (rows [cap, ...], count); rows past ``count`` are zero.
    """
    rank, ws = world()
    cap = emit_global.numel()
    pos = torch.arange(cap, device=emit_global.device)
    src = emit_global.long()
    mine = (pos < count.reshape(()).long()) & (src >= lo) & (src < hi)
    n_local = local_rows.shape[0]
    if n_local == 0:
        out = torch.zeros((cap,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    else:
        local_idx = (src - lo).clamp_(0, n_local - 1)
        shape = (cap,) + (1,) * (local_rows.dim() - 1)
        out = torch.where(mine.view(shape), local_rows[local_idx], torch.zeros((), dtype=local_rows.dtype,
                                                                             device=local_rows.device))
    if ws > 1:
        out = out.contiguous()
        dist.all_reduce(out.view(torch.int32) if out.dtype == torch.float32 else out, op=dist.ReduceOp.SUM, group=group)
    return out, count


def reduce_stats_async(n_accepted, score_sum, score_max, group=None):
    """Acceptance / score statistics as ONE device tensor [3] = (sum n, sum score, max score), float64; no host sync
    (``n_accepted`` may be a device tensor).  One all_gather of 24 bytes per rank."""
    rank, ws = world()
    dev = score_sum.device
    mine = torch.stack([torch.as_tensor(v, device=dev).to(torch.float64).reshape(()) for v in (n_accepted, score_sum, score_max)])
    if ws == 1:
        return mine
    allv = torch.empty(ws * 3, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.view(ws, 3)
    return torch.stack([allv[:, 0].sum(), allv[:, 1].sum(), allv[:, 2].max()])


def reduce_stats(n_accepted, score_sum, score_max, group=None):
    """(sum, sum, max) all-reduce of the acceptance / score statistics; returns python floats."""
    rank, ws = world()
    dev = score_sum.device if isinstance(score_sum, torch.Tensor) else "cpu"
    if ws == 1:
        return float(n_accepted), float(score_sum), float(score_max)
    # one exchange: every rank contributes (n, sum, max); sums and the max are formed locally in rank order
    mine = torch.stack([torch.as_tensor(v, dtype=torch.float64, device=dev).reshape(()) for v in (n_accepted, score_sum, score_max)])
    allv = torch.empty(ws * 3, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.view(ws, 3).cpu()
    return float(allv[:, 0].sum()), float(allv[:, 1].sum()), float(allv[:, 2].max())

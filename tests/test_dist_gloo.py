"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a batch, gather scores, and all-gather(v) the
accepted rows; the result must equal the single-process result bit-for-bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "collaborative-gan-sampling_b200")


def _worker(rank, world, port, n_global, ragged, ret):
    sys.path.insert(0, PKG)
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cgs import dist as D
    from oracle import sampling_np as snp
    rng = np.random.RandomState(0)
    scores = rng.beta(2, 5, size=n_global).astype(np.float32)
    rows = rng.randn(n_global, 3).astype(np.float32)
    u = rng.rand(n_global)
    bounds = [D.shard_bounds(n_global, r, world) for r in range(world)]
    assert ragged == (len({hi - lo for lo, hi in bounds}) > 1)
    lo, hi = bounds[rank]
    local_scores = torch.from_numpy(scores[lo:hi])
    local_rows = torch.from_numpy(rows[lo:hi])
    g = D.gather_scores(local_scores)
    assert torch.equal(g, torch.from_numpy(scores))
    # every rank runs the global chain redundantly on identical inputs -> identical emitted rows
    emit, _, _, _ = snp.mh_chain(g.numpy().reshape(-1, 1), u, np.float32(0.4), 1, 3, 0)
    acc = D.gather_accepted(local_rows, torch.from_numpy(emit), bounds)
    assert torch.equal(acc, torch.from_numpy(rows[emit]))
    # device-resident variant (no host read-backs): padded emit list + count, one int32 all-reduce merges the rows
    rows_nz = rows.copy()
    rows_nz[::7, 1] = -0.0                                        # sign bit of a negative zero must survive the merge
    cap = len(emit) + 5
    emit_pad = torch.zeros(cap, dtype=torch.int32)
    emit_pad[:len(emit)] = torch.from_numpy(emit).int()
    emit_pad[len(emit):] = 12345678                               # garbage past the count must be ignored
    cnt = torch.tensor([len(emit)], dtype=torch.int32)
    acc2, cnt2 = D.gather_accepted_async(torch.from_numpy(rows_nz[lo:hi]), emit_pad, cnt, lo, hi)
    assert int(cnt2) == len(emit) and acc2.shape[0] == cap
    assert np.array_equal(acc2[:len(emit)].numpy().view(np.int32), rows_nz[emit].view(np.int32))
    assert not acc2[len(emit):].any()
    st = D.reduce_stats_async(torch.tensor(hi - lo), local_scores.double().sum(), local_scores.max())
    assert float(st[0]) == n_global and float(st[2]) == float(scores.max())
    n, ssum, smax = D.reduce_stats(hi - lo, float(local_scores.double().sum()), float(local_scores.max()))
    assert n == n_global and abs(ssum - float(scores.astype(np.float64).sum())) < 1e-9 and smax == float(scores.max())
    if rank == 0:
        ret.put(int(acc.shape[0]))
    dist.barrier()
    dist.destroy_process_group()


def _run(n_global, ragged, port):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, ragged, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) > 0


def test_two_rank_gather_equal_shards():
    _run(512, False, 29611)


def test_two_rank_gather_ragged_shards():
    _run(515, True, 29612)


def test_shard_bounds_cover_exactly():
    sys.path.insert(0, PKG)
    from cgs import dist as D
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            b = [D.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))

"""§8(f) rows: proposal head on the device and the accept-reject fill-up driver."""
import numpy as np
import pytest
import torch

from oracle import nets as onets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,math,tol", [("mnist", "fp32", 2e-5), ("mnist", "tf32", 3e-3), ("dcgan64_l1", "fp32", 2e-5),
                                           ("dcgan32_l3", "tf32", 3e-3)])
def test_proposal_head_matches_oracle(cgs_lib, cuda_device, name, math, tol):
    from cgs import nets as N
    arch = N.get_arch(name)
    w = onets.init_weights(arch, seed=21)
    for k in list(w):                       # O(1) signal through the fc stack
        if k.endswith("/Matrix") or k.endswith("/w"):
            w[k] = (w[k] * 4).astype(np.float32)
    z = np.random.RandomState(0).uniform(-1, 1, (9, arch["z_dim"])).astype(np.float32)   # nsgan/GAN.py:216
    ref = onets.input_to_feature(z, arch, w).numpy()
    head = N.ProposalHead(arch, w, cuda_device, math=math)
    got = head(z).cpu().numpy()
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(name, math, "proposal head rel-L2 %.2e" % rel)
    assert got.shape == ref.shape and rel <= tol


def _stream(seed, dim=3):
    rng = np.random.RandomState(seed)

    def propose(n):
        return rng.randn(n, dim).astype(np.float32), rng.beta(2, 5, size=(n, 1)).astype(np.float32)
    return propose


@pytest.mark.parametrize("kind,T,eval_size,batch", [("rejection", 0, 300, 64), ("hastings", 3, 300, 64),
                                                    ("hastings", 20, 128, 32), ("synthetic", 5, 257, 0)])
def test_fill_up_matches_oracle_loop(cgs_lib, cuda_device, kind, T, eval_size, batch):
    """cgs.fillup vs the oracle restatement of nsgan/GAN.py:311-433 / synthetic/main.py:149-214 under the SAME proposals
    and the SAME uniforms (both sides draw from the global numpy RNG exactly like the reference): accepted rows in
    order, cnt, cnt_propose, back-fill count."""
    from cgs import fillup as P
    from oracle import fillup_np as F
    from sampling.idpsampler import IndependenceSampler
    from sampling.rejector import Rejector
    res = {}
    for side in ("oracle", "product"):
        propose = _stream(3)
        base_x, base_s = propose(eval_size)
        np.random.seed(11)
        if kind == "rejection":
            smp = F.OracleRejector() if side == "oracle" else Rejector()
            smp.set_score_max(np.float32(0.9))
            sampling = lambda x, s, smp=smp: smp.sampling(x, s, shift_percent=100.0)
            guard = "running"
        else:
            smp = F.OracleIndependenceSampler(T=T) if side == "oracle" else IndependenceSampler(T=T)
            smp.set_score_curr(np.float32(0.3))
            sampling = smp.sampling
            guard = "batch"
        if side == "oracle":
            r = F.fill_up_synthetic(base_x, base_s, propose, sampling) if kind == "synthetic" else \
                F.fill_up_nsgan(base_x, base_s, propose, sampling, eval_size, batch, store_guard=guard)
            res[side] = (r["samples"].astype(np.float32), r["cnt"], r["cnt_propose"], r["n_backfilled"], r["n_batches"])
        else:
            def propose_dev(n):
                x, sc = propose(n)
                return torch.from_numpy(x).to(cuda_device), torch.from_numpy(sc).to(cuda_device)
            bx, bs = torch.from_numpy(base_x).to(cuda_device), torch.from_numpy(base_s).to(cuda_device)
            r = P.fill_up_synthetic(bx, bs, propose_dev, sampling) if kind == "synthetic" else \
                P.fill_up(bx, bs, propose_dev, sampling, eval_size, batch, store_guard=guard)
            assert r.samples.is_cuda
            res[side] = (r.samples.cpu().numpy(), r.cnt, r.cnt_propose, r.n_backfilled, r.n_batches)
    o, g = res["oracle"], res["product"]
    assert o[1:] == g[1:], (o[1:], g[1:])
    assert np.array_equal(o[0], g[0])
    if kind == "hastings" and T == 20:
        assert g[3] > 0                       # efficiency <= 1/21 < MIN_EFFICIENCY: un-filtered back-fill (App. C8)


def test_fill_up_driver(cgs_lib, cuda_device):
    """z -> head -> refine -> MH accept, filled up to eval_size like nsgan/GAN.py:384-433, all on the device."""
    from cgs import nets as N, synthetic as S
    from cgs.fillup import fill_up
    from sampling.collaborator import Refiner
    from sampling.idpsampler import IndependenceSampler
    arch = N.get_arch("mnist")
    w = S.init_weights(arch, gain=3.0)
    spec = N.NetSpec(arch, w, cuda_device)
    head = N.ProposalHead(arch, w, cuda_device)
    refiner = Refiner(3, 0.1)
    refiner.set_env(N.discriminator_spec(spec), N.feature_to_data_spec(spec), N.loss_refine)
    g = torch.Generator(device="cpu").manual_seed(0)

    def propose(n):
        x = refiner.build_refiner(head(torch.rand(n, arch["z_dim"], generator=g) * 2 - 1))
        return x, torch.sigmoid(refiner.optimal_logit)

    mh = IndependenceSampler(T=3, rng="philox", seed=1)
    mh.set_score_curr(np.float32(0.5))
    r = fill_up(*propose(96), propose, mh.sampling, eval_size=96, batch_size=32)
    assert r.samples.shape == (96, 28, 28, 1) and r.samples.is_cuda and torch.isfinite(r.samples).all()
    assert 0.0 < r.efficiency <= 0.25 + 0.05       # T=3 -> about one emission per 4 proposals (+ back-filled rows)
    # with T=20 efficiency < MIN_EFFICIENCY: the driver back-fills un-filtered batches like the reference (App. C8)
    mh2 = IndependenceSampler(T=20, rng="philox", seed=1)
    mh2.set_score_curr(np.float32(0.5))
    r2 = fill_up(*propose(64), propose, mh2.sampling, eval_size=64, batch_size=32)
    assert r2.samples.shape[0] == 64 and r2.n_backfilled > 0

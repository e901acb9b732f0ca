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
        return head(torch.rand(n, arch["z_dim"], generator=g) * 2 - 1)

    def refine(h):
        return refiner.build_refiner(h)

    def score(x):
        return torch.sigmoid(refiner.optimal_logit)

    mh = IndependenceSampler(T=3, rng="philox", seed=1)
    mh.set_score_curr(np.float32(0.5))
    out, eff, backfilled = fill_up(propose, score, mh, eval_size=96, batch_size=32, refine=refine)
    assert out.shape == (96, 28, 28, 1) and out.is_cuda and torch.isfinite(out).all()
    assert 0.0 < eff <= 0.25 + 1e-6          # T=3 -> at most one emission per 4 proposals
    # with T=20 efficiency < MIN_EFFICIENCY: the driver back-fills un-filtered batches like the reference (App. C8)
    mh2 = IndependenceSampler(T=20, rng="philox", seed=1)
    mh2.set_score_curr(np.float32(0.5))
    out2, eff2, backfilled2 = fill_up(propose, score, mh2, eval_size=64, batch_size=32, refine=refine)
    assert out2.shape[0] == 64 and backfilled2 > 0

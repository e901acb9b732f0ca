"""GPU kernels vs the committed golden vectors (reference-generated for the sampling stage, oracle-generated for the
conv path) and size-independent properties at the BASELINE batch sizes."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import nets as onets

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "sampling_ref.npz"))


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_drs_vs_reference_vectors(cgs_lib, cuda_device, gold, tag):
    from sampling.rejector import Rejector
    rej = Rejector()
    rej.set_score_max(gold["drs_%s_smax" % tag])
    assert rej.D_tilde_M == float(gold["drs_%s_M0" % tag])
    for call, sp in enumerate((100.0, 60.0, None)):
        sig, u = gold["drs_%s_%d_sig" % (tag, call)], gold["drs_%s_%d_u" % (tag, call)]
        rows = np.arange(len(sig), dtype=np.float32).reshape(-1, 1)
        good = rej.sampling(rows, sig, shift_percent=sp, uniforms=u)
        assert np.array_equal(good[:, 0].astype(np.int64), gold["drs_%s_%d_accepted_rows" % (tag, call)])
        assert abs(rej.D_tilde_M - float(gold["drs_%s_%d_M" % (tag, call)])) <= 4e-16 * abs(rej.D_tilde_M)


@pytest.mark.parametrize("tag,dt", [("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("T,B", [(0, 0), (5, 3), (20, 0)])
def test_mh_vs_reference_vectors(cgs_lib, cuda_device, gold, tag, dt, T, B):
    from sampling.idpsampler import IndependenceSampler
    key = "mh_%s_T%d_B%d" % (tag, T, B)
    smp = IndependenceSampler(T=T, B=B)
    smp.set_score_curr(dt(gold[key + "_d0"]))
    for call in range(3):
        sig, u = gold["%s_%d_sig" % (key, call)], gold["%s_%d_u" % (key, call)]
        rows = np.arange(len(sig), dtype=np.float32).reshape(-1, 1)
        good = smp.sampling(rows, sig, uniforms=u)
        assert np.array_equal(good.reshape(-1).astype(np.int64), gold["%s_%d_emit" % (key, call)])
        assert float(smp.d_curr) == float(gold["%s_%d_d" % (key, call)])
        assert smp.cnt_chain == int(gold["%s_%d_cnt" % (key, call)])


@pytest.mark.parametrize("method", ["sgd", "momentum", "ladam"])
def test_policy_vs_reference_vectors(cgs_lib, cuda_device, gold, method):
    from sampling.policy import PolicyAdaptive
    pol = PolicyAdaptive(0.1, method)
    theta = torch.from_numpy(gold["policy_%s_theta0" % method].copy()).to(cuda_device)
    for it in range(5):
        pol.apply_gradient(theta, torch.from_numpy(gold["policy_%s_%d_grad" % (method, it)]).to(cuda_device),
                           torch.from_numpy(gold["policy_%s_%d_loss" % (method, it)]).to(cuda_device))
        assert np.array_equal(theta.cpu().numpy(), gold["policy_%s_%d_theta" % (method, it)])


@pytest.mark.parametrize("K", [10, 50])
def test_refine2d_vs_reference_vectors(cgs_lib, cuda_device, gold, K):
    from sampling.refiner_cpu import MlpSpec, Refiner
    ws = onets.init_mlp2d(64, 6, seed=2019, gain=1.5)
    key = "r2d_K%d" % K

    class Data:
        def next_batch(self, n):
            return gold[key + "_real"][:n]

    ref = Refiner(types.SimpleNamespace(rollout_steps=K, rollout_rate=0.1, rollout_method="ladam"))
    ref.set_env(MlpSpec(ws, cuda_device), None, Data())
    out = ref.manipulate_sample(gold[key + "_x0"], "deterministic")
    err = np.abs(out - gold[key + "_out"]).max(axis=1)
    print("2-D refine K=%d vs reference output: max-abs %.3e" % (K, err.max()))
    assert (err <= 1e-4 * 10.0).mean() >= 0.999          # BASELINE.md §5, scale = 10


@pytest.mark.parametrize("name", ["mnist", "dcgan32_l2", "dcgan64_l1"])
def test_graph_refiner_vs_oracle_fixture(cgs_lib, cuda_device, name):
    from cgs import nets as N
    from sampling.collaborator import Refiner
    g = np.load(os.path.join(GOLD, "graph_refiner.npz"))
    B, K, gain = g[name + "_cfg"]
    arch = N.get_arch(name)
    w = onets.scale_weights_for_signal(arch, onets.init_weights(arch, seed=2019), float(gain))
    # stated tolerances: FP32 SIMT mode 2e-4 max-abs (summation order only); TF32 tensor mode 6e-2 max-abs and
    # 1e-2 relative L2 on images after K steps (10-bit mantissa operands; ReLU masks of near-zero units may flip)
    for math, tol in (("fp32", 2e-4), ("tf32", 6e-2)):
        spec = N.NetSpec(arch, w, cuda_device, math=math)
        ref = Refiner(int(K), 0.1)
        ref.set_env(N.discriminator_spec(spec), N.feature_to_data_spec(spec), N.loss_refine)
        out = ref.build_refiner(torch.from_numpy(g[name + "_h0"]).to(cuda_device))
        e_img = np.abs(out.cpu().numpy() - g[name + "_refined"]).max()
        e_log = np.abs(ref.optimal_logit.cpu().numpy() - g[name + "_optimal_logit"]).max()
        print(name, math, "image max-abs %.2e  optimal_logit max-abs %.2e" % (e_img, e_log))
        rel = np.linalg.norm(out.cpu().numpy() - g[name + "_refined"]) / np.linalg.norm(g[name + "_refined"])
        assert e_img <= tol and e_log <= tol * max(1.0, np.abs(g[name + "_optimal_logit"]).max())
        assert rel <= (1e-5 if math == "fp32" else 1e-2), rel
        if math == "fp32":
            assert np.array_equal(ref.optimal_step.cpu().numpy(), g[name + "_optimal_step"])


@pytest.mark.parametrize("math", ["tf32", "fp32"])
def test_full_size_properties_mnist_batch_1024(cgs_lib, cuda_device, math):
    """BASELINE config 2 size (B=1024): shard invariance (bit-exact), monotone best-of-K, K=0 identity."""
    from cgs import nets as N, synthetic as S
    from sampling.collaborator import Refiner
    arch = N.get_arch("mnist")
    spec = N.NetSpec(arch, S.init_weights(arch, gain=3.0), cuda_device, math=math)
    B, K = (1024, 6) if math == "tf32" else (256, 2)
    h0 = torch.from_numpy(S.proposal_features(arch, B, seed=5)).to(cuda_device)
    ref = Refiner(K, 0.1)
    ref.set_env(N.discriminator_spec(spec), N.feature_to_data_spec(spec), N.loss_refine)
    full = ref.build_refiner(h0).clone()
    full_logit, full_step, dflt = ref.optimal_logit.clone(), ref.optimal_step.clone(), ref.default_logit.clone()
    assert bool((full_logit >= dflt).all())                       # best-of-K never loses to the proposal
    assert bool(((full_step >= 1) & (full_step <= K)).all())
    assert bool(torch.isfinite(full).all()) and float(full.abs().max()) <= 1.0
    # per-sample independence: any split of the batch gives bit-identical rows (what makes multi-GPU sharding exact)
    cut = 384 if B == 1024 else 100
    a = ref.build_refiner(h0[:cut]).clone()
    la = ref.optimal_logit.clone()
    b = ref.build_refiner(h0[cut:]).clone()
    assert torch.equal(torch.cat([a, b]), full)
    assert torch.equal(torch.cat([la, ref.optimal_logit]), full_logit)
    # K = 0: the refined batch is the proposals' own image and the default logit
    r0 = Refiner(0, 0.1)
    r0.set_env(N.discriminator_spec(spec), N.feature_to_data_spec(spec), N.loss_refine)
    img0 = r0.build_refiner(h0[:64])
    img_direct, logit_direct = ref.feature_to_image(h0[:64])
    assert torch.equal(img0, img_direct) and torch.equal(r0.optimal_logit, logit_direct)


def test_accept_properties_at_eval_size(cgs_lib, cuda_device):
    """nsgan eval size (49 984 rows): ordered compaction, thinning period, determinism of the Philox stream."""
    from sampling.idpsampler import IndependenceSampler
    from sampling.rejector import Rejector
    n = 49984
    g = torch.Generator(device="cpu").manual_seed(1)
    sig = torch.distributions.Beta(2.0, 5.0).sample((n, 1)).float()
    rows = torch.arange(n, dtype=torch.float32).reshape(n, 1).to(cuda_device)
    rej = Rejector(rng="philox", seed=7)
    out1 = rej.sampling(rows, sig.to(cuda_device), shift_percent=100.0)
    idx = rej.last_indices.cpu().numpy()
    assert np.all(np.diff(idx) > 0) and np.array_equal(out1[:, 0].cpu().numpy(), idx.astype(np.float32))
    rej2 = Rejector(rng="philox", seed=7)
    rej2.D_tilde_M = 0.0
    out2 = rej2.sampling(rows, sig.to(cuda_device), shift_percent=100.0)
    assert torch.equal(out1, out2)
    mh = IndependenceSampler(T=20, rng="philox", seed=3)
    mh.set_score_curr(np.float32(0.3))
    out = mh.sampling(rows, sig.to(cuda_device))
    src = mh.last_emit_src.cpu().numpy()
    acc = mh.last_accepted.cpu().numpy().astype(bool)
    assert np.all(np.diff(src) >= 0) and acc[src].all()
    first = np.nonzero(acc)[0][0] + 20                           # cnt_chain starts at 1: first emission T rows later
    assert len(src) == (n - 1 - first) // 21 + 1
    assert out.shape == (len(src), 1)

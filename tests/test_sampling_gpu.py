"""GPU parity: policy / DRS / MH kernels through the drop-in classes vs the pinned numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import sampling_np as snp

import philox_ref

pytestmark = pytest.mark.gpu


def _scores(rng, n, dtype):
    return rng.beta(2, 5, size=(n, 1)).astype(dtype)


@pytest.mark.parametrize("method", ["sgd", "momentum", "ladam"])
def test_policy_bit_exact(cgs_lib, cuda_device, method):
    from sampling.policy import PolicyAdaptive
    rng = np.random.RandomState(3)
    n = 1000
    theta = (rng.randn(n, 2) * 3).astype(np.float32)
    ref_theta = theta.copy()
    pol = PolicyAdaptive(0.1, method)
    state = snp.policy_new_state()
    th_dev = torch.from_numpy(theta.copy()).to(cuda_device)
    for it in range(6):
        g = (rng.randn(n, 2) * 10 ** rng.uniform(-6, 0, size=(n, 1))).astype(np.float32)
        loss = (rng.rand(n) - 0.5).astype(np.float32)
        snp.policy_step(method, ref_theta, g, state, 0.1, loss)
        out = pol.apply_gradient(th_dev, torch.from_numpy(g).to(cuda_device), torch.from_numpy(loss).to(cuda_device))
        assert out.data_ptr() == th_dev.data_ptr()
        assert np.array_equal(th_dev.cpu().numpy(), ref_theta), (method, it)
    pol.reset_moving_average()
    assert pol.momentum is None and pol.mean_square is None and pol.loss is None


def test_policy_numpy_inplace_and_errors(cgs_lib, cuda_device):
    from sampling.policy import PolicyAdaptive
    pol = PolicyAdaptive(0.1, "momentum")
    th = np.ones((5, 2), np.float32)
    r = pol.apply_gradient(th, np.ones((5, 2), np.float32))
    assert r is th and np.allclose(th, 0.9)
    with pytest.raises(NotImplementedError):
        PolicyAdaptive(0.1, "adamw").apply_gradient(th, th)
    with pytest.raises(TypeError):
        PolicyAdaptive(0.1, "ladam").apply_gradient(th, th)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 31, 1000, 50000])
def test_drs_decisions_bit_exact(cgs_lib, cuda_device, dtype, n):
    from sampling.rejector import Rejector
    rng = np.random.RandomState(n)
    rej = Rejector()
    smax = np.amax(_scores(rng, 200, dtype))
    rej.set_score_max(smax)
    m_ref = snp.drs_score_max(smax)
    assert rej.D_tilde_M == float(m_ref)
    samples = rng.randn(n, 3).astype(np.float32)
    for call, sp in enumerate([100.0, 60.0, None, 100.0, 0.0, 37.5]):
        sig = _scores(rng, n, dtype)
        if call == 3:
            sig[n // 2] = 0.999999     # fake score above the running max (SURVEY App. A9 collapse)
        np.random.seed(100 + call)
        good = rej.sampling(samples, sig, shift_percent=sp)
        np.random.seed(100 + call)
        u = np.random.rand(n)
        acc, m_ref = snp.drs_accept(sig, u, m_ref, shift_percent=sp)
        assert np.array_equal(rej.last_accept.cpu().numpy().astype(bool), acc), (dtype, n, sp)
        assert np.array_equal(good, samples[acc])
        assert np.array_equal(rej.last_indices.cpu().numpy(), np.nonzero(acc)[0])
        assert abs(rej.D_tilde_M - float(m_ref)) <= 4e-16 * abs(float(m_ref))


def test_drs_edge_cases(cgs_lib, cuda_device):
    from sampling.rejector import Rejector
    rej = Rejector()
    with pytest.raises(NotImplementedError):
        rej.sampling(np.zeros((4, 2), np.float32), np.full((4, 1), .5, np.float32), ranking=[1])
    # scores at the clip bounds
    sig = np.array([[0.0], [1.0], [1e-14], [1 - 1e-14], [0.5]], dtype=np.float64)
    u = np.array([0.5, 0.5, 0.5, 0.5, 0.5])
    good = rej.sampling(np.arange(5, dtype=np.float32).reshape(5, 1), sig, shift_percent=100.0, uniforms=u)
    acc, _ = snp.drs_accept(sig, u, 0.0, shift_percent=100.0)
    assert np.array_equal(good[:, 0], np.arange(5, dtype=np.float32)[acc])
    # empty accept set keeps the sample shape
    rej2 = Rejector()
    rej2.set_score_max(np.float32(0.99))
    out = rej2.sampling(np.zeros((8, 2, 2), np.float32), np.full((8, 1), 1e-3, np.float32), shift_percent=None,
                        uniforms=np.full(8, 0.999999))
    assert out.shape == (0, 2, 2)
    # torch in -> torch out, on device
    t = rej2.sampling(torch.zeros(8, 2, device=cuda_device), torch.full((8, 1), .3, device=cuda_device))
    assert isinstance(t, torch.Tensor) and t.is_cuda


def test_drs_philox_stream(cgs_lib, cuda_device):
    from sampling.rejector import Rejector
    rng = np.random.RandomState(5)
    n = 4097
    sig = _scores(rng, n, np.float32)
    rej = Rejector(rng="philox", seed=0x1234567890ABCDEF)
    rej.offset = 77
    rej.sampling(np.zeros((n, 1), np.float32), sig, shift_percent=100.0)
    u = philox_ref.philox_uniform_f64(0x1234567890ABCDEF, np.arange(77, 77 + n, dtype=np.uint64))
    acc, _ = snp.drs_accept(sig, u, 0.0, shift_percent=100.0)
    assert np.array_equal(rej.last_accept.cpu().numpy().astype(bool), acc)
    assert rej.offset == 77 + n


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("T,B", [(0, 0), (1, 0), (5, 3), (20, 0), (20, 7)])
def test_mh_chain_bit_exact(cgs_lib, cuda_device, dtype, T, B):
    from sampling.idpsampler import IndependenceSampler
    rng = np.random.RandomState(T * 10 + B)
    smp = IndependenceSampler(T=T, B=B)
    real = _scores(rng, 100, dtype)
    smp.set_score_curr(np.mean(real))
    d_ref, c_ref = np.mean(real), 1
    for call, n in enumerate([3000, 1, 777, 2049]):
        sig = _scores(rng, n, dtype)
        samples = rng.randn(n, 4).astype(np.float32)
        np.random.seed(call)
        good = smp.sampling(samples, sig)
        np.random.seed(call)
        u = np.random.rand(n)
        emit, d_ref, c_ref, acc = snp.mh_chain(sig, u, d_ref, c_ref, T, B)
        assert np.array_equal(smp.last_accepted.cpu().numpy().astype(bool), acc), (dtype, T, B, call)
        assert np.array_equal(smp.last_emit_src.cpu().numpy(), emit), (dtype, T, B, call)
        exp = samples[emit] if len(emit) else np.zeros((0,), np.float32)
        assert good.dtype == np.float32 and np.array_equal(good, exp)
        assert smp.cnt_chain == c_ref
        assert float(smp.d_curr) == float(np.squeeze(d_ref))


def test_mh_edge_cases(cgs_lib, cuda_device):
    from sampling.idpsampler import IndependenceSampler
    rng = np.random.RandomState(0)
    # d_curr None: first row moves unconditionally and draws no uniform
    smp = IndependenceSampler(T=2)
    sig = _scores(rng, 500, np.float32)
    u = rng.rand(499)
    smp.sampling(np.zeros((500, 1), np.float32), sig, uniforms=u)
    emit, d, c, acc = snp.mh_chain(sig, np.concatenate([[0.0], u]), None, 1, 2, 0)
    # the oracle consumes u[i] at row i; with d_curr None row 0 draws nothing, so shift by one
    emit2, d2, c2, acc2 = snp.mh_chain(sig[1:], u, sig[0, 0], 1, 2, 0)
    assert acc[0] and np.array_equal(smp.last_accepted.cpu().numpy().astype(bool)[1:], acc2)
    # pathological scores: 0 and 1 (division by zero -> inf / nan handled like python's min)
    sig = np.array([0.0, 1.0, 0.5, 1.0, 0.0, 0.3], dtype=np.float32).reshape(-1, 1)
    u = np.array([0.1, 0.9, 0.5, 0.2, 0.7, 0.4])
    for d0 in (np.float32(0.5), np.float64(0.0), np.float32(1.0)):
        s = IndependenceSampler(T=0)
        s.set_score_curr(d0)
        s.sampling(np.zeros((6, 1), np.float32), sig, uniforms=u)
        emit, d, c, acc = snp.mh_chain(sig, u, d0, 1, 0, 0)
        assert np.array_equal(s.last_accepted.cpu().numpy().astype(bool), acc), d0
    with pytest.raises(AssertionError):
        IndependenceSampler().sampling(np.zeros((2, 1), np.float32), np.array([[0.5], [1.5]], np.float32))
    # outlier with a very high score forces a long forward scan
    sig = _scores(rng, 20000, np.float32)
    sig[10] = 0.9999
    u = rng.rand(20000)
    s = IndependenceSampler(T=20)
    s.set_score_curr(np.float32(0.2))
    s.sampling(np.zeros((20000, 1), np.float32), sig, uniforms=u)
    emit, d, c, acc = snp.mh_chain(sig, u, np.float32(0.2), 1, 20, 0)
    assert np.array_equal(s.last_accepted.cpu().numpy().astype(bool), acc)
    assert np.array_equal(s.last_emit_src.cpu().numpy(), emit)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mh_sticky_chain(cgs_lib, cuda_device, dtype):
    """Scores spread over many decades of 1-d: most states reject for hundreds of rows, so almost every row of
    the successor table is finished by the warp-wide scan (sampling_kernels.cu mh_next_kernel)."""
    from sampling.idpsampler import IndependenceSampler
    rng = np.random.RandomState(5)
    n = 40000
    hi = 6 if dtype == np.float32 else 12
    sig = (1.0 - 10.0 ** (-rng.uniform(0.5, hi, size=(n, 1)))).astype(dtype)
    sig = np.clip(sig, 0.0, np.nextafter(dtype(1.0), dtype(0.0)))
    for d0, u in ((dtype(0.5), rng.rand(n)), (None, rng.rand(n - 1))):
        s = IndependenceSampler(T=3, B=2)
        if d0 is not None:
            s.set_score_curr(d0)
            emit, d, c, acc = snp.mh_chain(sig, u, d0, 1, 3, 2)
        else:
            emit, d, c, acc = snp.mh_chain(sig, np.concatenate([[0.0], u]), None, 1, 3, 2)
        s.sampling(np.zeros((n, 1), np.float32), sig, uniforms=u)
        got = s.last_accepted.cpu().numpy().astype(bool)
        assert np.array_equal(got, acc)
        assert 1 <= acc.sum() < n // 4                     # mean gap well beyond the per-thread scan for many rows
        assert np.array_equal(s.last_emit_src.cpu().numpy(), emit)
        assert float(s.d_curr) == float(np.squeeze(d))
    # counter-based uniforms through the same path
    s = IndependenceSampler(T=0, rng="philox", seed=99)
    s.set_score_curr(dtype(0.9))
    s.sampling(np.zeros((n, 1), np.float32), sig)
    u = philox_ref.philox_uniform_f64(99, np.arange(0, n, dtype=np.uint64))
    emit, d, c, acc = snp.mh_chain(sig, u, dtype(0.9), 1, 0, 0)
    assert np.array_equal(s.last_accepted.cpu().numpy().astype(bool), acc)

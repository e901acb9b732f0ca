"""Oracle vs the reference's own modules executed in place (build container only: /root/reference must exist)."""
import types

import numpy as np
import pytest

from oracle import nets as onets
from oracle import ref_shims
from oracle import sampling_np as snp

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shims.load_reference_sampling()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_drs_bit_exact(ref, dt):
    rng = np.random.RandomState(1)
    R = ref.rejector.Rejector()
    R.set_score_max(np.amax(rng.beta(2, 5, size=200).astype(dt)))
    m = R.D_tilde_M
    samples = rng.randn(1000, 2).astype(np.float32)
    for call, sp in enumerate((100.0, 60.0, None, 100.0, 12.5)):
        sig = rng.beta(2, 5, size=(1000, 1)).astype(dt)
        if call == 3:
            sig[17] = 1.0 - 1e-9
        np.random.seed(call)
        good = R.sampling(samples, sig, shift_percent=sp)
        np.random.seed(call)
        acc, m = snp.drs_accept(sig, np.random.rand(1000), m, shift_percent=sp)
        assert np.array_equal(samples[acc], good)
        assert float(m) == float(R.D_tilde_M)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("T,B", [(0, 0), (1, 0), (5, 2), (20, 0), (20, 5)])
def test_mh_bit_exact_with_carried_state(ref, dt, T, B):
    rng = np.random.RandomState(T + 7 * B)
    S = ref.idpsampler.IndependenceSampler(T=T, B=B)
    S.set_score_curr(np.mean(rng.beta(2, 5, size=64).astype(dt)))
    d, cnt = S.d_curr, S.cnt_chain
    for call in range(4):
        n = (900, 1, 333, 64)[call]
        sig = rng.beta(2, 5, size=(n, 1)).astype(dt)
        samples = rng.randn(n, 3).astype(np.float32)
        np.random.seed(call)
        good = S.sampling(samples, sig)
        np.random.seed(call)
        emit, d, cnt, _ = snp.mh_chain(sig, np.random.rand(n), d, cnt, T, B)
        exp = samples[emit] if len(emit) else np.zeros((0,), np.float32)
        assert np.array_equal(exp.astype(np.float32), good)
        assert cnt == S.cnt_chain and float(np.squeeze(d)) == float(np.squeeze(S.d_curr))


def test_mh_uniform_stream_equals_rand(ref):
    np.random.seed(3)
    a = np.array([np.random.uniform(0, 1) for _ in range(50)])
    np.random.seed(3)
    assert np.array_equal(a, np.random.rand(50))


@pytest.mark.parametrize("method", ["sgd", "momentum", "ladam"])
def test_policy_bit_exact(ref, method):
    rng = np.random.RandomState(5)
    P = ref.policy.PolicyAdaptive(0.1, method)
    a = (rng.randn(100, 2)).astype(np.float32)
    b = a.copy()
    st = snp.policy_new_state()
    for _ in range(4):
        g = rng.randn(100, 2).astype(np.float32) * 1e-3
        l = (rng.rand(100) - .5).astype(np.float32)
        P.apply_gradient(a, g, l)
        snp.policy_step(method, b, g, st, 0.1, l)
        assert np.array_equal(a, b)
    with pytest.raises(NotImplementedError):
        ref.policy.PolicyAdaptive(0.1, "nope").apply_gradient(a, a)


@pytest.mark.parametrize("mode", ["deterministic", "probabilistic"])
def test_refiner_cpu_bit_exact(ref, mode):
    ws = onets.init_mlp2d(64, 6, seed=4, gain=1.5)
    sess = ref_shims.FakeSession(ws)
    data = ref.Datasets.ToyDataset("Imbal-8Gaussians", scale=10, ratio=0.9)
    K, n = 12, 400
    Rf = ref.refiner_cpu.Refiner(types.SimpleNamespace(rollout_steps=K, rollout_rate=0.1, rollout_method="ladam"))
    Rf.set_env(ref_shims.FakeGan, sess, data)
    x0 = (np.random.RandomState(0).randn(n, 2) * 4).astype(np.float32)
    np.random.seed(9)
    out = Rf.manipulate_sample(x0, mode)
    assert sess.calls == K + 2                       # K+2 host<->runtime crossings per call
    np.random.seed(9)
    real = data.next_batch(n)
    rs, _ = onets.mlp2d_sigmoid_saliency(real.astype(np.float32), ws)
    idx = np.random.randint(K + 1, size=n) if mode == "probabilistic" else None
    o = snp.refine_2d(x0, lambda x: onets.mlp2d_sigmoid_saliency(x, ws), np.mean(rs), K, 0.1, "ladam", idx)
    exp = o["probabilistic"] if mode == "probabilistic" else o["optimal_batch"]
    assert out.dtype == exp.dtype and np.array_equal(out, exp)

"""The K=50 parity metrics themselves (tests/parity_metrics.py) checked on the CPU against textbook definitions."""
import numpy as np
import scipy.linalg

from parity_metrics import frechet_distance, jaccard, k50_metrics


def test_frechet_distance_matches_sqrtm_definition():
    rng = np.random.RandomState(0)
    x1 = rng.randn(200, 12) @ rng.randn(12, 12)
    x2 = rng.randn(150, 12) @ rng.randn(12, 12) + 0.3
    c1, c2 = np.cov(x1, rowvar=False), np.cov(x2, rowvar=False)
    want = ((x1.mean(0) - x2.mean(0)) ** 2).sum() + np.trace(c1 + c2 - 2 * scipy.linalg.sqrtm(c1 @ c2).real)
    assert abs(frechet_distance(x1, x2) - want) <= 1e-8 * max(1.0, abs(want))
    assert abs(frechet_distance(x1, x1)) <= 1e-9 * np.trace(c1)


def test_frechet_distance_rank_deficient_sets():
    """d >> n (the DCGAN-64 case: 8192 features, 16 samples): still exact, zero for identical sets."""
    rng = np.random.RandomState(1)
    x = rng.randn(16, 4096)
    assert abs(frechet_distance(x, x)) <= 1e-8 * float((x * x).sum())
    assert frechet_distance(x, x + 0.5) > 0.24 * 4096


def test_jaccard_and_metric_dict():
    assert jaccard([1, 2, 3], [2, 3, 4]) == 0.5 and jaccard([], []) == 1.0
    from oracle import graph_refiner as gr, nets as onets
    import torch
    arch = onets.get_arch("mnist")
    w = onets.scale_weights_for_signal(arch, onets.init_weights(arch, seed=9), 3.0)
    h0 = torch.relu(torch.randn(6, *arch["feature_shape"], generator=torch.Generator().manual_seed(2)))
    o = gr.build_refiner(h0, arch, w, 2, 0.1)
    d = {k: o[k].numpy() for k in ("refined", "optimal_logit", "optimal_step", "default_logit")}
    m = k50_metrics(d, d, arch, w)
    assert m["img_rel_l2"] == 0 and m["optimal_step_agree"] == 1.0 and m["mh_T0_emit_jaccard"] == 1.0
    assert abs(m["frechet_d_feature"]) <= 1e-6 * max(m["d_feature_trace_cov"], 1.0)

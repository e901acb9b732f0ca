"""K=50 parity of the image path on the BENCHMARK configurations.

The reference's default is rollout_steps = 50 (nsgan/main.py:28-48; loop at sampling/collaborator.py:63-83) with the
momentum policy (policy.py:31-37), which integrates per-step gradient error: these tests run the full 50 steps and
compare the CUDA path with the CPU oracle (oracle/graph_refiner.py) on the same weights and proposals, in both math
modes, and the TF32 tensor path with the exact-FP32 SIMT path on the GPU at the benchmark batch.

Stated tolerances at K=50 (measured values in profiles/round2_k50_parity.json):
  TF32 tensor path   image rel-L2 <= 1e-2, max-abs <= 0.4 (images in [-1,1]); |d optimal_logit| <= 2 % of the mean
                     logit gain (>= 0.05); optimal_step agreement >= 95 %; sign(logit - threshold) agreement >= 99 %;
                     MH(T=20) emitted-set Jaccard >= 0.9 under the same uniforms; Frechet distance in D's penultimate
                     feature space <= 0.5 (BASELINE.md §5's FID substitute)
  FP32 SIMT path     image rel-L2 <= 5e-4; |d optimal_logit| <= 2e-3; optimal_step agreement >= 99 %; Frechet <= 1e-3
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import graph_refiner as gr
from oracle import nets as onets
from parity_metrics import k50_metrics

pytestmark = pytest.mark.gpu

K = 50
GAIN = {"mnist": 3.0}
TOL = {
    "tf32": dict(img_rel_l2=1e-2, img_max_abs=0.4, logit_frac=0.02, logit_floor=0.05, step=0.95, sign=0.99,
                 jaccard20=0.9, jaccard0=0.97, frechet=0.5),
    "fp32": dict(img_rel_l2=5e-4, img_max_abs=0.05, logit_frac=0.0, logit_floor=2e-3, step=0.99, sign=0.999,
                 jaccard20=0.99, jaccard0=0.99, frechet=1e-3),
}


def _setup(name, B):
    from cgs import nets as N
    arch = N.get_arch(name)
    w = onets.scale_weights_for_signal(arch, onets.init_weights(arch, seed=9), GAIN.get(name, 2.5))
    h0 = torch.relu(torch.randn(B, *arch["feature_shape"], generator=torch.Generator().manual_seed(2)))
    return arch, w, h0


def _gpu(arch, w, h0, math, dev):
    from cgs import nets as N
    from sampling.collaborator import Refiner
    spec = N.NetSpec(arch, w, dev, math=math)
    ref = Refiner(K, 0.1, "momentum")
    ref.set_env(N.discriminator_spec(spec), N.feature_to_data_spec(spec), N.loss_refine)
    x = ref.build_refiner(h0.to(dev), None, "deterministic")
    torch.cuda.synchronize()
    return dict(refined=x.cpu().numpy(), optimal_logit=ref.optimal_logit.cpu().numpy(),
                optimal_step=ref.optimal_step.cpu().numpy(), default_logit=ref.default_logit.cpu().numpy())


def _check(m, math, tag):
    t = TOL[math]
    print(tag, json.dumps(m))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "k50_parity_tests.jsonl"), "a") as f:
            f.write(json.dumps(dict(m, tag=tag, math=math)) + "\n")
    assert m["img_rel_l2"] <= t["img_rel_l2"] and m["img_max_abs"] <= t["img_max_abs"], m
    assert m["optimal_logit_max_abs"] <= max(t["logit_floor"], t["logit_frac"] * abs(m["logit_gain_mean_ref"])), m
    assert m["optimal_step_agree"] >= t["step"], m
    assert m["sign_agree_logit0"] >= t["sign"] and m["sign_agree_median"] >= t["sign"], m
    assert m["mh_T20_emit_jaccard"] >= t["jaccard20"] and m["mh_T0_emit_jaccard"] >= t["jaccard0"], m
    assert m["frechet_d_feature"] <= t["frechet"], m
    assert m["logit_gain_mean_ref"] > 1.0, "the case must actually refine (logit gain over 50 steps)"


@pytest.mark.parametrize("math", ["fp32", "tf32"])
@pytest.mark.parametrize("name,B", [("mnist", 256), ("dcgan32_l1", 16), ("dcgan64_l1", 16)])
def test_k50_matches_cpu_oracle(cgs_lib, cuda_device, name, B, math):
    arch, w, h0 = _setup(name, B)
    o = gr.build_refiner(h0, onets.get_arch(name), w, K, 0.1, method="momentum")
    ref = dict(refined=o["refined"].numpy(), optimal_logit=o["optimal_logit"].numpy(),
               optimal_step=o["optimal_step"].numpy(), default_logit=o["default_logit"].numpy())
    got = _gpu(arch, w, h0, math, cuda_device)
    _check(k50_metrics(ref, got, onets.get_arch(name), w), math, "%s B=%d vs cpu oracle" % (name, B))


@pytest.mark.parametrize("name,B", [("mnist", 1024), ("dcgan64_l1", 256)])
def test_k50_tf32_matches_fp32_simt_at_benchmark_batch(cgs_lib, cuda_device, name, B):
    """The exact-FP32 SIMT path (itself checked against the oracle above) as the reference at a batch the CPU oracle
    would take minutes for."""
    arch, w, h0 = _setup(name, B)
    ref = _gpu(arch, w, h0, "fp32", cuda_device)
    got = _gpu(arch, w, h0, "tf32", cuda_device)
    _check(k50_metrics(ref, got, onets.get_arch(name), w), "tf32", "%s B=%d tf32 vs gpu fp32" % (name, B))

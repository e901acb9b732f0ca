"""K=50 parity report of the image path: CUDA path vs the CPU oracle, and TF32 vs FP32-SIMT on the
GPU at the benchmark batch.  Prints one JSON object per case and writes them all to --out.

    python tools/k50_parity.py --out gpurun_out/k50_parity.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "collaborative-gan-sampling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from cgs import nets as N  # noqa: E402
from oracle import graph_refiner as gr, nets as onets  # noqa: E402
from parity_metrics import k50_metrics  # noqa: E402
from sampling.collaborator import Refiner  # noqa: E402

GAIN = {"mnist": 3.0}


def gpu_refine(arch, w, h0, K, math, dev, method="momentum", rate=0.1):
    spec = N.NetSpec(arch, w, dev, math=math)
    ref = Refiner(K, rate, method)
    ref.set_env(N.discriminator_spec(spec), N.feature_to_data_spec(spec), N.loss_refine)
    x = ref.build_refiner(h0.to(dev), None, "deterministic")
    torch.cuda.synchronize()
    return dict(refined=x.cpu().numpy(), optimal_logit=ref.optimal_logit.cpu().numpy(),
                optimal_step=ref.optimal_step.cpu().numpy(), default_logit=ref.default_logit.cpu().numpy())


def oracle_refine(arch, w, h0, K, method="momentum", rate=0.1):
    o = gr.build_refiner(h0, arch, w, K, rate, method=method)
    return dict(refined=o["refined"].numpy(), optimal_logit=o["optimal_logit"].numpy(),
                optimal_step=o["optimal_step"].numpy(), default_logit=o["default_logit"].numpy())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/k50_parity.json")
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--cases", default="mnist:256,dcgan32_l1:16,dcgan64_l1:16")
    ap.add_argument("--gpu-cases", default="mnist:1024,dcgan64_l1:256")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    rows = []
    for case in [c for c in a.cases.split(",") if c]:
        name, B = case.split(":")
        B = int(B)
        arch = N.get_arch(name)
        w = onets.scale_weights_for_signal(arch, onets.init_weights(arch, seed=9), GAIN.get(name, 2.5))
        h0 = torch.relu(torch.randn(B, *arch["feature_shape"], generator=torch.Generator().manual_seed(2)))
        t0 = time.time()
        ref = oracle_refine(arch, w, h0, a.K)
        t_cpu = time.time() - t0
        for math in ("fp32", "tf32"):
            got = gpu_refine(arch, w, h0, a.K, math, dev)
            m = k50_metrics(ref, got, onets.get_arch(name), w)
            m.update(case=name, K=a.K, math=math, against="cpu_oracle", oracle_seconds=round(t_cpu, 2))
            rows.append(m)
            print(json.dumps(m), flush=True)
    for case in [c for c in a.gpu_cases.split(",") if c]:
        name, B = case.split(":")
        B = int(B)
        arch = N.get_arch(name)
        w = onets.scale_weights_for_signal(arch, onets.init_weights(arch, seed=9), GAIN.get(name, 2.5))
        h0 = torch.relu(torch.randn(B, *arch["feature_shape"], generator=torch.Generator().manual_seed(2)))
        ref = gpu_refine(arch, w, h0, a.K, "fp32", dev)
        got = gpu_refine(arch, w, h0, a.K, "tf32", dev)
        m = k50_metrics(ref, got, onets.get_arch(name), w)
        m.update(case=name, K=a.K, math="tf32", against="gpu_fp32_simt")
        rows.append(m)
        print(json.dumps(m), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()

"""Wall-clock breakdown of one end-to-end step of bench.py's MNIST workload (host proposals in, host results out).

    python tools/e2e_breakdown.py [--arch mnist] [--batch 1024] [--steps 50]

Every stage is bracketed by torch.cuda.synchronize() so the figures are additive; the last line is the same loop
without the extra synchronisation (what bench.py's e2e measures).
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "collaborative-gan-sampling_b200"))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from cgs import nets as N
from cgs import synthetic as S
from sampling.collaborator import Refiner
from sampling.idpsampler import IndependenceSampler


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="mnist")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    arch = N.get_arch(a.arch)
    spec = N.NetSpec(arch, S.init_weights(arch, seed=2019, gain=3.0), dev, math="tf32")
    refiner = Refiner(a.steps, 0.1, "momentum", cuda_graph=True)
    refiner.set_env(N.discriminator_spec(spec), N.feature_to_data_spec(spec), N.loss_refine)
    mh = IndependenceSampler(T=20, rng="philox", seed=2019)
    mh.set_score_curr(np.float32(0.5))
    h0_host = torch.from_numpy(S.proposal_features(arch, a.batch, seed=1000)).pin_memory()
    out_host = torch.empty((a.batch,) + tuple(arch["image_shape"]), dtype=torch.float32).pin_memory()

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    stages = {}

    def mark(name, t0):
        t1 = sync()
        stages.setdefault(name, []).append((t1 - t0) * 1e3)
        return t1

    for rep in range(a.reps + 2):
        t = sync()
        h0 = h0_host.to(dev, non_blocking=True)
        t = mark("h2d proposals", t)
        x = refiner.build_refiner(h0, None, "deterministic")
        t = mark("build_refiner (device input)", t)
        sig = torch.sigmoid(refiner.optimal_logit)
        emit = mh.select(sig)
        t = mark("mh.select", t)
        acc = mh.gather(x)
        t = mark("mh.gather", t)
        out_host.copy_(x, non_blocking=True)
        t = mark("d2h refined", t)
        acc_host = acc.cpu()
        t = mark("d2h accepted (pageable)", t)
        stats = (float(acc.shape[0]), float(sig.sum()), float(sig.max()))
        t = mark("stats", t)
    for k, v in stages.items():
        print("%-32s %8.3f ms" % (k, float(np.median(v[2:]))))
    print("%-32s %8.3f ms" % ("sum of stages", sum(float(np.median(v[2:])) for v in stages.values())))

    import gc
    gc_log = []

    def gc_cb(phase, info):
        if phase == "start":
            gc_log.append([info["generation"], time.perf_counter()])
        else:
            gc_log[-1][1] = (time.perf_counter() - gc_log[-1][1]) * 1e3
    gc.callbacks.append(gc_cb)

    def loop(src, label, d2h=True, do_sync=True):
        per = []
        dev_ms = []
        sync()
        for _ in range(a.reps + 3):
            t0 = sync() if do_sync else time.perf_counter()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            x = refiner.build_refiner(src, None, "deterministic")
            ev1.record()
            sig = torch.sigmoid(refiner.optimal_logit)
            mh.select(sig)
            acc = mh.gather(x)
            dev_ms.append((ev0, ev1))
            if d2h:
                out_host.copy_(x if torch.is_tensor(x) else torch.from_numpy(x), non_blocking=True)
                acc_host = acc.cpu() if torch.is_tensor(acc) else acc
                stats = (float(acc.shape[0]), float(sig.sum()), float(sig.max()))
            per.append(((sync() if do_sync else time.perf_counter()) - t0) * 1e3)
        print("%-32s %s" % (label, " ".join("%.2f" % v for v in per)))
        sync()
        print("%-32s %s" % ("   refine, device events", " ".join("%.2f" % a_.elapsed_time(b_) for a_, b_ in dev_ms)))
        print("%-32s %s" % ("   gc (gen, ms)", " ".join("g%d:%.1f" % (g_, t_) for g_, t_ in gc_log)))
        gc_log.clear()

    loop(h0_host.to(dev), "device input, no d2h", d2h=False)
    loop(h0_host.to(dev), "device input, d2h")
    loop(h0_host, "pinned host input, d2h")
    loop(h0_host, "pinned host input, d2h, nosync", do_sync=False)
    loop(h0_host.to(dev), "device input, d2h, nosync", do_sync=False)
    loop(h0_host.numpy(), "numpy input, d2h")
    loop(h0_host, "pinned host input, d2h, nosync", do_sync=False)
    g = [v for v in refiner._graphs.values()][0][0]
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.reps):
        g.replay()
    e.record()
    e.synchronize()
    print("%-32s %8.3f ms" % ("graph replay only", s.elapsed_time(e) / a.reps))


if __name__ == "__main__":
    main()

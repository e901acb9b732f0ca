"""Per-layer GEMM timing (CUDA events, best of N) for a workload; prints a table.  Used for A/B experiments."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "collaborative-gan-sampling_b200"))

import torch  # noqa: E402

from cgs import lib as L, nets as N, synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mnist")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--math", default="tf32")
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--inner", type=int, default=20, help="launches per event pair (hides host launch latency)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = L.load()
    arch = N.get_arch(a.workload)
    spec = N.NetSpec(arch, S.init_weights(arch, gain=2.5), dev, math=a.math)
    chain = [(l, spec.gtail.layer_desc(i)) for i, l in enumerate(arch["gtail"])] + \
            [(l, spec.d.layer_desc(i)) for i, l in enumerate(arch["d"][:-1])]
    B = a.batch
    tot = 0.0
    totf = 0.0
    for layer, desc in chain:
        if a.only and layer["name"] not in a.only.split(","):
            continue
        cin, cout = layer["cin"], layer["cout"]
        if layer["type"] == "fc":
            xs, ys = (B, cin), (B, N.cstride(cout))
        elif layer["type"] == "conv":
            xs = (B, layer["hin"], layer["win"], N.cstride(cin))
            ys = (B, (layer["hin"] + 1) // 2, (layer["win"] + 1) // 2, N.cstride(cout))
        else:
            xs = (B, layer["hin"], layer["win"], N.cstride(cin))
            ys = (B, layer["hin"] * 2, layer["win"] * 2, N.cstride(cout))
        x = torch.randn(xs, device=dev)
        y = torch.empty(ys, device=dev)
        dy = torch.randn(ys, device=dev)
        dx = torch.empty(xs, device=dev)
        ws = torch.empty(int(lib.cgs_layer_workspace_bytes(C.byref(desc), B)), dtype=torch.uint8, device=dev)
        flops = 2.0 * S.layer_macs(layer) * B
        for name, fn in (("fwd", lambda: lib.cgs_layer_forward(C.byref(desc), L.MATH_IDS[a.math], B, L.ptr(x), L.ptr(y), L.ptr(ws), ws.numel(), L.stream_ptr())),
                         ("bwd", lambda: lib.cgs_layer_backward(C.byref(desc), L.MATH_IDS[a.math], B, L.ptr(dy), L.ptr(dx), L.ptr(x), 1, L.ptr(ws), ws.numel(), L.stream_ptr()))):
            L.check(fn())
            best = 1e9
            for _ in range(a.reps):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(a.inner):
                    fn()
                e.record()
                e.synchronize()
                best = min(best, s.elapsed_time(e) * 1e-3 / a.inner)
            tot += best
            totf += flops
            print("%-14s %8.1f us %7.1f TFLOP/s" % (layer["name"] + "." + name, best * 1e6, flops / best / 1e12))
    print("TOTAL %.1f us  %.1f TFLOP/s  (CGS_DEBUG=%s)" % (tot * 1e6, totf / tot / 1e12, os.environ.get("CGS_DEBUG", "0")))


if __name__ == "__main__":
    main()

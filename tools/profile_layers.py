"""Launch every GEMM pass of one refinement step once (after one warm-up each), for use under ncu.

    ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -o gpurun_out/prof \
        python tools/profile_layers.py --workload mnist --batch 1024
"""
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
    ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = L.load()
    arch = N.get_arch(a.workload)
    spec = N.NetSpec(arch, S.init_weights(arch, gain=2.5), dev, math=a.math)
    chain = [(l, spec.gtail.layer_desc(i)) for i, l in enumerate(arch["gtail"])] + \
            [(l, spec.d.layer_desc(i)) for i, l in enumerate(arch["d"][:-1])]
    B = a.batch
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
        for _ in range(a.reps + 1):
            L.check(lib.cgs_layer_forward(C.byref(desc), L.MATH_IDS[a.math], B, L.ptr(x), L.ptr(y), L.ptr(ws), ws.numel(), L.stream_ptr()))
            L.check(lib.cgs_layer_backward(C.byref(desc), L.MATH_IDS[a.math], B, L.ptr(dy), L.ptr(dx), L.ptr(x), 1, L.ptr(ws), ws.numel(), L.stream_ptr()))
        torch.cuda.synchronize()
        print(layer["name"], "ok")


if __name__ == "__main__":
    main()

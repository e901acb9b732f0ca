"""CPU replay of the gathered-GEMM lowering (test helper, not product code).

Given a layer description and its packed weights, recompute the layer pass exactly the way the CUDA kernel is
parameterised (``cgs_debug_gemm_params`` + packed matrices), in float64 numpy.  Comparing this replay with the
oracle's conv / deconv / autograd proves the host-side lowering and the packing; the GPU tests then only have to
prove the device code.
"""
import ctypes as C

import numpy as np

from cgs import lib as L
from cgs import nets as N

KMAXTAPS = 32


def gemm_params(layer, backward, B):
    lib = L.load()
    d = N._layer_desc(layer)
    n = L.check(lib.cgs_debug_gemm_params(C.byref(d), int(backward), B, None, 0))
    buf = np.zeros(n, np.int32)
    L.check(lib.cgs_debug_gemm_params(C.byref(d), int(backward), B, buf.ctypes.data, n))
    keys = ["IH", "IW", "Cs", "cblocks", "MH", "MW", "S", "M", "OH", "OW", "ON", "os", "N", "nclasses"]
    p = dict(zip(keys, buf[:14].tolist()))
    off = 14
    p["cls"] = []
    for _ in range(p["nclasses"]):
        k0, nkb, ntaps, oy0, ox0 = buf[off:off + 5].tolist()
        dy = buf[off + 5:off + 5 + KMAXTAPS].tolist()
        dx = buf[off + 5 + KMAXTAPS:off + 5 + 2 * KMAXTAPS].tolist()
        p["cls"].append(dict(k0=k0, nkb=nkb, ntaps=ntaps, oy0=oy0, ox0=ox0, dy=dy, dx=dx))
        off += 5 + 2 * KMAXTAPS
    return p


def replay(p, x, w, B):
    """x: [B, IH, IW, Cs] float64, w: [rows, K] -> raw accumulators laid out as [B, OH, OW, ON]."""
    x = np.asarray(x, np.float64).reshape(B, p["IH"], p["IW"], p["Cs"])
    w = np.asarray(w, np.float64)
    out = np.zeros((B, p["OH"], p["OW"], p["ON"]))
    cin = p["cblocks"] * 32 if p["cblocks"] else 4
    nvalid = min(p["N"], w.shape[0])
    for g in p["cls"]:
        # A matrix of this class: [B, MH, MW, nkb*32]
        A = np.zeros((B, p["MH"], p["MW"], g["nkb"] * 32))
        for t in range(g["ntaps"]):
            for j in range(p["MH"]):
                y = j * p["S"] + g["dy"][t]
                if not 0 <= y < p["IH"]:
                    continue
                for i in range(p["MW"]):
                    xx = i * p["S"] + g["dx"][t]
                    if not 0 <= xx < p["IW"]:
                        continue
                    A[:, j, i, t * cin:(t + 1) * cin] = x[:, y, xx, :cin]
        Wc = w[:nvalid, g["k0"]:g["k0"] + g["nkb"] * 32]
        R = A @ Wc.T                                         # [B, MH, MW, nvalid]
        out[:, g["oy0"]::p["os"], g["ox0"]::p["os"], :nvalid][:, :p["MH"], :p["MW"]] = R
    return out

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
    keys = ["IH", "IW", "Cs", "cblocks", "MH", "MW", "S", "M", "OH", "OW", "ON", "os", "N", "nclasses",
            "window", "win_k", "win_x0", "in_pitch_px"]
    p = dict(zip(keys, buf[:18].tolist()))
    off = 18
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


def pass_layout(layer, backward):
    return L.load().cgs_pass_layout(C.byref(N._layer_desc(layer)), int(backward))


def same_pad_before(size, k):
    out = (size + 1) // 2
    return max((out - 1) * 2 + k - size, 0) // 2


def col2im(col, layer, backward, B):
    """CPU twin of col2im_kernel: col [B, IH, IW, k*k*4] -> raw accumulators [B, OH, OW, 4]."""
    k = layer["k"]
    if not backward:     # deconv forward
        IH, IW = layer["hin"], layer["win"]
        OH, OW = IH * 2, IW * 2
    else:                # conv data-gradient
        OH, OW = layer["hin"], layer["win"]
        IH, IW = (OH + 1) // 2, (OW + 1) // 2
    py, px = same_pad_before(OH, k), same_pad_before(OW, k)
    col = np.asarray(col, np.float64).reshape(B, IH, IW, k * k * 4)
    out = np.zeros((B, OH, OW, 4))
    for y in range(OH):
        for ky in range(k):
            ty = y + py - ky
            if ty < 0 or ty % 2 or ty // 2 >= IH:
                continue
            for x in range(OW):
                for kx in range(k):
                    tx = x + px - kx
                    if tx < 0 or tx % 2 or tx // 2 >= IW:
                        continue
                    out[:, y, x, :] += col[:, ty // 2, tx // 2, (ky * k + kx) * 4:(ky * k + kx) * 4 + 4]
    return out


IMG_XOFF = 2


def replay_window(p, x_dense, w, B):
    """Window lowering: x_dense [B, IH, IW, 4] is re-laid out pitched ([IH][IW+8][4], data at column 2) and every
    K block ky of output pixel (j, i) is the 32 contiguous floats at stored pixel (2j + dy[ky], 2i + win_x0)."""
    IH, IW, P = p["IH"], p["IW"], p["in_pitch_px"]
    assert P == IW + 8
    img = np.zeros((B, IH, P, 4))
    img[:, :, IMG_XOFF:IMG_XOFF + IW, :] = np.asarray(x_dense, np.float64).reshape(B, IH, IW, 4)
    flat = img.reshape(B, IH, P * 4)
    g = p["cls"][0]
    w = np.asarray(w, np.float64)
    out = np.zeros((B, p["OH"], p["OW"], p["ON"]))
    for j in range(p["MH"]):
        for ky in range(g["ntaps"]):
            y = j * p["S"] + g["dy"][ky]
            if not 0 <= y < IH:
                continue
            for i in range(p["MW"]):
                s0 = (2 * i + p["win_x0"]) * 4
                out[:, j, i, :p["N"]] += flat[:, y, s0:s0 + 32] @ w[:p["N"], ky * 32:(ky + 1) * 32].T
    return out


def layer_pass(layer, backward, B, x_padded, w_packed):
    """Raw accumulators of one pass, replayed on the CPU whichever lowering the library picks."""
    p = gemm_params(layer, backward, B)
    if pass_layout(layer, backward) == 2:
        return replay_window(p, x_padded, w_packed, B)
    if pass_layout(layer, backward) == 1:
        col = replay(p, np.asarray(x_padded, np.float64).reshape(p["M"], 1, 1, p["Cs"]), w_packed, p["M"])
        return col2im(col.reshape(p["M"], p["ON"]), layer, backward, B)
    return replay(p, x_padded, w_packed, B)


def fusion_plan(layer, backward, B):
    """Decoded cgs_debug_fusion_plan (None when the pass is never class-fused)."""
    lib = L.load()
    d = N._layer_desc(layer)
    n = L.check(lib.cgs_debug_fusion_plan(C.byref(d), int(backward), B, None, 0))
    if n == 0:
        return None
    buf = np.zeros(n, np.int32)
    L.check(lib.cgs_debug_fusion_plan(C.byref(d), int(backward), B, buf.ctypes.data, n))
    v = buf.tolist()
    plan = dict(ns=v[0], groups=[], shifts=[])
    off = 2
    for _ in range(v[1]):
        plan["groups"].append(dict(nshifts=v[off], shift0=v[off + 1], ncls=v[off + 2], cls=v[off + 3:off + 7]))
        off += 7
    nshf = sum(g["nshifts"] for g in plan["groups"])
    for _ in range(nshf):
        s = v[off:off + 48]
        plan["shifts"].append(dict(dy=s[0], dx=s[1], ncls=s[2], nrun=s[3], slot=s[4:8], katom0=s[8:12], run_slot=s[12:16],
                                   run_len=s[16:20], run_acc=s[20:24], pc_slot=[s[24:28], s[28:32]],
                                   pc_half=[s[32:36], s[36:40]], pc_katom=[s[40:44], s[44:48]]))
        off += 48
    return plan

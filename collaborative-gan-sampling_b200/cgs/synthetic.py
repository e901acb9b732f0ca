"""Synthetic weights and proposals for benchmarks / smoke runs (no datasets or checkpoints are reachable).

Initialisers follow the reference: conv ``truncated_normal(0.02)`` (nsgan/ops.py:40), deconv /
linear ``normal(0.02)`` (ops.py:52,77), BN gamma~1 / beta~0 with non-trivial moving statistics so that folding is
exercised, seed 2019 (nsgan/main.py:13-16).  ``gain`` scales the kernels so a random-init net produces logits and
gradients of O(1) like a trained one (otherwise K refinement steps barely move anything).
"""
from __future__ import annotations

import math

import numpy as np


def init_weights(arch, seed=2019, gain=1.0, dtype=np.float32):
    rng = np.random.RandomState(seed)
    w = {}

    def tn(shape, std):
        x = rng.standard_normal(shape)
        bad = np.abs(x) > 2
        while bad.any():
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2
        return (x * std).astype(dtype)

    for scope, layers in (("generator", arch["gtail"]), ("discriminator", arch["d"]), ("generator", arch.get("head", []))):
        for L in layers:
            p = "%s/%s/" % (scope, L["name"])
            if (p + "w") in w or (p + "Matrix") in w:
                continue
            if L["type"] == "conv":
                w[p + "w"] = tn((L["k"], L["k"], L["cin"], L["cout"]), 0.02)
                w[p + "biases"] = (0.01 * rng.standard_normal(L["cout"])).astype(dtype)
            elif L["type"] == "deconv":
                w[p + "w"] = (0.02 * rng.standard_normal((L["k"], L["k"], L["cout"], L["cin"]))).astype(dtype)
                w[p + "biases"] = (0.01 * rng.standard_normal(L["cout"])).astype(dtype)
            else:
                w[p + "Matrix"] = (0.02 * rng.standard_normal((L["cin"], L["cout"]))).astype(dtype)
                w[p + "bias"] = (0.01 * rng.standard_normal(L["cout"])).astype(dtype)
            if L["bn"]:
                q = "%s/%s/" % (scope, L["bn"])
                c = L.get("bn_channels") or L["cout"]
                w[q + "gamma"] = (1.0 + 0.1 * rng.standard_normal(c)).astype(dtype)
                w[q + "beta"] = (0.05 * rng.standard_normal(c)).astype(dtype)
                w[q + "moving_mean"] = (0.1 * rng.standard_normal(c)).astype(dtype)
                w[q + "moving_variance"] = rng.uniform(0.5, 1.5, c).astype(dtype)
    if gain != 1.0:
        for k in list(w):
            if k.endswith("/w") or k.endswith("/Matrix"):
                w[k] = (w[k] * gain).astype(dtype)
    return w


def init_mlp2d(nhidden=64, nlayers=6, seed=2019, gain=1.0):
    rng = np.random.RandomState(seed)
    dims = [2] + [nhidden] * (nlayers - 1) + [1]
    ws = []
    for i in range(nlayers):
        lim = gain * math.sqrt(6.0 / (dims[i] + dims[i + 1]))
        k = rng.uniform(-lim, lim, (dims[i], dims[i + 1])).astype(np.float32)
        b = (0.1 * rng.standard_normal(dims[i + 1])).astype(np.float32)
        ws.append((k, b))
    return ws


def proposal_features(arch, batch, seed=0):
    """h0 = relu(N(0,1)) of the feature shape as a float32 numpy array [B,H,W,C]."""
    rng = np.random.RandomState(seed)
    return np.maximum(rng.standard_normal((batch,) + tuple(arch["feature_shape"])), 0).astype(np.float32)


def layer_macs(layer):
    """Forward MACs per sample, true taps only: conv k^2, stride-2 deconv k^2/4 per output."""
    if layer["type"] == "fc":
        return layer["cin"] * layer["cout"]
    if layer["type"] == "conv":
        ho, wo = (layer["hin"] + 1) // 2, (layer["win"] + 1) // 2
        return ho * wo * layer["cin"] * layer["cout"] * layer["k"] ** 2
    ho, wo = layer["hin"] * 2, layer["win"] * 2
    return ho * wo * layer["cin"] * layer["cout"] * layer["k"] ** 2 / 4.0


def refine_flops_per_sample(arch, steps):
    """2 * [(K+1) * MACs_fwd + K * MACs_bwd]  (SURVEY.md §8d)."""
    macs = sum(layer_macs(l) for l in arch["gtail"] + arch["d"])
    return 2.0 * ((steps + 1) * macs + steps * macs)

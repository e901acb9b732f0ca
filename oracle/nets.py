"""FP32 torch-CPU restatement of the reference's networks (TEST INFRASTRUCTURE ONLY).

Follows, line by line:
  * ``nsgan/ops.py:19-26``  bn  (decay .9, eps 1e-5, scale=True)
  * ``nsgan/ops.py:37-46``  conv2d   filter [kh,kw,Cin,Cout], SAME, stride 2, bias after conv
  * ``nsgan/ops.py:48-67``  deconv2d filter [kh,kw,Cout,Cin], conv2d_transpose, stride 2
  * ``nsgan/ops.py:69-70``  lrelu = max(x, 0.2 x)
  * ``nsgan/ops.py:72-83``  linear  Matrix [in,out] + bias
  * ``nsgan/GAN.py:59-70``  discriminator (infoGAN MNIST)
  * ``nsgan/GAN.py:87-101`` input_to_feature / feature_to_data
  * ``synthetic/GAN.py:28-37,108-111`` 2-D discriminator MLP, fake_sigmoid, fake_saliency
DCGAN-32/64 (BASELINE configs C3/C4) are NOT in the reference tree; their shapes are
the upstream carpedm20/DCGAN-tensorflow ones the reference's ``ops.py`` was taken from
(``nsgan/ops.py:2``), k=5 s=2 defaults at ``nsgan/ops.py:37,48``.

PARITY UNPINNED for the image nets (no TF here, no reference vectors) -- see
``oracle/__init__.py``.  All tensors are NHWC at the interface, like TF.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5      # nsgan/ops.py:23
BN_DECAY = 0.9     # nsgan/ops.py:21
LRELU_LEAK = 0.2   # nsgan/ops.py:69


# --------------------------------------------------------------------------------------
# architecture descriptions (plain data; the product has its own builder, tests compare)
# --------------------------------------------------------------------------------------

def conv_out_size_same(size, stride):      # nsgan/ops.py:28-29
    return int(math.ceil(float(size) / float(stride)))


def arch_mnist_infogan(layer=1):
    """nsgan/GAN.py:59-101; refinement at the [7,7,128] map (GAN.py:172, 87-92) for layer=1, or at the
    [14,14,64] map feeding the last deconv for layer=2 (BASELINE config wording; same code path)."""
    assert layer in (1, 2)
    gtail = [
        dict(type="deconv", name="g_dc3", k=4, cin=128, cout=64, hin=7, win=7, bn="g_bn3", act="relu"),
        dict(type="deconv", name="g_dc4", k=4, cin=64, cout=1, hin=14, win=14, bn=None, act="tanh"),
    ][layer - 1:]
    return {
        "name": "mnist_infogan" if layer == 1 else "mnist_infogan_l2",
        "feature_shape": [7, 7, 128] if layer == 1 else [14, 14, 64],
        "image_shape": [28, 28, 1],
        # proposal head = nsgan/GAN.py:87-92 input_to_feature (BN inference); only used to produce proposals from z
        "z_dim": 62,
        "head": [
            dict(type="fc", name="g_fc1", cin=62, cout=1024, bn="g_bn1", act="relu"),
            dict(type="fc", name="g_fc2", cin=1024, cout=6272, bn="g_bn2", act="relu"),
        ] + ([dict(type="deconv", name="g_dc3", k=4, cin=128, cout=64, hin=7, win=7, bn="g_bn3", act="relu")]
             if layer == 2 else []),
        "head_reshape": [7, 7, 128],
        "gtail": gtail,
        "d": [
            dict(type="conv", name="d_conv1", k=4, cin=1, cout=64, hin=28, win=28, bn=None, act="lrelu"),
            dict(type="conv", name="d_conv2", k=4, cin=64, cout=128, hin=14, win=14, bn="d_bn2", act="lrelu"),
            dict(type="fc", name="d_fc3", cin=6272, cout=1024, bn="d_bn3", act="lrelu"),
            dict(type="fc", name="d_fc4", cin=1024, cout=1, bn=None, act="none"),
        ],
    }


def arch_dcgan(size=64, layer=1, gf=64, df=64, c_dim=3, k=5):
    """Upstream DCGAN-tensorflow generator tail from activation map `layer` (1..4) + discriminator.

    G: h0 [s16,s16,gf*8] -dc1-> [s8,s8,gf*4] -dc2-> [s4,s4,gf*2] -dc3-> [s2,s2,gf] -dc4-> [s,s,c] tanh.
    D: conv(c->df) lrelu, conv(->2df)+bn lrelu, conv(->4df)+bn lrelu, conv(->8df)+bn lrelu, linear->1.
    """
    assert 1 <= layer <= 4
    s = [size]
    for _ in range(4):
        s.append(conv_out_size_same(s[-1], 2))
    s16, s8, s4, s2 = s[4], s[3], s[2], s[1]
    gch = [gf * 8, gf * 4, gf * 2, gf, c_dim]
    gsz = [s16, s8, s4, s2, size]
    gtail = []
    for i in range(layer - 1, 4):
        last = i == 3
        gtail.append(dict(type="deconv", name="g_h%d" % (i + 1), k=k, cin=gch[i], cout=gch[i + 1],
                          hin=gsz[i], win=gsz[i], bn=None if last else "g_bn%d" % (i + 1),
                          act="tanh" if last else "relu"))
    dch = [c_dim, df, df * 2, df * 4, df * 8]
    dsz = [size, s2, s4, s8, s16]
    d = []
    for i in range(4):
        d.append(dict(type="conv", name="d_h%d_conv" % i, k=k, cin=dch[i], cout=dch[i + 1],
                      hin=dsz[i], win=dsz[i], bn=None if i == 0 else "d_bn%d" % i, act="lrelu"))
    d.append(dict(type="fc", name="d_h4_lin", cin=s16 * s16 * df * 8, cout=1, bn=None, act="none"))
    head = [dict(type="fc", name="g_h0_lin", cin=100, cout=gsz[0] * gsz[0] * gch[0], bn="g_bn0", bn_channels=gch[0],
                 act="relu")]
    for i in range(layer - 1):
        head.append(dict(type="deconv", name="g_h%d" % (i + 1), k=k, cin=gch[i], cout=gch[i + 1], hin=gsz[i], win=gsz[i],
                         bn="g_bn%d" % (i + 1), act="relu"))
    return {
        "name": "dcgan%d_l%d" % (size, layer),
        "feature_shape": [gsz[layer - 1], gsz[layer - 1], gch[layer - 1]],
        "image_shape": [size, size, c_dim],
        "z_dim": 100, "head": head, "head_reshape": [gsz[0], gsz[0], gch[0]],
        "gtail": gtail,
        "d": d,
    }


def get_arch(name):
    if name in ("mnist", "mnist_infogan"):
        return arch_mnist_infogan()
    if name in ("mnist_l2", "mnist_infogan_l2"):
        return arch_mnist_infogan(2)
    if name.startswith("dcgan"):
        body = name[len("dcgan"):]
        size, _, layer = body.partition("_l")
        return arch_dcgan(int(size), int(layer or 1))
    raise KeyError(name)


# --------------------------------------------------------------------------------------
# weights (TF variable layouts; synthetic init per SURVEY.md §8d)
# --------------------------------------------------------------------------------------

def init_weights(arch, seed=2019, dtype=np.float32):
    """Random-init weights in TF variable layout.

    conv w ~ truncated_normal(.02) (ops.py:40), deconv/linear ~ normal(.02) (ops.py:52,77),
    biases 0 (ops.py:43,61,78).  BN gamma=1, beta=0 as TF initialises them, but with
    NON-trivial moving statistics (mean~N(0,.1), var~U(.5,1.5)) and small non-zero
    biases/betas so folding and bias paths are exercised (synthetic).
    Seed 2019 = nsgan/main.py:13-16.
    """
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
    return w


def scale_weights_for_signal(arch, w, gain=4.0):
    """Optional: multiply conv/deconv/linear kernels so that random-init nets give logits and
    gradients of O(1) instead of O(1e-3) (a trained net's regime).  Pure test utility."""
    out = dict(w)
    for k, v in w.items():
        if k.endswith("/w") or k.endswith("/Matrix"):
            out[k] = (v * gain).astype(v.dtype)
    return out


# --------------------------------------------------------------------------------------
# TF-semantics ops in torch (NHWC in / NHWC out)
# --------------------------------------------------------------------------------------

def _same_pads(size, k, stride=2):
    out = conv_out_size_same(size, stride)
    total = max((out - 1) * stride + k - size, 0)
    before = total // 2          # TF puts the extra pixel AFTER
    return before, total - before


def conv2d_same(x_nhwc, w_tf, bias, stride=2):
    """tf.nn.conv2d(x, w, [1,2,2,1], 'SAME') + bias_add  (nsgan/ops.py:41-44)."""
    kh, kw = w_tf.shape[0], w_tf.shape[1]
    x = x_nhwc.permute(0, 3, 1, 2)
    pt, pb = _same_pads(x.shape[2], kh, stride)
    pl, pr = _same_pads(x.shape[3], kw, stride)
    x = F.pad(x, (pl, pr, pt, pb))
    wt = w_tf.permute(3, 2, 0, 1).contiguous()            # [Cout,Cin,kh,kw]
    y = F.conv2d(x, wt, bias=None, stride=stride)
    y = y + bias.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1).contiguous()


def deconv2d_same(x_nhwc, w_tf, bias, stride=2):
    """tf.nn.conv2d_transpose(x, w, output_shape=2*in, [1,2,2,1]) + bias (nsgan/ops.py:55-62).

    conv2d_transpose is the data-gradient of the SAME conv whose input is the [2H,2W] output:
    y[2i+ky-pt, 2j+kx-pl] += x[i,j] w[ky,kx], pt = pad-before of that conv, cropped to [0,2H).
    """
    kh, kw = w_tf.shape[0], w_tf.shape[1]
    x = x_nhwc.permute(0, 3, 1, 2)
    H, W = x.shape[2], x.shape[3]
    pt, _ = _same_pads(H * stride, kh, stride)
    pl, _ = _same_pads(W * stride, kw, stride)
    wt = w_tf.permute(3, 2, 0, 1).contiguous()            # [Cin,Cout,kh,kw]
    full = F.conv_transpose2d(x, wt, bias=None, stride=stride)   # [(H-1)*s+k]
    y = full[:, :, pt:pt + H * stride, pl:pl + W * stride]
    y = y + bias.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1).contiguous()


def batch_norm(x, gamma, beta, mmean, mvar, mode):
    """tf.contrib.layers.batch_norm (nsgan/ops.py:19-26) over the last (channel) axis.

    mode 'inference': moving statistics.  mode 'batch': batch mean / biased variance with the
    gradient flowing through the statistics (is_training=True, nsgan/GAN.py:175).  The
    moving-average side effect of training mode (updates_collections=None) is returned by
    ``bn_batch_stat_update`` for completeness; it does not affect the value computed here.
    """
    if mode == "inference":
        mean, var = mmean, mvar
    elif mode == "batch":
        dims = tuple(range(x.dim() - 1))
        mean = x.mean(dim=dims)
        var = ((x - mean) ** 2).mean(dim=dims)
    else:
        raise ValueError(mode)
    return (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta


def bn_batch_stat_update(x, mmean, mvar):
    dims = tuple(range(x.dim() - 1))
    mean = x.mean(dim=dims)
    var = ((x - mean) ** 2).mean(dim=dims)
    return BN_DECAY * mmean + (1 - BN_DECAY) * mean, BN_DECAY * mvar + (1 - BN_DECAY) * var


def activation(x, act):
    if act == "relu":
        return torch.relu(x)
    if act == "lrelu":
        return torch.maximum(x, LRELU_LEAK * x)        # nsgan/ops.py:69-70
    if act == "tanh":
        return torch.tanh(x)
    if act == "none":
        return x
    raise ValueError(act)


def _t(w, key):
    v = w[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


def run_layers(x, layers, scope, w, bn_mode, collect=None):
    """Apply conv/deconv/fc layers (bias -> BN -> activation, as nsgan/GAN.py:64-68,96-100)."""
    for L in layers:
        p = "%s/%s/" % (scope, L["name"])
        if L["type"] == "conv":
            x = conv2d_same(x, _t(w, p + "w"), _t(w, p + "biases"))
        elif L["type"] == "deconv":
            x = deconv2d_same(x, _t(w, p + "w"), _t(w, p + "biases"))
        else:
            x = x.reshape(x.shape[0], -1)                 # NHWC flatten, nsgan/GAN.py:66
            x = x @ _t(w, p + "Matrix") + _t(w, p + "bias")
        if L["bn"]:
            q = "%s/%s/" % (scope, L["bn"])
            shp = x.shape
            if L.get("bn_channels"):             # DCGAN: linear -> reshape [.., C] -> bn0 (per channel)
                x = x.reshape(-1, L["bn_channels"])
            x = batch_norm(x, _t(w, q + "gamma"), _t(w, q + "beta"),
                           _t(w, q + "moving_mean"), _t(w, q + "moving_variance"), bn_mode)
            x = x.reshape(shp)
        x = activation(x, L["act"])
        if collect is not None:
            collect.append(x)
    return x


def input_to_feature(z, arch, w):
    """nsgan/GAN.py:87-92 (MNIST) / upstream DCGAN generator up to the refined map; BN in inference mode."""
    x = torch.as_tensor(z, dtype=torch.float32)
    fcs = [L for L in arch["head"] if L["type"] == "fc"]
    rest = [L for L in arch["head"] if L["type"] != "fc"]
    x = run_layers(x, fcs, "generator", w, "inference")
    x = x.reshape(x.shape[0], *arch["head_reshape"])
    if rest:
        x = run_layers(x, rest, "generator", w, "inference")
    return x


def feature_to_data(h, arch, w, collect=None):
    """nsgan/GAN.py:94-101 -- generator tail, BN in inference mode (is_training=False default)."""
    return run_layers(h, arch["gtail"], "generator", w, "inference", collect)


def discriminator(x, arch, w, d_bn="inference", collect=None):
    """nsgan/GAN.py:59-70.  The reference refines through D with is_training=True
    (GAN.py:175, d_bn='batch'); the B200 build uses inference statistics (north_star)."""
    return run_layers(x, arch["d"], "discriminator", w, d_bn, collect)


# --------------------------------------------------------------------------------------
# 2-D discriminator MLP (synthetic/GAN.py:28-37) + fake_sigmoid / fake_saliency (:108-111)
# --------------------------------------------------------------------------------------

def init_mlp2d(nhidden=64, nlayers=6, seed=2019, gain=1.0):
    """tf.layers.dense default init = glorot_uniform kernels, zero bias; small non-zero biases
    are drawn instead so the bias path is exercised (synthetic weights)."""
    rng = np.random.RandomState(seed)
    dims = [2] + [nhidden] * (nlayers - 1) + [1]
    ws = []
    for i in range(nlayers):
        lim = gain * math.sqrt(6.0 / (dims[i] + dims[i + 1]))
        k = rng.uniform(-lim, lim, (dims[i], dims[i + 1])).astype(np.float32)
        b = (0.1 * rng.standard_normal(dims[i + 1])).astype(np.float32)
        ws.append((k, b))
    return ws


def mlp2d_logit(x_t, ws):
    net = x_t
    for i, (k, b) in enumerate(ws):
        net = net @ torch.from_numpy(k) + torch.from_numpy(b)
        if i < len(ws) - 1:
            net = torch.relu(net)            # synthetic/GAN.py:31,34
    return net                               # [N,1]


def mlp2d_sigmoid_saliency(x, ws):
    """(fake_sigmoid [N,1], fake_saliency [N,2]) for a fed batch, FP32.

    fake_loss = reduce_mean(sigmoid_cross_entropy_with_logits(logit, ones))  (GAN.py:109-110)
    fake_saliency = d fake_loss / d fake_samples                           (GAN.py:111)
    => carries a 1/N factor, N = rows fed.
    """
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).requires_grad_(True)
    logit = mlp2d_logit(xt, ws)
    loss = F.softplus(-logit).mean()
    (g,) = torch.autograd.grad(loss, xt)
    return torch.sigmoid(logit).detach().numpy(), g.numpy()

"""Weight interchange: externally stored variables -> the name -> array mapping ``NetSpec`` packs.

The canonical form is the reference's TF-1.x variables by name and layout (``nsgan/ops.py``):
  ``<scope>/<layer>/w``       conv   [kh, kw, Cin, Cout] (ops.py:39)      deconv [kh, kw, Cout, Cin] (ops.py:51)
  ``<scope>/<layer>/biases``  [Cout] (ops.py:43,61)
  ``<scope>/<layer>/Matrix``  linear [in, out], ``<scope>/<layer>/bias`` [out] (ops.py:75-79)
  ``<scope>/<bn>/beta|gamma|moving_mean|moving_variance`` [C] (tf.contrib.layers.batch_norm, ops.py:19-26)
with scope ``generator`` / ``discriminator`` (nsgan/GAN.py:62,75).

A TF checkpoint is turned into such a file where TensorFlow is installed with
``r = tf.train.load_checkpoint(ckpt); np.savez(out, **{n: r.get_tensor(n) for n in r.get_variable_to_shape_map()})``
(the checkpoint container format itself is not parsed here).  ``normalize_names`` strips what a Saver adds (``:0``
suffixes, optimizer slots, counters); ``from_torch_layout`` converts PyTorch-ordered kernels
(conv [Cout, Cin, kh, kw], transposed conv [Cin, Cout, kh, kw], linear [out, in]) to the TF order.
"""
from __future__ import annotations

import re

import numpy as np

_SLOT = re.compile(r"/(Adam(_\d+)?|Momentum|RMSProp(_\d+)?)$")
_SKIP = ("beta1_power", "beta2_power", "global_step")


def normalize_names(variables):
    """Drop ``:0`` suffixes, optimizer slot variables and Saver counters; arrays become float32 numpy."""
    out = {}
    for name, value in variables.items():
        n = name[:-2] if name.endswith(":0") else name
        if _SLOT.search(n) or n.split("/")[-1] in _SKIP or n in _SKIP:
            continue
        v = value.detach().cpu().numpy() if hasattr(value, "detach") else np.asarray(value)
        out[n] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def load_npz(path):
    with np.load(path) as z:
        return normalize_names({k: z[k] for k in z.files})


def save_npz(path, variables):
    np.savez(path, **{k: np.asarray(v) for k, v in variables.items()})


def _layers(arch, include_head):
    groups = [("generator", arch["gtail"]), ("discriminator", arch["d"])]
    if include_head:
        groups.append(("generator", arch.get("head", [])))
    seen = set()
    for scope, layers in groups:
        for layer in layers:
            key = (scope, layer["name"])
            if key not in seen:
                seen.add(key)
                yield scope, layer


def expected_shapes(arch, include_head=False):
    """name -> shape of every variable the nets of ``arch`` need, in the canonical (TF) layout."""
    shapes = {}
    for scope, L in _layers(arch, include_head):
        p = "%s/%s/" % (scope, L["name"])
        if L["type"] == "conv":
            shapes[p + "w"] = (L["k"], L["k"], L["cin"], L["cout"])
            shapes[p + "biases"] = (L["cout"],)
        elif L["type"] == "deconv":
            shapes[p + "w"] = (L["k"], L["k"], L["cout"], L["cin"])
            shapes[p + "biases"] = (L["cout"],)
        else:
            shapes[p + "Matrix"] = (L["cin"], L["cout"])
            shapes[p + "bias"] = (L["cout"],)
        if L["bn"]:
            q = "%s/%s/" % (scope, L["bn"])
            c = L.get("bn_channels") or L["cout"]
            for f in ("gamma", "beta", "moving_mean", "moving_variance"):
                shapes[q + f] = (c,)
    return shapes


def validate(arch, variables, include_head=False):
    """Raise KeyError / ValueError with the offending name when a variable is missing or laid out differently."""
    for name, shape in expected_shapes(arch, include_head).items():
        if name not in variables:
            raise KeyError("weight file has no variable %r (needed by architecture %s)" % (name, arch["name"]))
        got = tuple(np.shape(variables[name]))
        if got != shape:
            hint = ""
            if len(shape) == 4 and sorted(got) == sorted(shape):
                hint = " -- same extents in another order: is this a PyTorch-ordered kernel? (see from_torch_layout)"
            raise ValueError("variable %r has shape %s, expected %s%s" % (name, got, shape, hint))
    return variables


def from_torch_layout(arch, variables, include_head=False):
    """PyTorch-ordered kernels -> TF order: conv [Cout,Cin,kh,kw] -> [kh,kw,Cin,Cout]; transposed conv
    [Cin,Cout,kh,kw] -> [kh,kw,Cout,Cin]; linear [out,in] -> [in,out] (inverse of SURVEY App. A8's permutes)."""
    out = dict(normalize_names(variables))
    for scope, L in _layers(arch, include_head):
        p = "%s/%s/" % (scope, L["name"])
        if L["type"] in ("conv", "deconv"):
            out[p + "w"] = np.ascontiguousarray(np.transpose(out[p + "w"], (2, 3, 1, 0)))
        else:
            out[p + "Matrix"] = np.ascontiguousarray(out[p + "Matrix"].T)
    return out


def to_torch_layout(arch, variables, include_head=False):
    """Inverse of ``from_torch_layout`` (used by the tests to produce externally laid-out files)."""
    out = dict(normalize_names(variables))
    for scope, L in _layers(arch, include_head):
        p = "%s/%s/" % (scope, L["name"])
        if L["type"] in ("conv", "deconv"):
            out[p + "w"] = np.ascontiguousarray(np.transpose(out[p + "w"], (3, 2, 0, 1)))
        else:
            out[p + "Matrix"] = np.ascontiguousarray(out[p + "Matrix"].T)
    return out

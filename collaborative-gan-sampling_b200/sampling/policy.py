"""Drop-in for the reference's ``sampling/policy.py`` (PolicyAdaptive), backed by ``cgs_policy_step``.

Same constructor, fields and method names as ``sampling/policy.py:5-64``.  ``apply_gradient`` updates ``theta``
IN PLACE on the GPU (numpy inputs are staged to the device and copied back in place, as the reference's numpy
branch mutates its argument, policy.py:28,36,61) and returns it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

try:
    from . import _paths  # noqa: F401
except ImportError:  # flat import, reference style
    import _paths  # noqa: F401
from cgs import lib as L
from cgs import runtime as R


class PolicyAdaptive(object):
    def __init__(self, step_size, method):
        self.method = method
        self.lambda_ = step_size
        self.alpha_ = 0.9
        self.momentum = None
        self.beta1_ = 0.9
        self.beta2_ = 0.5
        self.beta3_ = 0.5
        self.degree_ = 2
        self.eps_ = 1e-8
        self.mean_square = None
        self.loss = None

    def reset_moving_average(self):
        self.momentum = None
        self.mean_square = None
        self.loss = None

    def config(self):
        if self.method not in L.POLICY_IDS:
            raise NotImplementedError(self.method)       # policy.py:64
        c = L.PolicyCfg()
        c.method = L.POLICY_IDS[self.method]
        c.degree = int(self.degree_)
        c.step_size, c.alpha = float(self.lambda_), float(self.alpha_)
        c.beta1, c.beta2, c.beta3, c.eps = float(self.beta1_), float(self.beta2_), float(self.beta3_), float(self.eps_)
        return c

    def apply_gradient(self, theta, grad, loss=None):
        cfg = self.config()
        lib = L.load()
        th, th_np = R.to_device(theta, torch.float32)
        g, _ = R.to_device(grad, torch.float32)
        if th.shape != g.shape:
            raise ValueError("theta and grad shapes differ")
        rows = th.shape[0]
        cols = th.numel() // max(rows, 1)
        first = self.momentum is None
        ls = None
        if self.method == "ladam":
            if loss is None:
                # the reference evaluates `None + 0.5` here (policy.py:51,56) -> TypeError
                raise TypeError("unsupported operand type(s) for +: 'NoneType' and 'float'")
            ls, _ = R.to_device(loss, torch.float32)
            ls = ls.reshape(-1)
            if ls.numel() != rows:
                raise ValueError("loss must have one entry per row")
        # the state of an earlier call must fit this theta: the reference's numpy / TF arithmetic raises a broadcast
        # error when the batch shape changes without reset_moving_average(); the kernel would index out of bounds
        for name in ("momentum", "mean_square"):
            st = getattr(self, name)
            if st is not None and (tuple(st.shape) != tuple(th.shape) or st.device != th.device):
                raise ValueError("operands could not be broadcast together with shapes %s %s (policy state `%s`; call "
                                 "reset_moving_average() when the batch changes)" % (tuple(st.shape), tuple(th.shape), name))
        if self.loss is not None and (self.loss.numel() != rows or self.loss.device != th.device):
            raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,) (policy state `loss`)"
                             % (self.loss.numel(), rows))
        if self.method != "sgd" and first:
            self.momentum = torch.empty_like(th)
        if self.method == "ladam" and first:
            self.mean_square = torch.empty_like(th)
            self.loss = torch.empty(rows, dtype=torch.float32, device=th.device)
        L.check(lib.cgs_policy_step(C.byref(cfg), L.ptr(th), L.ptr(g), L.ptr(ls), L.ptr(self.momentum),
                                    L.ptr(self.mean_square), L.ptr(self.loss), int(first), rows, cols,
                                    L.stream_ptr()))
        if th_np:
            np.copyto(theta, th.cpu().numpy().reshape(theta.shape))
            return theta
        if th.data_ptr() != theta.data_ptr():
            theta.copy_(th.reshape(theta.shape))
        return theta

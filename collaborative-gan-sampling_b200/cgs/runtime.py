"""Small host-side helpers shared by the drop-in sampling classes (device staging, workspace reuse)."""
from __future__ import annotations

import numpy as np
import torch

from . import lib as L


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("libcgs needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(x, dtype=None, device=None):
    """numpy array / torch tensor (host or device) -> contiguous device tensor.  Returns (tensor, was_numpy)."""
    dev = device or require_cuda()
    was_numpy = not isinstance(x, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(x)) if was_numpy else x
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    return t.contiguous(), was_numpy


def back(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


class Workspace:
    """Grow-only device scratch buffer owned by the caller side of the ABI."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        nbytes = int(nbytes)
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        return self.buf


def score_dtype(t):
    if t.dtype == torch.float64:
        return L.F64
    if t.dtype == torch.float32:
        return L.F32
    raise TypeError("scores must be float32 or float64, got %s" % t.dtype)

"""Drop-in for the reference's ``sampling/rejector.py`` (Rejector), backed by ``cgs_drs_accept``.

Same names and defaults as ``sampling/rejector.py:7-38``: ``D_tilde_M``, ``set_score_max``,
``sampling(samples, sigmoids, epsilon=1e-8, shift_percent=60.0, ranking=None)``.

Randomness: by default one ``np.random.rand(N)`` is drawn on the host per call, exactly where the reference
draws it (rejector.py:33), so the same ``np.random.seed`` gives the same decisions.  ``rng='philox'`` switches
to the on-device counter-based stream (no host RNG, no H2D copy); ``uniforms=`` passes them explicitly.
"""
from __future__ import annotations

import numpy as np
import torch

try:
    from . import _paths  # noqa: F401
except ImportError:
    import _paths  # noqa: F401
from cgs import lib as L
from cgs import runtime as R


class Rejector(object):
    def __init__(self, rng="numpy", seed=0):
        self._m = None            # device scalar (float64)
        self._m_host = 0.0        # rejector.py:9
        self.rng = rng
        self.seed = int(seed)
        self.offset = 0
        self._ws = R.Workspace()
        self.last_accept = None   # flags of the last call (device uint8)
        self.last_indices = None  # accepted row indices of the last call (device int32, ascending)

    # D_tilde_M mirrors the reference attribute; reading it synchronises.
    @property
    def D_tilde_M(self):
        return self._m_host if self._m is None else float(self._m.item())

    @D_tilde_M.setter
    def D_tilde_M(self, v):
        self._m_host = float(v)
        self._m = None

    def _state(self, device):
        if self._m is None or self._m.device != device:
            self._m = torch.tensor([self._m_host], dtype=torch.float64, device=device)
        return self._m

    def set_score_max(self, score_max):
        dev = R.require_cuda()
        s, _ = R.to_device(np.asarray(score_max) if not isinstance(score_max, torch.Tensor) else score_max)
        if s.dtype not in (torch.float32, torch.float64):
            s = s.to(torch.float64)
        s = s.reshape(-1)[:1].contiguous()
        m = self._state(dev)
        L.check(L.load().cgs_drs_set_score_max(L.ptr(s), R.score_dtype(s), L.ptr(m), L.stream_ptr()))

    def select(self, sigmoids, epsilon=1e-8, shift_percent=60.0, uniforms=None):
        """Accept / reject on the scores only; returns the accepted rows (device int32, ascending)."""
        dev = R.require_cuda()
        lib = L.load()
        sig, _ = R.to_device(sigmoids)
        if sig.dtype not in (torch.float32, torch.float64):
            sig = sig.to(torch.float64)
        sig = sig.reshape(-1)
        n = sig.numel()
        u = None
        seed, offset = 0, 0
        if n >= 2 ** 31 - 2:
            raise ValueError("n too large for 32-bit row indices; split the batch")
        if uniforms is not None:
            u, _ = R.to_device(uniforms, torch.float64)
            u = u.reshape(-1)
            if u.numel() < n:                                         # the kernel reads uniforms[i] for every row
                raise ValueError("uniforms has %d entries, %d rows need one each (rejector.py:33)" % (u.numel(), n))
        elif self.rng == "numpy":
            u, _ = R.to_device(np.random.rand(n), torch.float64)      # rejector.py:33
        elif self.rng == "philox":
            seed, offset = self.seed, self.offset
            self.offset += n
        else:
            raise ValueError("rng must be 'numpy' or 'philox'")
        m = self._state(dev)
        accept = torch.empty(n, dtype=torch.uint8, device=dev)
        idx = torch.empty(n, dtype=torch.int32, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = self._ws.get(lib.cgs_drs_workspace_bytes(n), dev)
        sp = -1.0 if shift_percent is None else float(shift_percent)
        L.check(lib.cgs_drs_accept(L.ptr(sig), R.score_dtype(sig), n, L.ptr(u), seed, offset, L.ptr(m), float(epsilon),
                                   sp, L.ptr(accept), L.ptr(idx), L.ptr(count), None, L.ptr(ws), ws.numel(),
                                   L.stream_ptr()))
        k = int(count.item())                 # the one sync: the output is sized by the count
        self.last_accept, self.last_indices = accept, idx[:k]
        self._last_count = count
        return self.last_indices

    def sampling(self, samples, sigmoids, epsilon=1e-8, shift_percent=60.0, ranking=None, uniforms=None):
        if ranking is not None:
            raise NotImplementedError        # rejector.py:35-36
        dev = R.require_cuda()
        lib = L.load()
        smp, smp_np = R.to_device(samples)
        nsig = sigmoids.shape[0] if hasattr(sigmoids, 'shape') else len(sigmoids)
        if smp.shape[0] != nsig:
            raise IndexError("boolean index did not match indexed array along dimension 0")
        idx = self.select(sigmoids, epsilon, shift_percent, uniforms)
        k = idx.numel()
        count = self._last_count
        out = torch.empty((k,) + tuple(smp.shape[1:]), dtype=smp.dtype, device=dev)
        if k:
            row_bytes = smp[0].numel() * smp.element_size()
            L.check(lib.cgs_gather_rows(L.ptr(smp), row_bytes, L.ptr(idx), L.ptr(count), k, L.ptr(out), L.stream_ptr()))
        return R.back(out, smp_np)

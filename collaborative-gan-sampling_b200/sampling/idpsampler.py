"""Drop-in for the reference's ``sampling/idpsampler.py`` (IndependenceSampler), backed by ``cgs_mh_accept``.

Same names as ``sampling/idpsampler.py:4-53``: ``IndependenceSampler(T=5, B=0)``, ``d_curr``, ``cnt_chain``,
``thin_period``, ``burn_in``, ``set_score_curr``, ``sampling``, ``next``.  The sequential chain is evaluated by
an exact parallel restatement on the GPU; emitted rows, ``d_curr`` and ``cnt_chain`` match the reference's
python loop given the same uniforms (one ``np.random.uniform`` per row, idpsampler.py:50).
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

_KIND_NONE, _KIND_F32, _KIND_F64, _KIND_WEAK = 0, 1, 2, 3


def _kind_of(v):
    """How numpy-2 promotion would type ``v`` inside idpsampler.py:48."""
    if v is None:
        return _KIND_NONE, 0.0
    if isinstance(v, torch.Tensor):
        k = _KIND_F64 if v.dtype == torch.float64 else _KIND_F32
        return k, float(v.reshape(-1)[0].item())
    if isinstance(v, (np.ndarray, np.generic)):
        a = np.asarray(v)
        k = _KIND_F64 if a.dtype == np.float64 else _KIND_F32
        return k, float(a.reshape(-1)[0])
    return _KIND_WEAK, float(v)      # python scalar: weak, adopts the dtype of the scores


class IndependenceSampler():
    def __init__(self, T=5, B=0, rng="numpy", seed=0):
        self._d_host = None
        self._state = None         # device: d_curr f64, kind i32, cnt i32
        self._cnt_host = 1
        self.thin_period = T
        self.burn_in = B
        self.rng = rng
        self.seed = int(seed)
        self.offset = 0
        self._ws = R.Workspace()
        self.last_accepted = None
        self.last_emit_src = None

    # --- state mirrors (reading synchronises) ---
    def _pull(self):
        if self._state is not None:
            d, kind, cnt = self._state
            k = int(kind.item())
            dv = float(d.item())
            self._cnt_host = int(cnt.item())
            if k == _KIND_NONE:
                self._d_host = None
            elif k == _KIND_F32:
                self._d_host = np.float32(dv)
            elif k == _KIND_F64:
                self._d_host = np.float64(dv)
            else:
                self._d_host = dv
            self._state = None

    @property
    def d_curr(self):
        self._pull()
        return self._d_host

    @d_curr.setter
    def d_curr(self, v):
        self._pull()
        self._d_host = v

    @property
    def cnt_chain(self):
        self._pull()
        return self._cnt_host

    @cnt_chain.setter
    def cnt_chain(self, v):
        self._pull()
        self._cnt_host = int(v)

    def set_score_curr(self, d_curr):
        self.d_curr = d_curr          # idpsampler.py:11-15

    def _push(self, dev):
        if self._state is None:
            kind, val = _kind_of(self._d_host)
            self._state = (torch.tensor([val], dtype=torch.float64, device=dev),
                           torch.tensor([kind], dtype=torch.int32, device=dev),
                           torch.tensor([self._cnt_host], dtype=torch.int32, device=dev))
        return self._state

    def emit_capacity(self, n):
        """Upper bound on the rows one call over ``n`` scores can emit (the thinning counter emits at most once per
        ``thin_period + 1`` processed rows, idpsampler.py:34-39; +2 covers the carried counter)."""
        return int(n) // (int(self.thin_period) + 1) + 2

    def select_async(self, sigmoids, uniforms=None):
        """``select`` without any host synchronisation: returns (emit_src [emit_capacity(n)] int32, count [1] int32),
        both on the device; rows past ``count`` are undefined.  The score-range assertion of idpsampler.py:22-23
        is skipped (it would need a host read); everything else is the same chain."""
        return self._select(sigmoids, uniforms, sync=False)

    def gather_async(self, samples):
        """Rows emitted by the last ``select_async`` -> (rows [emit_capacity, ...] float32, count [1]); device only."""
        dev = R.require_cuda()
        smp, _ = R.to_device(samples)
        emit, count = self.last_emit_src, self._last_count
        src = (smp if smp.dtype == torch.float32 else smp.to(torch.float32)).contiguous()
        cap = emit.numel()
        out = torch.empty((cap,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
        if cap and src.shape[0]:
            L.check(L.load().cgs_gather_rows(L.ptr(src), src[0].numel() * 4, L.ptr(emit), L.ptr(count), cap, L.ptr(out),
                                             L.stream_ptr()))
        return out, count

    def select(self, sigmoids, uniforms=None):
        """Run the chain over the scores only; returns the emitted source rows (device int32, ascending)."""
        return self._select(sigmoids, uniforms, sync=True)

    def _select(self, sigmoids, uniforms, sync):
        dev = R.require_cuda()
        lib = L.load()
        sig, _ = R.to_device(sigmoids)
        if sig.dtype not in (torch.float32, torch.float64):
            sig = sig.to(torch.float64)
        sig = sig.reshape(sig.shape[0], -1)[:, 0].contiguous()
        n = sig.numel()
        if n and sync:
            lo, hi = torch.aminmax(sig)
            assert float(lo) >= 0.0                             # idpsampler.py:22
            assert float(hi) <= 1.0                             # idpsampler.py:23
        u = None
        seed, offset = 0, 0
        if uniforms is not None:
            u, _ = R.to_device(uniforms, torch.float64)
            u = u.reshape(-1)
            if u.numel() < n:                                         # the kernel reads one uniform per processed row
                self._pull()                                          # (host read of the state only in this rare case)
                need = n - 1 if (self._d_host is None and n > 0) else n   # no draw for the first move without d_curr
                if u.numel() < need:
                    raise ValueError("uniforms has %d entries, the chain over %d rows needs %d (idpsampler.py:50)"
                                     % (u.numel(), n, need))
        elif self.rng == "numpy":
            # one np.random.uniform(0, 1) per row (idpsampler.py:50); none for the first row when d_curr is None
            self._pull()
            ndraw = n - 1 if (self._d_host is None and n > 0) else n
            u, _ = R.to_device(np.random.uniform(0, 1, size=ndraw) if ndraw else np.zeros(1), torch.float64)
        elif self.rng == "philox":
            seed, offset = self.seed, self.offset
            self.offset += n
        else:
            raise ValueError("rng must be 'numpy' or 'philox'")
        d, kind, cnt = self._push(dev)
        accepted = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        emit = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = self._ws.get(lib.cgs_mh_workspace_bytes(n), dev)
        L.check(lib.cgs_mh_accept(L.ptr(sig), R.score_dtype(sig), n, L.ptr(u), seed, offset, L.ptr(d), L.ptr(kind),
                                  L.ptr(cnt), int(self.thin_period), int(self.burn_in), L.ptr(accepted), L.ptr(emit),
                                  L.ptr(count), L.ptr(ws), ws.numel(), L.stream_ptr()))
        self._last_count = count
        if not sync:
            cap = min(self.emit_capacity(n), emit.numel())
            self.last_accepted, self.last_emit_src = accepted[:n], emit[:cap]
            return self.last_emit_src, count
        k = int(count.item())
        self.last_accepted, self.last_emit_src = accepted[:n], emit[:k]
        return self.last_emit_src

    def gather(self, samples):
        """Rows of ``samples`` emitted by the last ``select`` call, float32, in emission order."""
        dev = R.require_cuda()
        smp, smp_np = R.to_device(samples)
        emit, count = self.last_emit_src, self._last_count
        k = emit.numel()
        if k == 0:
            return R.back(torch.empty((0,), dtype=torch.float32, device=dev), smp_np)
        src = (smp if smp.dtype == torch.float32 else smp.to(torch.float32)).contiguous()
        out = torch.empty((k,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
        L.check(L.load().cgs_gather_rows(L.ptr(src), src[0].numel() * 4, L.ptr(emit), L.ptr(count), k, L.ptr(out),
                                         L.stream_ptr()))
        return R.back(out, smp_np)

    def sampling(self, samples, sigmoids, uniforms=None):
        dev = R.require_cuda()
        lib = L.load()
        smp, smp_np = R.to_device(samples)
        nsig = sigmoids.shape[0]
        assert smp.shape[0] == nsig                              # idpsampler.py:21
        emit = self.select(sigmoids, uniforms)
        k = emit.numel()
        count = self._last_count
        if k == 0:
            out = torch.empty((0,), dtype=torch.float32, device=dev)        # np.asarray([], float32): shape (0,)
            return R.back(out, smp_np)
        src = smp if smp.dtype == torch.float32 else smp.to(torch.float32)  # idpsampler.py:41 -> float32
        out = torch.empty((k,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
        row_bytes = src[0].numel() * 4
        L.check(lib.cgs_gather_rows(L.ptr(src), row_bytes, L.ptr(emit), L.ptr(count), k, L.ptr(out), L.stream_ptr()))
        return R.back(out, smp_np)

    def next(self, d_next):
        """One MH move for a single score (idpsampler.py:43-53); runs the same kernel with n = 1."""
        dev = R.require_cuda()
        s = d_next if isinstance(d_next, torch.Tensor) else np.asarray(d_next)
        sig, _ = R.to_device(s)
        sig = sig.reshape(-1)[:1]
        saved_t, saved_b, saved_cnt = self.thin_period, self.burn_in, self.cnt_chain
        try:
            self.thin_period, self.burn_in = 0, 0
            self.sampling(torch.zeros(1, 1, device=dev), sig)
            moved = bool(self.last_accepted[0].item())
        finally:
            self.thin_period, self.burn_in = saved_t, saved_b
            self.cnt_chain = saved_cnt
        return moved

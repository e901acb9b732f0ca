"""CPU restatement of the reference's accept-reject FILL-UP loops.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED for the loop control flow: the loops live
inside ``GAN.evaluate`` / ``synthetic/main.evaluate`` next to TensorFlow session calls and cannot be executed here;
the restatement follows them statement by statement (citations below).  The samplers plugged into the loops ARE pinned
(``oracle/sampling_np.py`` vs the reference's own ``rejector.py`` / ``idpsampler.py``), and
``tests/test_fillup_cpu.py`` drives these loops with the reference's own sampler classes as well.

Two drivers:
* ``fill_up_nsgan``     nsgan/GAN.py:311-339 (rejection), :348-375 (hastings), :400-427 (collaborate)
* ``fill_up_synthetic`` synthetic/main.py:149-169 (rejection), :181-199 (hastings), :228-247 (collaborate)
"""
from __future__ import annotations

import numpy as np

MIN_EFFICIENCY = 0.2          # nsgan/GAN.py:18


def fill_up_nsgan(base_samples, base_scores, propose, sampling, eval_size, batch_size,
                  min_efficiency=MIN_EFFICIENCY, store_guard="batch", max_batches=None):
    """nsgan/GAN.py:311-339 / :348-375 / :400-427.

    base_samples / base_scores: the ``eval_size`` rows of the base call (:315, :351, :403).
    propose(batch_size) -> (samples, scores) of one fill-up batch (:319-320, :356-357, :408-411; scoring is skipped
    by the reference once the loop back-fills, :409-411, which the caller may mirror by returning scores lazily).
    sampling(samples, scores) -> accepted rows (``Rejector.sampling(..., shift_percent=100.0)`` /
    ``IndependenceSampler.sampling``), shape (0,) or (0, ...) when empty.
    store_guard: 'batch' = ``if cnt_batch > 0`` (:360, :412, hastings / collaborate); 'running' = ``if cnt_reject >
    0`` (:323, the rejection loop tests the RUNNING count, sic: when the base call accepted nothing, accepted rows of
    the first batches are counted but never stored).
    Returns dict(samples [eval_size, ...] (float64 like np.empty; unwritten slots are zero here), cnt, cnt_propose,
    efficiency = cnt / cnt_propose (:344, :380, :432; cnt may overshoot eval_size), n_backfilled, n_batches).
    """
    max_num_propose = eval_size / min_efficiency                         # :283
    base_samples = np.asarray(base_samples)
    out = np.zeros((eval_size,) + tuple(base_samples.shape[1:]))         # np.empty in the reference
    cnt_propose = eval_size                                              # :312, :349, :401
    base = np.asarray(sampling(base_samples, base_scores))               # :315, :351, :403
    cnt = base.shape[0]                                                  # :316
    if cnt > 0:                                                          # :317-318
        out[:cnt] = base
    backfilled = 0
    batches = 0
    while cnt < eval_size:                                               # :320
        batch_samples, batch_scores = propose(batch_size)
        batch_samples = np.asarray(batch_samples)
        if cnt_propose < max_num_propose:                                # :322
            acc = np.asarray(sampling(batch_samples, batch_scores() if callable(batch_scores) else batch_scores))
            cnt_batch = acc.shape[0]
            guard = (cnt > 0) if store_guard == "running" else (cnt_batch > 0)
            if guard:
                if cnt + cnt_batch < eval_size:                          # :326-329
                    out[cnt:cnt + cnt_batch] = acc
                else:
                    out[cnt:eval_size] = acc[:eval_size - cnt]
            cnt = cnt + cnt_batch                                        # :330
        else:                                                            # :331-337 "Oops, too inefficient"
            if cnt + batch_size < eval_size:
                out[cnt:cnt + batch_size] = batch_samples
            else:
                out[cnt:eval_size] = batch_samples[:eval_size - cnt]
            cnt = cnt + batch_size
            backfilled += batch_size
        cnt_propose = cnt_propose + batch_size                           # :338
        batches += 1
        if max_batches is not None and batches >= max_batches:
            break
    return dict(samples=out, cnt=cnt, cnt_propose=cnt_propose, efficiency=cnt / cnt_propose,
                n_backfilled=backfilled, n_batches=batches)


def fill_up_synthetic(base_samples, base_scores, propose, sampling, max_batches=None):
    """synthetic/main.py:149-169 / :181-199 / :228-247: batches of ``eval_size`` rows, NO efficiency guard, and
    ``cnt_propose`` only advances when a batch accepted something (sic, :163-169)."""
    base_samples = np.asarray(base_samples)
    eval_size = base_samples.shape[0]                                    # :133
    cnt_propose = eval_size                                              # :150
    out = np.zeros_like(base_samples)                                    # np.empty_like, :151
    base = np.asarray(sampling(base_samples, base_scores))               # :153
    cnt = base.shape[0]
    if cnt > 0:
        out[:cnt] = base
    batches = 0
    while cnt < eval_size:                                               # :157
        extra, scores = propose(eval_size)                               # :158-160
        acc = np.asarray(sampling(np.asarray(extra), scores() if callable(scores) else scores))   # :161
        cnt_extra = acc.shape[0]
        if cnt_extra > 0:                                                # :163
            if cnt + cnt_extra < eval_size:
                out[cnt:cnt + cnt_extra] = acc
            else:
                out[cnt:eval_size] = acc[:eval_size - cnt]
            cnt = cnt + cnt_extra
            cnt_propose = cnt_propose + eval_size                        # :169 (inside the if, sic)
        batches += 1
        if max_batches is not None and batches >= max_batches:           # the reference would spin forever
            break
    return dict(samples=out, cnt=cnt, cnt_propose=cnt_propose, efficiency=cnt / cnt_propose, n_backfilled=0,
                n_batches=batches)


# --------------------------------------------------------------------------------------------------------------
# samplers with the reference's interfaces on top of the pinned oracle kernels (numpy global RNG, like the reference)
# --------------------------------------------------------------------------------------------------------------
class OracleRejector:
    """rejector.py:7-38 through ``sampling_np.drs_accept``; one ``np.random.rand(N)`` per call (:33)."""

    def __init__(self):
        self.D_tilde_M = 0.0

    def set_score_max(self, score_max):
        from . import sampling_np as snp
        self.D_tilde_M = snp.drs_score_max(score_max)

    def sampling(self, samples, sigmoids, epsilon=1e-8, shift_percent=60.0):
        from . import sampling_np as snp
        u = np.random.rand(len(samples))
        acc, self.D_tilde_M = snp.drs_accept(sigmoids, u, self.D_tilde_M, epsilon, shift_percent)
        return np.asarray(samples)[acc]


class OracleIndependenceSampler:
    """idpsampler.py:4-53 through ``sampling_np.mh_chain``; one ``np.random.uniform`` per row (:50)."""

    def __init__(self, T=5, B=0):
        self.d_curr, self.cnt_chain, self.thin_period, self.burn_in = None, 1, T, B

    def set_score_curr(self, d_curr):
        self.d_curr = d_curr

    def sampling(self, samples, sigmoids):
        from . import sampling_np as snp
        n = len(samples)
        ndraw = n - 1 if (self.d_curr is None and n > 0) else n
        u = np.random.uniform(0, 1, size=ndraw)
        if ndraw < n:
            u = np.concatenate([[0.0], u])
        emit, self.d_curr, self.cnt_chain, _ = snp.mh_chain(sigmoids, u, self.d_curr, self.cnt_chain,
                                                            self.thin_period, self.burn_in)
        return np.asarray(np.asarray(samples)[emit], dtype=np.float32) if len(emit) else np.asarray([], np.float32)

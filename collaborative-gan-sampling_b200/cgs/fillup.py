"""Accept-reject fill-up drivers: the loops that turn refined batches into ``eval_size`` accepted
samples, mirroring the reference statement by statement:

* ``fill_up``            nsgan/GAN.py:311-339 (rejection), :348-375 (hastings), :400-427 (collaborate): a base call over
  ``eval_size`` rows, then ``batch_size`` rows at a time; once ``cnt_propose`` reaches ``eval_size / MIN_EFFICIENCY``
  (nsgan/GAN.py:18,283,322) the remaining slots are back-filled with UN-FILTERED batches (:331-337); the efficiency
  is ``cnt / cnt_propose`` with ``cnt`` allowed to overshoot (:344);
* ``fill_up_synthetic``  synthetic/main.py:149-169 / :181-199 / :228-247: ``eval_size`` rows per batch, no efficiency
  guard, ``cnt_propose`` advances only when a batch accepted something (sic).

Everything between proposal and acceptance stays on the device; per batch the host learns only the accepted-row count
(the loop condition needs it).  The samplers are the drop-in ``Rejector`` / ``IndependenceSampler``.
"""
from __future__ import annotations

from collections import namedtuple

import torch

MIN_EFFICIENCY = 0.2          # nsgan/GAN.py:18

FillUpResult = namedtuple("FillUpResult", "samples cnt cnt_propose efficiency n_backfilled n_batches")


def _rows(good, like):
    good = good if isinstance(good, torch.Tensor) else torch.as_tensor(good)
    if good.dim() == 1 and good.numel() == 0:          # the reference's empty MH result has shape (0,)
        return like[:0]
    return good.to(like.device)


def fill_up(base_samples, base_scores, propose, sampling, eval_size, batch_size, min_efficiency=MIN_EFFICIENCY,
            store_guard="batch", max_batches=None):
    """Collect ``eval_size`` samples like nsgan/GAN.py:311-339.

    base_samples / base_scores: the ``eval_size`` rows of the base call (:315 / :351 / :403 -- the "collaborate"
    variant of the reference passes the UN-refined standard samples here; that is the caller's choice).
    propose(batch_size) -> (samples, scores) of one fill-up batch; ``scores`` may be a callable, evaluated only while
    the loop still filters (the reference skips the scoring pass once it back-fills, :409-411).
    sampling(samples, scores) -> accepted rows (``Rejector.sampling`` / ``IndependenceSampler.sampling`` or a partial).
    store_guard: 'batch' (``if cnt_batch > 0``, :360/:412) or 'running' (``if cnt_reject > 0``, :323, sic).
    """
    if store_guard not in ("batch", "running"):
        raise ValueError("store_guard must be 'batch' or 'running'")
    max_num_propose = eval_size / min_efficiency                         # nsgan/GAN.py:283
    base_samples = base_samples if isinstance(base_samples, torch.Tensor) else torch.as_tensor(base_samples)
    out = torch.zeros((eval_size,) + tuple(base_samples.shape[1:]), dtype=base_samples.dtype, device=base_samples.device)
    cnt_propose = eval_size
    base = _rows(sampling(base_samples, base_scores), base_samples)
    cnt = int(base.shape[0])
    if cnt > 0:
        out[:cnt] = base.to(out.dtype)
    backfilled, batches = 0, 0
    while cnt < eval_size:
        batch_samples, batch_scores = propose(batch_size)
        if cnt_propose < max_num_propose:
            acc = _rows(sampling(batch_samples, batch_scores() if callable(batch_scores) else batch_scores), out)
            cnt_batch = int(acc.shape[0])
            if (cnt > 0) if store_guard == "running" else (cnt_batch > 0):
                take = cnt_batch if cnt + cnt_batch < eval_size else eval_size - cnt
                out[cnt:cnt + take] = acc[:take].to(out.dtype)
            cnt += cnt_batch
        else:                                                            # "Oops, too inefficient" :331-337
            take = batch_size if cnt + batch_size < eval_size else eval_size - cnt
            out[cnt:cnt + take] = batch_samples[:take].to(out.device, out.dtype)
            cnt += batch_size
            backfilled += batch_size
        cnt_propose += batch_size
        batches += 1
        if max_batches is not None and batches >= max_batches:
            break
    return FillUpResult(out, cnt, cnt_propose, cnt / cnt_propose, backfilled, batches)


def fill_up_synthetic(base_samples, base_scores, propose, sampling, max_batches=None):
    """synthetic/main.py:149-169: ``eval_size`` rows per batch, no efficiency guard (``max_batches`` bounds the loop
    the reference would spin in when nothing is ever accepted)."""
    base_samples = base_samples if isinstance(base_samples, torch.Tensor) else torch.as_tensor(base_samples)
    eval_size = int(base_samples.shape[0])
    out = torch.zeros_like(base_samples)
    cnt_propose = eval_size
    base = _rows(sampling(base_samples, base_scores), base_samples)
    cnt = int(base.shape[0])
    if cnt > 0:
        out[:cnt] = base.to(out.dtype)
    batches = 0
    while cnt < eval_size:
        extra, scores = propose(eval_size)
        acc = _rows(sampling(extra, scores() if callable(scores) else scores), out)
        cnt_extra = int(acc.shape[0])
        if cnt_extra > 0:
            take = cnt_extra if cnt + cnt_extra < eval_size else eval_size - cnt
            out[cnt:cnt + take] = acc[:take].to(out.dtype)
            cnt += cnt_extra
            cnt_propose += eval_size                                     # inside the if (sic, :169)
        batches += 1
        if max_batches is not None and batches >= max_batches:
            break
    return FillUpResult(out, cnt, cnt_propose, cnt / cnt_propose, 0, batches)

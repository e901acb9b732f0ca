"""Accept-reject fill-up driver: propose -> [refine] -> score -> accept until ``eval_size`` samples.

Mirrors the reference's evaluation loops (nsgan/GAN.py:311-433, synthetic/main.py:149-263): a base call over the
first ``eval_size`` proposals, then batch-by-batch fill-up; once ``cnt_propose`` exceeds ``eval_size / MIN_EFFICIENCY``
(nsgan/GAN.py:18,283) the remaining slots are back-filled with un-filtered batches.  Everything between proposal and
acceptance stays on the device; the only host traffic per batch is the accepted-row count.
"""
from __future__ import annotations

import torch

MIN_EFFICIENCY = 0.2          # nsgan/GAN.py:18


def fill_up(propose, score, sampler, eval_size, batch_size, refine=None, min_efficiency=MIN_EFFICIENCY,
            max_batches=None):
    """Collect ``eval_size`` accepted samples.

    propose(n) -> proposals [n, ...] (e.g. ProposalHead(z)); refine(x) -> refined samples (optional, e.g.
    ``Refiner.build_refiner``); score(x) -> sigmoid scores [n] or [n,1]; sampler = Rejector / IndependenceSampler
    drop-in (``sampling(samples, scores)`` returning the accepted rows).
    Returns (samples [eval_size, ...], efficiency = accepted / proposed, n_backfilled).
    """
    out, have, proposed, backfilled, batches = None, 0, 0, 0, 0
    max_propose = eval_size / min_efficiency                      # nsgan/GAN.py:283
    first = True
    while have < eval_size:
        n = eval_size if first else batch_size                    # base call over eval_size rows, then per batch
        first = False
        x = propose(n)
        if refine is not None:
            x = refine(x)
        if proposed < max_propose or have == 0 and proposed == 0:
            good = sampler.sampling(x, score(x))
            good = good if isinstance(good, torch.Tensor) else torch.as_tensor(good)
            if good.dim() == 1 and good.numel() == 0:              # the reference's empty result has shape (0,)
                good = x[:0]
        else:                                                      # "Oops, too inefficient": nsgan/GAN.py:332-338
            good = x
            backfilled += good.shape[0]
        proposed += n
        batches += 1
        if good.shape[0]:
            if out is None:
                out = torch.empty((eval_size,) + tuple(good.shape[1:]), dtype=good.dtype, device=good.device)
            take = min(good.shape[0], eval_size - have)
            out[have:have + take] = good[:take].to(out.dtype)
            have += take
        if max_batches is not None and batches >= max_batches:
            break
    return (out[:have] if out is not None else None), have / max(proposed, 1), backfilled

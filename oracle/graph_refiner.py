"""FP32 torch-CPU restatement of ``sampling/collaborator.py`` (the TF-graph refiner).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no TensorFlow here, the reference ships no
vectors for this path) -- the restatement follows the reference line by line and its
gradients come from ``torch.autograd`` exactly where the reference calls ``tf.gradients``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import nets


def loss_refine(logits):
    """nsgan/GAN.py:176-177: un-reduced sigmoid_cross_entropy_with_logits(labels=1) = softplus(-l)."""
    return F.softplus(-logits)


def forward_logits_and_grad(h, arch, w, d_bn="inference"):
    """collaborator.py:26-39.  Returns (per-sample logit mean [B], d sum(loss)/d h, image)."""
    h = h.detach().clone().requires_grad_(True)
    x = nets.feature_to_data(h, arch, w)                       # :28
    logits = nets.discriminator(x, arch, w, d_bn)              # :29
    loss = loss_refine(logits)                                 # :30
    (g,) = torch.autograd.grad(loss.sum(), h)                  # :31 (tf.gradients sums ys)
    logit_mean = logits.reshape(logits.shape[0], -1).mean(dim=1)   # :34-37
    return logit_mean.detach(), g.detach(), x.detach()


def build_refiner(h0, arch, w, steps, rate, method="momentum", mode="deterministic",
                  d_bn="inference", prob_indices=None, vmin=None, vmax=None, alpha=0.9,
                  return_trace=False, exit_logit=None):
    """collaborator.py:41-88 evaluated eagerly on a concrete batch.

    h0: [B,H,W,C] float32 NHWC.  Returns dict(refined [B,h,w,c], optimal_logit, optimal_step,
    default_logit, optimal_feature[, trace]).  ``prob_indices`` replaces the build-time
    ``np.random.randint(K+1, size=B)`` of :54-56 for mode='probabilistic'.
    Policy: policy.py:26-37 (sgd / momentum; 'ladam' crashes in the reference here because
    collaborator.py:66 passes no loss -- policy.py:51 -- so it is rejected).
    """
    if method not in ("sgd", "momentum"):
        raise NotImplementedError("graph refiner supports sgd/momentum only (policy.py:51 with loss=None)")
    h0 = torch.as_tensor(h0, dtype=torch.float32)
    cur = h0.clone()                                            # :48
    cur_logit, grad, img = forward_logits_and_grad(cur, arch, w, d_bn)   # :49
    default_logit = cur_logit.clone()                           # :52
    if mode == "probabilistic":
        assert prob_indices is not None
        idx = torch.as_tensor(np.asarray(prob_indices))
    best_feat = h0.clone()                                      # :58
    best_logit = cur_logit.clone()                              # :59
    best_step = torch.ones_like(best_logit)                     # :60  (sic: starts at 1)
    best_img = img.clone()
    momentum = None
    trace = []
    # opt-in early exit (README.md:13; NOT in the reference code, which always runs K steps): a sample whose logit
    # reaches exit_logit keeps its best state and its feature as of that evaluation
    done = (cur_logit >= exit_logit) if exit_logit is not None else torch.zeros_like(cur_logit, dtype=torch.bool)
    final_feat = cur.clone()
    for i in range(steps):                                      # :63
        if method == "sgd":                                     # policy.py:27-29
            cur = cur - rate * grad
        else:                                                   # policy.py:31-37
            momentum = rate * grad if momentum is None else alpha * momentum + rate * grad
            cur = cur - momentum
        if vmin and vmax:                                       # :69 (truthiness test, sic)
            cur = cur.clamp(min=vmin, max=vmax)
        cur_logit, grad, img = forward_logits_and_grad(cur, arch, w, d_bn)   # :73
        if mode == "probabilistic":
            upd = idx == i                                      # :77
        elif mode == "deterministic":
            upd = cur_logit > best_logit                        # :79
        else:
            raise NotImplementedError(mode)
        upd = upd & ~done
        if exit_logit is not None:
            newly = (~done) & (cur_logit >= exit_logit)
            final_feat = torch.where((~done).view(-1, *([1] * (cur.dim() - 1))), cur, final_feat)
            done = done | newly
        best_logit = torch.where(upd, cur_logit, best_logit)    # :81
        m = upd.view(-1, *([1] * (cur.dim() - 1)))
        best_feat = torch.where(m, cur, best_feat)              # :82
        best_img = torch.where(m, img, best_img)
        best_step = torch.where(upd, torch.full_like(best_step, i + 1), best_step)   # :83
        if return_trace:
            trace.append(dict(logit=cur_logit.clone(), feature=cur.clone()))
    refined = nets.feature_to_data(best_feat, arch, w).detach()   # :88
    out = dict(refined=refined, optimal_logit=best_logit, optimal_step=best_step,
               default_logit=default_logit, optimal_feature=best_feat, best_img_kept=best_img,
               final_feature=cur if exit_logit is None else final_feat, done=done)
    if return_trace:
        out["trace"] = trace
    return out

"""Functional numpy restatement of the reference's host-side sampling algorithms.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: every function here is
checked bit-for-bit against the reference's own module run in the build container
(``tests/test_oracle_vs_reference.py``) and against ``tests/golden/*.npz``.

Unlike the reference classes these are pure functions: random uniforms and carried state
are explicit arguments / return values, which is what the C ABI of the CUDA build takes.
"""
from __future__ import annotations

import numpy as np

CLIP_LO = 1e-14            # sampling/rejector.py:12,18
CLIP_HI = 1 - 1e-14


# --------------------------------------------------------------------------------------
# update policy  (sampling/policy.py:26-64), numpy branch
# --------------------------------------------------------------------------------------

def policy_new_state():
    return {"momentum": None, "mean_square": None, "loss": None}


def policy_step(method, theta, grad, state, step_size, loss=None,
                alpha=0.9, beta1=0.9, beta2=0.5, beta3=0.5, degree=2, eps=1e-8):
    """One in-place update of ``theta`` ([N,2] or any shape for sgd/momentum); returns theta.

    sgd       policy.py:27-29   theta -= lambda g
    momentum  policy.py:31-37   m = lambda g (first) | alpha m + lambda g ; theta -= m
    ladam     policy.py:39-51,61  m = g | b1 m + (1-b1) g ; v = g^2 | b2 v + (1-b2) g^2 ;
                                l = loss | b3 l + (1-b3) loss ;
                                theta -= lambda m / (sqrt(v)+eps) * clip(l+.5, 0, inf)^degree
    (numpy branch only: the TF branch at policy.py:52-59 additionally clips at 1e4.)
    """
    if method == "sgd":
        theta -= step_size * grad
        return theta
    if method == "momentum":
        if state["momentum"] is None:
            state["momentum"] = step_size * grad
        else:
            state["momentum"] = alpha * state["momentum"] + step_size * grad
        theta -= state["momentum"]
        return theta
    if method == "ladam":
        g = grad
        state["momentum"] = g if state["momentum"] is None else beta1 * state["momentum"] + (1.0 - beta1) * g
        state["mean_square"] = g ** 2 if state["mean_square"] is None else \
            beta2 * state["mean_square"] + (1.0 - beta2) * g ** 2
        state["loss"] = loss if state["loss"] is None else beta3 * state["loss"] + (1.0 - beta3) * loss
        rescale = np.expand_dims((state["loss"] + 0.5).clip(min=0.0), axis=1) ** degree
        theta -= step_size * state["momentum"] / (np.sqrt(state["mean_square"]) + eps) * rescale
        return theta
    raise NotImplementedError(method)


# --------------------------------------------------------------------------------------
# Discriminator Rejection Sampling  (sampling/rejector.py:11-38)
# --------------------------------------------------------------------------------------

def _logit(p):
    return np.log(p / (1.0 - p))        # scipy.special.logit == log(p/(1-p)) in float64


def _expit(x):
    from scipy.special import expit
    return expit(x)


def drs_score_max(score_max):
    """rejector.py:11-14 -> D_tilde_M (float64)."""
    s = np.clip(np.asarray(score_max).astype(np.float64), CLIP_LO, CLIP_HI)
    from scipy.special import logit
    return logit(s)


def drs_probabilities(sigmoids, d_tilde_m, epsilon=1e-8, shift_percent=60.0):
    """rejector.py:18-31 -> (P [N] float64, new D_tilde_M)."""
    from scipy.special import logit
    s = np.clip(np.asarray(sigmoids).astype(np.float64), CLIP_LO, CLIP_HI)
    d_tilde = logit(s)
    m_new = np.maximum(d_tilde_m, np.amax(d_tilde))
    delta = d_tilde - m_new
    f = delta - np.log(1 - np.exp(delta - epsilon))
    if shift_percent is not None:
        f = f - np.percentile(f, shift_percent)
    return np.squeeze(_expit(f)), m_new


def drs_accept(sigmoids, uniforms, d_tilde_m, epsilon=1e-8, shift_percent=60.0):
    """rejector.py:33 with explicit uniforms -> (accept mask [N] bool, new D_tilde_M).

    ``uniforms`` must be what ``np.random.rand(N)`` would have returned."""
    p, m_new = drs_probabilities(sigmoids, d_tilde_m, epsilon, shift_percent)
    return np.asarray(uniforms) < p, m_new


# --------------------------------------------------------------------------------------
# MH-GAN independence sampler  (sampling/idpsampler.py:17-53)
# --------------------------------------------------------------------------------------

def mh_chain(sigmoids, uniforms, d_curr, cnt_chain, thin_period, burn_in=0):
    """Sequential chain with explicit uniforms.

    Returns (emit_src [n_emit] int64: for every emitted sample the row index of ``samples`` it
    copies, new d_curr, new cnt_chain, accepted mask [N]).  idpsampler.py:27-39 (loop), :43-53
    (``next``): alpha = min(1, d'(1-d)/(d(1-d'))) evaluated in the dtype the scores arrive in
    (numpy-2 promotion); reject iff u > alpha; one uniform per row.
    ``min(1.0, x)`` is Python's: it returns x only if ``x < 1.0`` (NaN -> 1.0).
    """
    sig = np.asarray(sigmoids)
    sig = sig.reshape(sig.shape[0], -1)[:, 0]
    n = sig.shape[0]
    u = np.asarray(uniforms, dtype=np.float64)
    emit = []
    accepted = np.zeros(n, dtype=bool)
    curr = -1
    cnt_good = 0
    one = 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(n):
            d_next = sig[i]
            move = True
            if d_curr is not None:
                ratio = d_next * (one - d_curr) / (d_curr * (one - d_next))
                alpha = ratio if ratio < 1.0 else 1.0
                if u[i] > alpha:
                    move = False
            if move:
                d_curr = d_next
                accepted[i] = True
                cnt_good += 1
                if cnt_good > burn_in:
                    curr = i
            if curr >= 0:
                if cnt_chain > thin_period:
                    emit.append(curr)
                    cnt_chain = 1
                else:
                    cnt_chain += 1
    return np.asarray(emit, dtype=np.int64), d_curr, cnt_chain, accepted


# --------------------------------------------------------------------------------------
# 2-D host-loop refiner  (sampling/refiner_cpu.py:19-81)
# --------------------------------------------------------------------------------------

def refine_2d(fake_batch, score_fn, real_sigmoid_mean, steps, step_size, method="ladam",
              prob_indices=None):
    """Data-space refinement.  ``score_fn(x) -> (sigmoid [N,1], saliency [N,2])`` (FP32).

    refiner_cpu.py:26-28 init, :46-66 loop (policy step -> score -> strict-improvement select),
    :58 ``optimal_loss - forward_loss > 0``, :33 optimal_step starts at 0.
    Returns dict(optimal_batch, optimal_loss, optimal_step, traj [N,steps+1,2] float64,
    loss_traj [N,steps+1] float64[, probabilistic [N,2] float64 if prob_indices is given]).
    ``real_sigmoid_mean`` = np.mean(real_sigmoid) computed by the caller (refiner_cpu.py:22-23,28).
    """
    fake_batch = np.asarray(fake_batch)
    x = fake_batch.copy()
    sig, grad = score_fn(x)
    loss = real_sigmoid_mean - np.squeeze(sig)
    best_x = x.copy()
    best_loss = loss.copy()
    best_step = np.zeros_like(best_loss)
    n = len(fake_batch)
    traj = np.zeros((n, steps + 1, 2))
    ltraj = np.zeros((n, steps + 1))
    traj[:, 0, :] = fake_batch
    ltraj[:, 0] = loss
    state = policy_new_state()
    for i in range(steps):
        policy_step(method, x, grad, state, step_size, loss)
        sig, grad = score_fn(x)
        loss = real_sigmoid_mean - np.squeeze(sig)
        upd = (best_loss - loss) > 0
        best_loss[upd] = loss[upd]
        best_x[upd, :] = x[upd, :]
        best_step[upd] = i + 1
        traj[:, i + 1, :] = x
        ltraj[:, i + 1] = loss
    out = dict(optimal_batch=best_x, optimal_loss=best_loss, optimal_step=best_step,
               traj=traj, loss_traj=ltraj)
    if prob_indices is not None:
        idx = np.asarray(prob_indices)
        out["probabilistic"] = traj[np.arange(n), idx, :]      # refiner_cpu.py:71-76 (float64)
    return out

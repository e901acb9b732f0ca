"""Metrics of BASELINE.md §5 for the image path at K=50 (shared by tests/test_k50_parity_gpu.py and tools/k50_parity.py).

TEST INFRASTRUCTURE: imports ``oracle`` (the CPU restatement); never imported by the product package.

* image max-abs / rel-L2, |delta optimal_logit|, ``optimal_step`` agreement rate;
* sign agreement of (logit - threshold): the threshold the README describes (README.md:13) is "D classifies the
  sample as real"; two thresholds are reported: logit 0 (sigmoid 0.5) and the batch median of the oracle's logits
  (the reference's unused ``real_logits_mean``, collaborator.py:44-45, is a batch statistic of that kind);
* Jaccard index of the MH-GAN emitted / accepted row sets (idpsampler.py:27-53) under the same uniforms;
* Frechet distance between the two refined sets in D's penultimate feature space (FID itself needs Inception
  weights that are not reachable here).  With n samples of dimension d >> n the covariances are
  rank-deficient, so Tr((C1 C2)^(1/2)) is evaluated exactly as the nuclear norm of A B^T, A = X1c / sqrt(n-1),
  B = X2c / sqrt(n-1) (the non-zero eigenvalues of C1 C2 = A^T A B^T B equal those of (A B^T)(A B^T)^T).
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import nets as onets
from oracle import sampling_np as snp


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def frechet_distance(x1, x2):
    """||mu1 - mu2||^2 + Tr(C1 + C2 - 2 (C1 C2)^(1/2)) for two [n, d] feature sets (float64)."""
    x1 = np.asarray(x1, np.float64).reshape(len(x1), -1)
    x2 = np.asarray(x2, np.float64).reshape(len(x2), -1)
    mu1, mu2 = x1.mean(0), x2.mean(0)
    a = (x1 - mu1) / np.sqrt(max(len(x1) - 1, 1))
    b = (x2 - mu2) / np.sqrt(max(len(x2) - 1, 1))
    tr1, tr2 = float((a * a).sum()), float((b * b).sum())
    cross = float(np.linalg.svd(a @ b.T, compute_uv=False).sum())
    return float(((mu1 - mu2) ** 2).sum() + tr1 + tr2 - 2.0 * cross)


def d_features(images, arch, w):
    """Penultimate activation of the ORACLE's discriminator (input of the 1-logit head), [n, d]."""
    col = []
    x = torch.as_tensor(np.asarray(images, np.float32))
    with torch.no_grad():
        onets.discriminator(x, arch, w, "inference", collect=col)
    f = col[-2]
    return f.reshape(f.shape[0], -1).numpy()


def jaccard(a, b):
    a, b = set(np.asarray(a).tolist()), set(np.asarray(b).tolist())
    return 1.0 if not (a or b) else len(a & b) / float(len(a | b))


def k50_metrics(ref, got, arch, w, seed=2019):
    """ref / got: dict(refined [B,h,w,c], optimal_logit [B], optimal_step [B], default_logit [B]) as numpy arrays."""
    B = len(ref["optimal_logit"])
    lo, lg = np.asarray(ref["optimal_logit"], np.float64), np.asarray(got["optimal_logit"], np.float64)
    m = {
        "B": int(B),
        "img_max_abs": float(np.abs(np.asarray(got["refined"], np.float64) - ref["refined"]).max()),
        "img_rel_l2": rel_l2(got["refined"], ref["refined"]),
        "optimal_logit_max_abs": float(np.abs(lg - lo).max()),
        "optimal_logit_mean_abs": float(np.abs(lg - lo).mean()),
        "default_logit_max_abs": float(np.abs(np.asarray(got["default_logit"], np.float64) - ref["default_logit"]).max()),
        "optimal_step_agree": float((np.asarray(got["optimal_step"]) == np.asarray(ref["optimal_step"])).mean()),
        "optimal_step_within_1": float((np.abs(np.asarray(got["optimal_step"], np.float64) - ref["optimal_step"]) <= 1).mean()),
        "sign_agree_logit0": float(((lg > 0) == (lo > 0)).mean()),
        "sign_agree_median": float(((lg > np.median(lo)) == (lo > np.median(lo))).mean()),
        "logit_gain_mean_ref": float((lo - np.asarray(ref["default_logit"], np.float64)).mean()),
    }
    u = np.random.RandomState(seed).rand(B)
    so = (1.0 / (1.0 + np.exp(-lo))).astype(np.float32).reshape(-1, 1)
    sg = (1.0 / (1.0 + np.exp(-lg))).astype(np.float32).reshape(-1, 1)
    for T in (0, 20):
        eo, _, _, ao = snp.mh_chain(so, u, np.float32(0.5), 1, T, 0)
        eg, _, _, ag = snp.mh_chain(sg, u, np.float32(0.5), 1, T, 0)
        m["mh_T%d_emit_jaccard" % T] = jaccard(eo, eg)
        if T == 0:
            m["mh_accept_mask_agree"] = float((ao == ag).mean())
    m["frechet_d_feature"] = frechet_distance(d_features(ref["refined"], arch, w), d_features(got["refined"], arch, w))
    # scale of the feature space, so the absolute distance can be read: Frechet distance between the oracle's refined
    # set and the un-refined proposals' images would be the natural yardstick; the trace of the covariance is cheap
    f = d_features(ref["refined"], arch, w)
    m["d_feature_trace_cov"] = float(((f - f.mean(0)) ** 2).sum() / max(len(f) - 1, 1))
    return m

"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

from oracle import nw_oracle as O


def clustered_features(n_classes, per_class, d, n_query, seed, spread=1.0, mean_scale=0.6):
    """ReLU-like, class-clustered, class-sorted synthetic features."""
    rng = np.random.default_rng(seed)
    y = np.repeat(np.arange(n_classes), per_class)
    mu = rng.normal(size=(n_classes, d)) * mean_scale
    s = np.maximum(mu[y] + rng.normal(size=(len(y), d)) * spread + 0.5, 0).astype(np.float32)
    qy = rng.integers(0, n_classes, n_query)
    q = np.maximum(mu[qy] + rng.normal(size=(n_query, d)) * spread + 0.5, 0).astype(np.float32)
    return q, s, y.astype(np.int64), qy


def probs_of(logp):
    return np.exp(np.asarray(logp, dtype=np.float64))


def assert_head_parity(logp_gpu, logp_ref, prob_tol=1e-3, top1=0.999):
    """north_star tolerances: class probabilities within 1e-3 max-abs, top-1 agreement >= 99.9 %."""
    lg = logp_gpu.detach().cpu().numpy() if torch.is_tensor(logp_gpu) else np.asarray(logp_gpu)
    lr = np.asarray(logp_ref)
    assert lg.shape == lr.shape
    assert np.isfinite(lg).all()
    perr = np.abs(probs_of(lg) - probs_of(lr)).max()
    agree = (lg.argmax(1) == lr.argmax(1)).mean()
    assert perr < prob_tol, f"class-probability max-abs error {perr:.3e} >= {prob_tol}"
    assert agree >= top1, f"top-1 agreement {agree:.4f} < {top1}"
    return perr, agree


def oracle_logp(q, s, y, C, kind):
    return O.nw_forward(q, s, y, C, kind)


def fp64_class_probs(q, feats, labels, n_classes, chunk=16384):
    """CHECKER for full BASELINE sizes (the numpy oracle does not fit them): float64 restatement of
    nwhead/nw.py:266-289 + nwhead/kernel.py:13-15 with torch on the GPU, batched over the queries and chunked over
    the bank with an online (running max) class aggregation.  |q|^2 + |s|^2 - 2 q.s in float64 carries ~1e-13 of
    cancellation error at these norms — far below the fp32 reference's own rounding."""
    qd = q.double()
    qn = (qd * qd).sum(1, keepdim=True)
    m = torch.full((q.shape[0], 1), float("-inf"), dtype=torch.float64, device=q.device)
    w = torch.zeros((q.shape[0], n_classes), dtype=torch.float64, device=q.device)
    for i in range(0, feats.shape[0], chunk):
        s = feats[i:i + chunk].double()
        d2 = qn + (s * s).sum(1)[None, :] - 2.0 * (qd @ s.t())
        sc = -d2.clamp_min_(0).sqrt_()
        m_new = torch.maximum(m, sc.max(dim=1, keepdim=True).values)
        w *= (m - m_new).exp()
        w.index_add_(1, labels[i:i + chunk], (sc - m_new).exp_())
        m = m_new
    return w / w.sum(1, keepdim=True)

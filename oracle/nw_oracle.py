"""Float64 numpy restatement of the reference's NW-head hot path.  TEST INFRASTRUCTURE ONLY.

Parity status: PINNED against the unmodified reference by ``oracle/gen_golden.py`` (run in the
build container, where /root/reference is importable) and by ``tests/test_oracle_vs_golden.py``
(fixtures committed under ``tests/golden/``).

Every function cites the reference lines it restates (paths relative to /root/reference).
The restatement is written in float64 so it can serve as the "true" answer for both the fp32
reference and the bf16-input / fp32-accumulate CUDA kernels.
"""
from __future__ import annotations

import numpy as np

LOG_EPS = 1e-12  # nwhead/nw.py:289  torch.log(output + 1e-12)
NORM_EPS = 1e-12  # torch.nn.functional.normalize default eps (nwhead/kernel.py:19-20, 25-26, 41-42)

KERNEL_KINDS = ("euclidean", "hypersphere_euclidean", "cosine", "dotproduct", "clip")
CLIP_LOGIT_SCALE_INIT = float(np.log(1.0 / 0.07))  # nwhead/kernel.py:38


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _f64(a):
    return np.asarray(a, dtype=np.float64)


def l2_normalize(x, eps=NORM_EPS):
    """F.normalize(x, dim=-1): x / max(||x||_2, eps)   (nwhead/kernel.py:19-20, 25-26)."""
    x = _f64(x)
    n = np.sqrt((x * x).sum(-1, keepdims=True))
    return x / np.maximum(n, eps)


def quantize_bf16(x):
    """Round-to-nearest-even fp32 -> bf16 -> fp32, bit exact with the CUDA __float2bfloat16_rn
    used by the bank-build / query-prep kernels.  Used to model the kernel's input rounding."""
    x32 = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    u = x32.view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    out = rounded.astype(np.uint32).view(np.float32).reshape(x32.shape)
    # NaN/Inf pass through unchanged in the exponent; keep them as is
    bad = ~np.isfinite(x32)
    if bad.any():
        out = out.copy()
        out[bad] = x32[bad]
    return out


# --------------------------------------------------------------------------------------
# a2/a3: similarity kernels  (nwhead/kernel.py)
# --------------------------------------------------------------------------------------
def pairwise_scores(q, s, kind="euclidean", logit_scale=CLIP_LOGIT_SCALE_INIT):
    """scores[b, j] for query rows q (B,d) against supports s (N,d) or (B,N,d).

    euclidean             -> -||q - s||_2            nwhead/kernel.py:13-15  (-torch.cdist, un-squared)
    hypersphere_euclidean -> -||q^ - s^||_2          nwhead/kernel.py:17-21
    cosine                -> q^ . s^                 nwhead/kernel.py:23-28
    dotproduct            -> q . s                   nwhead/kernel.py:30-33
    clip                  -> exp(logit_scale) q^.s^  nwhead/kernel.py:35-44
    """
    q = _f64(q)
    s = _f64(s)
    if kind not in KERNEL_KINDS:
        raise NotImplementedError(kind)  # nwhead/kernel.py:95-96
    if kind in ("hypersphere_euclidean", "cosine", "clip"):
        q = l2_normalize(q)
        s = l2_normalize(s)
    if s.ndim == 2:
        dots = q @ s.T
        if kind in ("euclidean", "hypersphere_euclidean"):
            # exact differences in float64, chunked to bound memory
            out = np.empty((q.shape[0], s.shape[0]), dtype=np.float64)
            for b in range(q.shape[0]):
                diff = s - q[b]
                out[b] = -np.sqrt((diff * diff).sum(-1))
            return out
    else:
        dots = np.einsum("bd,bnd->bn", q, s)
        if kind in ("euclidean", "hypersphere_euclidean"):
            diff = s - q[:, None, :]
            return -np.sqrt((diff * diff).sum(-1))
    if kind == "clip":
        return np.exp(logit_scale) * dots
    return dots


# --------------------------------------------------------------------------------------
# a1: NWHead.forward  (nwhead/nw.py:266-289)
# --------------------------------------------------------------------------------------
def class_lse(scores, labels, n_classes):
    """L[b,c] = log sum_{j: y_j = c} exp(scores[b,j]); -inf for classes absent from the support
    (SURVEY.md B.1).  labels: (N,) shared or (B,N) per-query."""
    scores = _f64(scores)
    labels = np.asarray(labels)
    B, N = scores.shape
    L = np.full((B, n_classes), -np.inf)
    m = scores.max(axis=1, keepdims=True) if N else np.zeros((B, 1))
    e = np.exp(scores - m)
    W = np.zeros((B, n_classes))
    if labels.ndim == 1:
        np.add.at(W.T, labels, e.T)
    else:
        rows = np.repeat(np.arange(B), N)
        np.add.at(W, (rows, labels.reshape(-1)), e.reshape(-1))
    with np.errstate(divide="ignore"):
        L = np.log(W) + m
    return L


def logp_from_class_lse(L):
    """out[b,c] = log( exp(L[b,c] - logsumexp_c L[b,:]) + 1e-12 )   (nwhead/nw.py:285-289)."""
    L = _f64(L)
    m = L.max(axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        Z = m + np.log(np.exp(L - m).sum(axis=1, keepdims=True))
        P = np.exp(L - Z)
    return np.log(P + LOG_EPS)


def nw_forward(q, s, y, n_classes, kind="euclidean", logit_scale=CLIP_LOGIT_SCALE_INIT):
    """NWHead.forward (nwhead/nw.py:266-289): log(softmax_j(kernel(q, s_j)) @ onehot(y) + 1e-12).
    q (B,d); s (N,d)|(B,N,d); y (N,)|(B,N) ints < n_classes.  Returns (B, n_classes) float64."""
    scores = pairwise_scores(q, s, kind, logit_scale)
    return logp_from_class_lse(class_lse(scores, y, n_classes))


def nw_probs(q, s, y, n_classes, kind="euclidean", logit_scale=CLIP_LOGIT_SCALE_INIT):
    """softmax-weighted one-hot aggregation without the log (nwhead/nw.py:285-287)."""
    L = class_lse(pairwise_scores(q, s, kind, logit_scale), y, n_classes)
    m = L.max(axis=1, keepdims=True)
    E = np.exp(L - m)
    return E / E.sum(axis=1, keepdims=True)


# --------------------------------------------------------------------------------------
# a4: backward of a1-a3  (autograd through nwhead/nw.py:276-289; closed form SURVEY.md B.2)
# --------------------------------------------------------------------------------------
def nw_backward(q, s, y, n_classes, grad_out, kind="euclidean", logit_scale=CLIP_LOGIT_SCALE_INIT):
    """Returns (dL/dq (B,d), dL/ds (N,d)) for a SHARED 2-D support, given grad_out (B,C) wrt the
    log-prob output.  For kind == 'clip' also returns d/d(logit_scale) as a third value."""
    q0 = _f64(q)
    s0 = _f64(s)
    g = _f64(grad_out)
    y = np.asarray(y)
    normalised = kind in ("hypersphere_euclidean", "cosine", "clip")
    if normalised:
        qn = np.maximum(np.sqrt((q0 * q0).sum(-1, keepdims=True)), NORM_EPS)
        sn = np.maximum(np.sqrt((s0 * s0).sum(-1, keepdims=True)), NORM_EPS)
        qq, ss = q0 / qn, s0 / sn
    else:
        qq, ss = q0, s0
    scores = pairwise_scores(q0, s0, kind, logit_scale)
    m = scores.max(axis=1, keepdims=True)
    e = np.exp(scores - m)
    p = e / e.sum(axis=1, keepdims=True)  # (B,N) softmax weights
    P = np.zeros((q0.shape[0], n_classes))
    np.add.at(P.T, y, p.T)
    gP = g / (P + LOG_EPS)  # d/dP of log(P + eps)
    gp = gP[:, y]  # (B,N)
    gs = p * (gp - (p * gp).sum(axis=1, keepdims=True))  # d/dscore
    extra = None
    if kind in ("euclidean", "hypersphere_euclidean"):
        dist = -scores
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.where(dist > 0, gs / dist, 0.0)  # cdist backward yields 0 at coincident points
        # score = -||q-s||  =>  dscore/dq = -(q-s)/dist
        gq = -(r.sum(axis=1, keepdims=True) * qq) + r @ ss
        gss = r.T @ qq - r.sum(axis=0)[:, None] * ss
    else:
        scale = np.exp(logit_scale) if kind == "clip" else 1.0
        gq = scale * (gs @ ss)
        gss = scale * (gs.T @ qq)
        if kind == "clip":
            extra = float((gs * (qq @ ss.T)).sum() * scale)
    if normalised:
        # chain through x^ = x / max(||x||, eps): J = (I - x^ x^T) / max(||x||, eps)
        gq = (gq - (gq * qq).sum(-1, keepdims=True) * qq) / qn
        gss = (gss - (gss * ss).sum(-1, keepdims=True) * ss) / sn
    if kind == "clip":
        return gq, gss, extra
    return gq, gss


# --------------------------------------------------------------------------------------
# (e) multi-GPU merge algebra  (new)
# --------------------------------------------------------------------------------------
def merge_class_lse(parts):
    """Exact merge of per-shard class-LSE tables: logaddexp over shards (elementwise)."""
    out = _f64(parts[0]).copy()
    for p in parts[1:]:
        out = np.logaddexp(out, _f64(p))
    return out


def rowstat_partials(scores, labels, n_classes):
    """(m, l, w) row-sharded partials of SURVEY.md B.3 for one shard."""
    scores = _f64(scores)
    m = scores.max(axis=1)
    e = np.exp(scores - m[:, None])
    w = np.zeros((scores.shape[0], n_classes))
    np.add.at(w.T, np.asarray(labels), e.T)
    return m, e.sum(axis=1), w


def merge_rowstat_partials(parts):
    """P[b,c] = sum_r e^{m_r-m} w_r[b,c] / sum_r e^{m_r-m} l_r[b]   (SURVEY.md B.3)."""
    m = np.max(np.stack([p[0] for p in parts]), axis=0)
    num = sum(np.exp(p[0] - m)[:, None] * p[2] for p in parts)
    den = sum(np.exp(p[0] - m) * p[1] for p in parts)
    return num / den[:, None]


# --------------------------------------------------------------------------------------
# a5: support-bank construction  (nwhead/utils.py:34-54, 142-159; nwhead/support.py:156-165)
# --------------------------------------------------------------------------------------
def separated_indices(targets):
    """get_separated_indices (nwhead/utils.py:142-159): list of per-class index lists, classes in
    sorted order of their label values, indices in dataset order."""
    targets = np.asarray(targets)
    uniq = np.unique(targets)
    return [np.nonzero(targets == u)[0].tolist() for u in uniq]


def full_bank_keys(targets, n_shot_full):
    """FullDataset.keys (nwhead/utils.py:40-48): first min(n_shot_full, min class count) indices of
    every class, concatenated class by class -> class-sorted, balanced bank order."""
    idx = separated_indices(targets)
    k = min(n_shot_full, min(len(l) for l in idx))
    keys = []
    for l in idx:
        keys += l[:k]
    return np.asarray(keys, dtype=np.int64)


def class_sort_permutation(labels):
    """Stable permutation that class-sorts an arbitrary label vector (what the bank-build kernel
    applies when handed an unsorted support) and the class offsets table (C+1 entries need C)."""
    labels = np.asarray(labels)
    return np.argsort(labels, kind="stable")


def class_offsets(sorted_labels, n_classes):
    counts = np.bincount(np.asarray(sorted_labels), minlength=n_classes)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


# --------------------------------------------------------------------------------------
# a6: cluster-mode centroids  (nwhead/utils.py:218-246 with n_clusters == 1)
# --------------------------------------------------------------------------------------
def class_centroids(feats, labels):
    """compute_clusters(..., n_clusters=1): KMeans with one cluster == the class mean
    (pinned to 1e-6 against sklearn by gen_golden.py).  Returns (centroids (U,d), class ids (U,))
    over the sorted unique labels (nwhead/utils.py:227, 232, 245)."""
    feats = _f64(feats)
    labels = np.asarray(labels)
    uniq = np.unique(labels)
    cent = np.stack([feats[labels == u].mean(axis=0) for u in uniq])
    return cent, uniq.astype(np.int64)


def sklearn_kmeans_restated(x, k, max_iter=300, tol=1e-4):
    """ONE class of compute_clusters(..., n_clusters=k) (nwhead/utils.py:230:
    `KMeans(n_clusters=k, random_state=0).fit(x).cluster_centers_`), restated from scikit-learn 1.9.0 — the
    un-vendored, unpinned third-party dependency that owns this arithmetic; pinned to the installed
    scikit-learn by oracle/gen_golden.py on ambiguous data.  What is restated, in sklearn's order:

      KMeans.fit (_kmeans.py): X -= X.mean(axis=0) in float32; tol_abs = mean(var(X, axis=0)) * tol; n_init = 1
        ('auto' with k-means++); a FRESH RandomState(0) for every fit — every class sees the same random numbers.
      _kmeans_plusplus: first centre = rs.choice(n, p=uniform) (one random_sample through the normalised cdf);
        each further centre: n_local_trials = 2 + int(log k) candidates at searchsorted(cumsum_f32(closest_dist_sq),
        rs.uniform(size=trials) * current_pot); distances in float64 rounded to float32
        (_euclidean_distances_upcast); the candidate with the smallest potential wins.
      _kmeans_single_lloyd: labels = argmin distance (first minimum), centres = means; stop when the labels repeat
        (strict convergence) or when the total squared centre shift <= tol_abs; max_iter = 300.
    Not restated: BLAS summation orders (float32 dot / GEMM) and the relocation of emptied clusters — an emptied
    cluster keeps its centre.  Returns (k, d) float32 centres in sklearn's order."""
    X = np.array(x, dtype=np.float32, order="C", copy=True)
    n = len(X)
    if n < k:
        raise ValueError("a class has fewer rows than n_clusters")
    mean = X.mean(axis=0)
    X -= mean
    tol_abs = float(np.mean(np.var(X, axis=0)) * tol)
    rs = np.random.RandomState(0)
    w = np.ones(n, dtype=np.float32)
    Xd = X.astype(np.float64)

    def dist2(rows):  # float64 arithmetic, rounded to float32, clipped at zero
        rd = Xd[rows]
        d2 = (rd * rd).sum(1)[:, None] + (Xd * Xd).sum(1)[None, :] - 2.0 * rd @ Xd.T
        return np.maximum(d2.astype(np.float32), 0)

    trials = 2 + int(np.log(k))
    ids = [int(rs.choice(n, p=w / w.sum()))]
    closest = dist2([ids[0]])[0]
    pot = np.float32(closest.astype(np.float64).sum())
    for _ in range(1, k):
        rand_vals = rs.uniform(size=trials) * pot
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)
        np.clip(cand, None, n - 1, out=cand)
        dc = np.minimum(closest, dist2(cand))
        cpot = dc.astype(np.float64).sum(1).astype(np.float32)
        best = int(np.argmin(cpot))
        pot, closest = cpot[best], dc[best]
        ids.append(int(cand[best]))
    cent = Xd[ids].copy()
    labels_old = np.full(n, -1)
    for _ in range(max_iter):
        d2 = (Xd * Xd).sum(1)[:, None] + (cent * cent).sum(1)[None, :] - 2.0 * Xd @ cent.T
        labels = d2.argmin(1)
        new = cent.copy()
        for j in range(k):
            if (labels == j).any():
                new[j] = X[labels == j].astype(np.float64).mean(0)
        shift = float(((new - cent) ** 2).sum())
        cent = new
        if np.array_equal(labels, labels_old) or shift <= tol_abs:
            break
        labels_old = labels
    return (cent + mean.astype(np.float64)).astype(np.float32)


def kmeans_centroids(feats, labels, k):
    """compute_clusters(feats, labels, k > 1) (nwhead/utils.py:218-233, closest=False) through
    sklearn_kmeans_restated: (centroids (U*k, d) float32, labels (U*k,)) over the sorted unique labels, the rows of a
    class in dataset order, the centroids of a class in scikit-learn's order.  PINNED: gen_golden.py asserts equality
    with the reference (scikit-learn) row for row on unambiguous AND on overlapping clusters."""
    feats = np.asarray(feats, dtype=np.float32)
    labels = np.asarray(labels)
    out, out_y = [], []
    for c in np.unique(labels):
        out.append(sklearn_kmeans_restated(feats[labels == c], k))
        out_y += [c] * k
    return np.concatenate(out), np.asarray(out_y, dtype=np.int64)


def kmeans_inertia(feats, labels, centroids, k):
    """Sum over rows of the squared distance to the nearest of their class's k centroids (the k-means objective);
    centroids (U*k, d) over the sorted unique labels.  Returns one value per class."""
    feats, centroids = _f64(feats), _f64(centroids)
    labels = np.asarray(labels)
    out = []
    for i, c in enumerate(np.unique(labels)):
        x = feats[labels == c]
        d2 = ((x[:, None, :] - centroids[None, i * k:(i + 1) * k, :]) ** 2).sum(-1)
        out.append(d2.min(1).sum())
    return np.asarray(out)


def match_centroid_sets(a, b, k):
    """Largest |difference| between two (U*k, d) centroid arrays after matching, inside every class, each row
    of `a` with its nearest row of `b` (the order of KMeans centres within a class carries no meaning).
    Returns inf when the matching is not one-to-one."""
    a, b = _f64(a), _f64(b)
    worst = 0.0
    for c0 in range(0, len(a), k):
        d2 = ((a[c0:c0 + k, None, :] - b[None, c0:c0 + k, :]) ** 2).sum(-1)
        nearest = d2.argmin(1)
        if len(set(nearest.tolist())) != k:
            return float("inf")
        worst = max(worst, np.abs(a[c0:c0 + k] - b[c0:c0 + k][nearest]).max())
    return worst


# --------------------------------------------------------------------------------------
# a8: support_influence  (util/metric.py:23-50)
# --------------------------------------------------------------------------------------
def support_influence(softmaxes, qlabel_idx, sweights, slabel_idx):
    """infl[b,j] = log( (p_b - p_b w[b,j]) / (p_b - w[b,j] 1[y_j == y_b]) ),  p_b = softmaxes[b, y_b]
    (util/metric.py:42-47).  Takes categorical labels (the argmax of the reference's one-hots).
    slabel_idx (N,) -> (B,N).  slabel_idx (B,N) reproduces the reference's broadcast quirk
    (util/metric.py:43 takes argmax over the WHOLE slabels) -> (B,B,N)  (SURVEY.md A.8)."""
    P = _f64(softmaxes)
    w = _f64(sweights)
    qy = np.asarray(qlabel_idx)
    sy = np.asarray(slabel_idx)
    B = P.shape[0]
    p = P[np.arange(B), qy]
    with np.errstate(divide="ignore", invalid="ignore"):
        if sy.ndim == 1:
            ind = (sy[None, :] == qy[:, None]).astype(np.float64)
            return np.log((p[:, None] - p[:, None] * w) / (p[:, None] - w * ind))
        ind = (sy[None, :, :] == qy[:, None, None]).astype(np.float64)  # (B, Bs, N)
        pb = p[:, None, None]
        wb = w[:, None, :]
        return np.log((pb - pb * wb) / (pb - wb * ind))


# --------------------------------------------------------------------------------------
# a9: neighbour ranking  (nwhead/nw.py:245-249)
# --------------------------------------------------------------------------------------
def neighbor_ranking(q, s, kind="euclidean"):
    """argsort of scores, descending (torch.argsort(..., descending=True)); ties-free data only."""
    sc = pairwise_scores(q, s, kind)
    return np.argsort(-sc, axis=-1, kind="stable")


def topk_neighbors(q, s, k, kind="euclidean"):
    return neighbor_ranking(q, s, kind)[:, :k]

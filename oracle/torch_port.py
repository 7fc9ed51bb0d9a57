"""fp32 torch-CPU port of the reference's op sequence.  TEST / BASELINE INFRASTRUCTURE ONLY.

Used (a) as the CPU baseline that bench.py times on the GPU box's host cores (the Python
reference at /root/reference cannot travel there) and (b) by tests as an fp32 stand-in whose
rounding behaviour matches the reference (torch.cdist's mm-vs-exact switch, fp32 softmax).
It issues the same library calls in the same order as the reference, so its cost and its
memory behaviour (materialised (B,N,d) / (B,N,C) operands) are representative.

Parity status: PINNED — bit-compared with the imported reference by oracle/gen_golden.py.
"""
import torch
import torch.nn.functional as F


def port_scores(x, y, kind, logit_scale=None):
    """nwhead/kernel.py:13-44 on (bs, num_x, d) x (bs, num_y, d) -> (bs, num_x, num_y)."""
    if kind == "euclidean":
        return -torch.cdist(x, y)
    if kind == "dotproduct":
        return torch.bmm(x, y.transpose(-2, -1))
    xn, yn = F.normalize(x, dim=-1), F.normalize(y, dim=-1)
    if kind == "hypersphere_euclidean":
        return -torch.cdist(xn, yn)
    if kind == "cosine":
        return torch.bmm(xn, yn.transpose(-2, -1))
    if kind == "clip":
        return logit_scale.exp() * torch.bmm(xn, yn.transpose(-2, -1))
    raise NotImplementedError(kind)


def port_nw_forward(x, sx, sy, n_classes, kind="euclidean", logit_scale=None):
    """nwhead/nw.py:266-289."""
    b = x.shape[0]
    onehot = F.one_hot(sy, n_classes).float()
    if sx.dim() == x.dim():
        sx = sx.unsqueeze(0).expand(b, *sx.shape)
        onehot = onehot.unsqueeze(0).expand(b, *onehot.shape)
    w = F.softmax(port_scores(x.unsqueeze(1), sx, kind, logit_scale), dim=-1)
    return torch.log(torch.bmm(w, onehot).squeeze(1) + 1e-12)


def port_support_influence(softmaxes, qlabels, sweights, slabels):
    """util/metric.py:23-50 including its per-query Python loop (that loop IS the baseline)."""
    rows = []
    for b in range(len(softmaxes)):
        qcat = qlabels[b].argmax(-1).item()
        scat = slabels.argmax(-1)
        p = softmaxes[b][qcat]
        ind = (scat == qcat).long()
        rows.append(torch.log((p - p * sweights[b]) / (p - sweights[b] * ind))[None])
    return torch.cat(rows, dim=0)


def port_class_centroids(feats, labels):
    """nwhead/utils.py:218-246 with n_clusters=1: per-class boolean-mask gather + mean (the k=1
    KMeans fixed point), classes in sorted-unique order."""
    cents, ys = [], []
    for c in torch.unique(labels).tolist():
        cents.append(feats[labels == c].mean(dim=0, keepdim=True))
        ys.append(c)
    return torch.cat(cents, dim=0), torch.tensor(ys)


def port_compute_clusters(embeddings, labels, n_clusters):
    """nwhead/utils.py:218-233 (closest=False): one scikit-learn KMeans(n_clusters, random_state=0) fit per class
    over a boolean-mask gather, classes in sorted-unique order.  This is the CPU baseline of cluster-mode
    precompute (BASELINE config 4); for n_clusters=1 its result is the class mean (port_class_centroids)."""
    from sklearn.cluster import KMeans

    cents, ys = [], []
    for c in torch.unique(labels).tolist():
        km = KMeans(n_clusters=n_clusters, random_state=0).fit(embeddings[labels == c].numpy())
        cents.append(torch.from_numpy(km.cluster_centers_).float())
        ys += [c] * n_clusters
    return torch.cat(cents, dim=0), torch.tensor(ys)

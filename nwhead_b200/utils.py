"""Host-side dataset adaptors and support samplers (API of the reference's nwhead/utils.py).

Only index bookkeeping lives here; it stays on the host so that a seeded run draws the same numpy
random numbers, in the same order, as the reference.  Feature gathers, centroid
reduction and ranking run on the GPU (compute_clusters -> nw_class_centroids, KNN -> nw_rank_rows).
"""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from . import _abi
from ._abi import check, load, ptr, stream_of


class DatasetMetadata(Dataset):
    """(x, y) dataset -> (x, y, metadata[idx]) triples (reference nwhead/utils.py:7-19)."""

    def __init__(self, dataset, metadata):
        super().__init__()
        self.dataset = dataset
        self.targets = dataset.targets
        self.metadata = metadata

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, idx):
        item = self.dataset[idx]
        return item[0], item[1], self.metadata[idx]


class FeatureDataset(Dataset):
    """Precomputed features as a dataset (reference nwhead/utils.py:21-32)."""

    def __init__(self, features, targets, metadata):
        super().__init__()
        self.features, self.targets, self.metadata = features, targets, metadata

    def __len__(self):
        return len(self.features)

    def __getitem__(self, idx):
        return self.features[idx], self.targets[idx], self.metadata[idx]


def get_separated_indices(vals):
    """Per-class index lists, classes ordered by sorted label value, indices in dataset order
    (reference nwhead/utils.py:142-159).  [0, 1, 1, 2, 3] -> [[0], [1, 2], [3], [4]]."""
    if torch.is_tensor(vals):
        vals = vals.cpu().detach().numpy()
    vals = np.asarray(vals)
    order = np.argsort(vals, kind="stable")
    sorted_vals = vals[order]
    cuts = np.flatnonzero(sorted_vals[1:] != sorted_vals[:-1]) + 1
    return [chunk.tolist() for chunk in np.split(order, cuts)]


class FullDataset(Dataset):
    """Class-balanced, class-major view used to build the full bank: the first
    min(n_shot_full, smallest class) items of every class (reference nwhead/utils.py:34-54)."""

    def __init__(self, underlying_dataset, n_shot_full):
        super().__init__()
        self.underlying_dataset = underlying_dataset
        self.indices = get_separated_indices(underlying_dataset.targets)
        per_class = min(n_shot_full, min(len(l) for l in self.indices))
        self.keys = [i for l in self.indices for i in l[:per_class]]

    def __getitem__(self, key):
        return self.underlying_dataset[self.keys[key]]

    def __len__(self):
        return len(self.keys)


class InfiniteUniformClassLoader(DataLoader):
    """n_way / n_shot support sampler (reference nwhead/utils.py:99-140).  ``sample_indices`` issues
    exactly the reference's np.random.choice calls; ``next`` additionally loads + collates the items."""

    def __init__(self, dataset, n_shot, n_way=None):
        self.dataset = dataset
        self.indices = get_separated_indices(dataset.targets)
        self.n_classes = len(self.indices)
        self.n_shot = n_shot
        self.n_way = n_way
        if n_way:
            assert n_way <= len(self.indices)
        super().__init__(dataset)

    def __iter__(self):
        return self

    def __next__(self):
        raise NotImplementedError

    def sample_indices(self, qy=None):
        if self.n_way:
            assert len(qy) <= self.n_way, "qy must be smaller than n_way"
            qy = qy.cpu().detach().numpy()
            probs = np.ones(len(self.indices))
            probs[qy] = 0
            probs /= probs.sum()
            extra = np.random.choice(self.n_classes, size=(self.n_way - len(qy)), replace=False, p=probs)
            rows = [self.indices[c] for c in np.concatenate([extra, qy])]
        else:
            rows = self.indices
        return np.array([np.random.choice(r, size=self.n_shot, replace=False) for r in rows]).flatten()

    def next(self, qy=None):
        idx = self.sample_indices(qy)
        return self.collate_fn([self.dataset[i] for i in idx])


def compute_clusters(embeddings, labels, n_clusters, closest=False):
    """Cluster-mode support (reference nwhead/utils.py:218-246): per class, KMeans(n_clusters, random_state=0).

    n_clusters == 1 (the NWNet default, nwhead/nw.py:24): KMeans with one cluster is the class mean, one pass of
    nw_class_centroids.  n_clusters > 1: k-means++ seeding + Lloyd iterations for ALL classes at once on the GPU
    (kmeans_centroids below: nw_kmeans_assign + nw_class_centroids per iteration).  scikit-learn's random stream
    cannot be reproduced, so for k > 1 the centroids are those of the same objective run to strict convergence —
    equal to the reference's wherever the clustering is unambiguous (tests/golden/clusters.npz), in any order
    within a class (the NW head is invariant to the order of its support rows).
    closest=True returns, for every centroid, the nearest real embedding of its class (nwhead/utils.py:234-241).
    Returns (centroids (U*k, d) fp32, labels (U*k,) int64) over the sorted unique labels."""
    dev = _abi.require_cuda(embeddings, labels)
    lib = load()
    feats = embeddings.detach()
    if feats.dtype != torch.float32:
        feats = feats.float()
    if feats.stride(1) != 1:
        feats = feats.contiguous()
    labels = labels.detach().to(torch.int64).contiguous()
    n, d = feats.shape
    n_classes = int(labels.max().item()) + 1
    st = stream_of(dev)
    lab32 = torch.empty((n,), dtype=torch.int32, device=dev)
    status = torch.empty((2,), dtype=torch.int32, device=dev)
    check(lib.nw_labels_to_i32(ptr(labels), None, n, n_classes, ptr(lab32), ptr(status), st), "nw_labels_to_i32")
    perm, sorted32 = None, lab32
    if status[1].item():
        perm = torch.sort(labels, stable=True).indices.contiguous()
        sorted32 = torch.empty_like(lab32)
        check(lib.nw_labels_to_i32(ptr(labels), ptr(perm), n, n_classes, ptr(sorted32), ptr(status), st),
              "nw_labels_to_i32")
    offsets = torch.empty((n_classes + 1,), dtype=torch.int32, device=dev)
    check(lib.nw_class_offsets(ptr(sorted32), n, n_classes, ptr(offsets), st), "nw_class_offsets")
    if n_clusters == 1 and not closest:
        return class_centroids(feats, perm, offsets, n_classes)
    return kmeans_centroids(feats, lab32, perm, offsets, n_classes, n_clusters, closest=closest)


def _kmeans_assign(feats, group, centroids, k, order=None):
    """nw_kmeans_assign: (assignment group*k + j, squared distance to the chosen centroid) for every row;
    order = the class-sorted visiting order of the rows (None: they are class-sorted already)."""
    lib = load()
    dev = feats.device
    n, d = feats.shape
    assign = torch.empty((n,), dtype=torch.int32, device=dev)
    dist = torch.empty((n,), dtype=torch.float32, device=dev)
    check(lib.nw_kmeans_assign(ptr(feats), d, feats.stride(0), ptr(group), ptr(order), n, ptr(centroids), k, ptr(assign),
                               ptr(dist), stream_of(dev)), "nw_kmeans_assign")
    return assign, dist


def kmeans_centroids(feats, group, perm, offsets, n_classes, k, closest=False, seed=0, max_iter=300):
    """Per-class k-means for all classes at once (reference nwhead/utils.py:227-231, one sklearn fit per class).

    feats (N, d) fp32; group (N,) int32 class of every row; perm/offsets: the class-sorted order of the rows
    (SupportBank.perm / .offsets).  Seeding is k-means++ (each further centre drawn with probability
    proportional to the squared distance to the nearest chosen one) from numpy RandomState(seed): one uniform per
    class and centre.  Lloyd iterations run until no row changes cluster (sklearn's strict-convergence stop) or
    max_iter (sklearn's default 300); an emptied cluster keeps its centre."""
    import numpy as np

    lib = load()
    dev = feats.device
    n, d = feats.shape
    st = stream_of(dev)
    lo, hi = offsets[:-1].long(), offsets[1:].long()
    cnt = hi - lo
    present = (cnt > 0).nonzero().flatten()
    if bool((cnt[present] < k).any()):
        raise ValueError(f"a class has fewer rows than n_clusters={k}")   # as sklearn's KMeans.fit does
    to_src = (lambda pos: pos) if perm is None else (lambda pos: perm[pos])
    rng = np.random.RandomState(seed)
    cent = torch.zeros((n_classes, k, d), dtype=torch.float32, device=dev)

    def draw():
        return torch.from_numpy(rng.random_sample(n_classes)).to(dev)

    pick = lo + torch.minimum((draw() * cnt).floor().long(), (cnt - 1).clamp_min(0))
    cent[:, 0] = feats[to_src(pick.clamp_max(n - 1))]
    mind = torch.full((n,), float("inf"), dtype=torch.float32, device=dev)
    last = (hi - 1).clamp_min(0)
    for j in range(1, k):
        _, dist = _kmeans_assign(feats, group, cent[:, j - 1].contiguous(), 1, perm)
        mind = torch.minimum(mind, dist)
        cs = torch.cumsum((mind if perm is None else mind[perm]).double(), 0)
        base = torch.where(lo > 0, cs[(lo - 1).clamp_min(0)], torch.zeros_like(cs[:1]))
        target = base + draw() * (cs[last] - base)
        pos = torch.minimum(torch.maximum(torch.searchsorted(cs, target, right=True), lo), last)
        cent[:, j] = feats[to_src(pos)]
    cent = cent.reshape(n_classes * k, d)

    ws_bytes = lib.nw_class_centroids_workspace_bytes(n_classes * k, d)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=dev)
    offs2 = torch.empty((n_classes * k + 1,), dtype=torch.int32, device=dev)
    new = torch.empty_like(cent)
    prev = None
    for _ in range(max_iter):
        assign, _ = _kmeans_assign(feats, group, cent, k, perm)
        if prev is not None and torch.equal(assign, prev):
            break
        prev = assign
        srt = torch.sort(assign, stable=True)
        check(lib.nw_class_offsets(ptr(srt.values), n, n_classes * k, ptr(offs2), st), "nw_class_offsets")
        check(lib.nw_class_centroids(ptr(feats), d, feats.stride(0), ptr(srt.indices), ptr(offs2), n_classes * k,
                                     ptr(new), ptr(ws), ws_bytes, st), "nw_class_centroids")
        cent = torch.where((offs2[1:] > offs2[:-1])[:, None], new, cent)

    cent = cent.reshape(n_classes, k, d)
    if closest:  # nearest real embedding of the class to every centroid (nwhead/utils.py:234-241)
        rows = torch.arange(n, device=dev)
        g64 = group.long()
        for j in range(k):
            _, dist = _kmeans_assign(feats, group, cent[:, j].contiguous(), 1, perm)
            best = torch.full((n_classes,), float("inf"), device=dev).scatter_reduce(0, g64, dist, "amin")
            first = torch.full((n_classes,), n, dtype=torch.int64, device=dev).scatter_reduce(
                0, g64, torch.where(dist == best[g64], rows, torch.full_like(rows, n)), "amin")
            cent[:, j] = feats[first.clamp_max(n - 1)]
    out = cent.index_select(0, present).reshape(-1, d)
    return out, present.to(torch.int64).repeat_interleave(k)


def class_centroids(feats, perm, offsets, n_classes):
    """nw_class_centroids + compaction to the classes that are present."""
    lib = load()
    dev = feats.device
    d = feats.shape[1]
    out = torch.empty((n_classes, d), dtype=torch.float32, device=dev)
    ws_bytes = lib.nw_class_centroids_workspace_bytes(n_classes, d)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=dev)
    check(lib.nw_class_centroids(ptr(feats), d, feats.stride(0), ptr(perm), ptr(offsets), n_classes, ptr(out),
                                 ptr(ws), ws_bytes, stream_of(dev)), "nw_class_centroids")
    present = (offsets[1:] > offsets[:-1]).nonzero().flatten()
    if present.numel() != n_classes:
        out = out.index_select(0, present)
    return out, present.to(torch.int64)


def rank_rows(scores: torch.Tensor, k=None) -> torch.Tensor:
    """Indices of each row's scores in descending order (first k): nw_rank_rows."""
    lib = load()
    dev = _abi.require_cuda(scores)
    scores = scores.detach().float().contiguous()
    r, n = scores.shape
    k = n if k is None else min(int(k), n)
    out = torch.empty((r, k), dtype=torch.int64, device=dev)
    ws_bytes = lib.nw_rank_rows_workspace_bytes(r, n)
    ws = torch.empty((ws_bytes // 8,), dtype=torch.int64, device=dev)
    check(lib.nw_rank_rows(ptr(scores), r, n, k, ptr(out), ptr(ws), ws_bytes, stream_of(dev)), "nw_rank_rows")
    return out


class KNN:
    """Exact k-nearest-neighbour support (reference nwhead/utils.py:178-193): the k nearest bank rows
    of every query, concatenated into ONE shared support of B*k rows."""

    BANK_SEARCH_MIN_ROWS = 65536

    def __init__(self, data, labels, n_neighbors=20, bank=None) -> None:
        self.data, self.labels, self.n_neighbors = data, labels, n_neighbors
        # an euclidean SupportBank over the same rows: large searches go through its tensor-core block search
        # (SupportBank.topk_exact: same ranking as the dense scores, bit for bit)
        self.bank = bank if bank is not None and bank.kind == "euclidean" else None

    def __call__(self, x):
        from .kernel import dense_scores

        if self.bank is not None and len(self.bank) >= self.BANK_SEARCH_MIN_ROWS:
            idx = self.bank.topk_exact(x, self.n_neighbors, self.data).flatten()
        else:
            idx = rank_rows(dense_scores("euclidean", x, self.data), self.n_neighbors).flatten()
        return self.data.index_select(0, idx), self.labels.index_select(0, idx)


class HNSW(KNN):
    """mode='hnsw' (reference nwhead/utils.py:195-216 builds an approximate hnswlib index, space='l2').  On the
    B200 the exact search over the bank (dense scores + ranking) is cheap, so this returns the EXACT k nearest
    neighbours under the same metric — the result an HNSW index approximates.  hnswlib is not needed."""

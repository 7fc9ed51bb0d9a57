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
    """Cluster-mode support (reference nwhead/utils.py:218-246).

    n_clusters == 1 (the NWNet default, nwhead/nw.py:24): KMeans with one cluster is the class mean,
    computed on the GPU by nw_class_centroids.  n_clusters > 1 (k-means++ / Lloyd on the host in the
    reference, parity unpinned) is not provided.
    Returns (centroids (U*k, d) fp32, labels (U*k,) int64) over the sorted unique labels."""
    if n_clusters != 1 or closest:
        raise NotImplementedError(
            "only n_clusters=1 (the NWNet default: class means, nw_class_centroids) runs on the B200 path; the "
            "reference's k>1 / closest=True variants are host-side scikit-learn KMeans (nwhead/utils.py:227-241) "
            "and there is no CPU fallback here")
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
    perm = None
    if status[1].item():
        perm = torch.sort(labels, stable=True).indices.contiguous()
        check(lib.nw_labels_to_i32(ptr(labels), ptr(perm), n, n_classes, ptr(lab32), ptr(status), st),
              "nw_labels_to_i32")
    offsets = torch.empty((n_classes + 1,), dtype=torch.int32, device=dev)
    check(lib.nw_class_offsets(ptr(lab32), n, n_classes, ptr(offsets), st), "nw_class_offsets")
    return class_centroids(feats, perm, offsets, n_classes)


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

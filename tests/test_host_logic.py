"""CPU: host-side support-set bookkeeping of the drop-in package (no GPU, no compute calls) — class
index separation, class-balanced bank order, and numpy-RNG-stream parity of the samplers with the
reference (when /root/reference is importable; the properties are checked either way)."""
import numpy as np
import pytest
import torch

from nwhead_b200 import utils as U
from oracle import nw_oracle as O
from oracle.ref_import import load_reference, reference_available


class ToyDataset(torch.utils.data.Dataset):
    def __init__(self, targets):
        self.targets = list(targets)

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        return torch.full((2,), float(i)), self.targets[i]


def test_separated_indices_docstring_example():
    assert U.get_separated_indices([0, 1, 1, 2, 3]) == [[0], [1, 2], [3], [4]]
    # non-consecutive labels map to consecutive slots in sorted order (reference nwhead/utils.py:152-153)
    assert U.get_separated_indices(torch.tensor([7, 3, 7, 5])) == [[1], [3], [0, 2]]


def test_full_dataset_is_class_sorted_balanced_truncated():
    rng = np.random.default_rng(0)
    targets = rng.integers(0, 7, 200).tolist()
    fd = U.FullDataset(ToyDataset(targets), 12)
    keys = np.asarray(fd.keys)
    assert np.array_equal(keys, O.full_bank_keys(targets, 12))
    labs = np.asarray(targets)[keys]
    assert (np.diff(labs) >= 0).all()                       # class-sorted
    assert len(set(np.bincount(labs))) == 1                  # balanced
    assert len(fd) == 7 * min(12, np.bincount(targets).min())
    assert fd[3][1] == labs[3]


@pytest.mark.skipif(not reference_available(), reason="reference not mounted (GPU box)")
def test_sampler_rng_stream_matches_reference():
    ref = load_reference()
    rng = np.random.default_rng(1)
    targets = rng.integers(0, 9, 120).tolist()
    ds = ToyDataset(targets)
    for n_way, n_shot in [(None, 2), (5, 1), (6, 3)]:
        a = U.InfiniteUniformClassLoader(U.DatasetMetadata(ds, np.zeros(len(ds))), n_shot, n_way)
        b = ref.InfiniteUniformClassLoader(ref_meta(ref, ds), n_shot, n_way)
        qy = torch.tensor([3, 3, 8]) if n_way else None
        np.random.seed(42)
        xa = a.next(qy)
        np.random.seed(42)
        xb = b.next(qy)
        for ta, tb in zip(xa, xb):
            assert torch.equal(torch.as_tensor(ta), torch.as_tensor(tb))
        assert U.get_separated_indices(targets) == ref.get_separated_indices(targets)


def ref_meta(ref, ds):
    import importlib

    return importlib.import_module("nwhead.utils").DatasetMetadata(ds, np.zeros(len(ds)))


def test_sampler_properties():
    rng = np.random.default_rng(2)
    targets = rng.integers(0, 9, 150).tolist()
    ld = U.InfiniteUniformClassLoader(U.DatasetMetadata(ToyDataset(targets), np.zeros(150)), 2, n_way=5)
    np.random.seed(0)
    qy = torch.tensor([1, 1, 4])
    idx = ld.sample_indices(qy)
    labs = np.asarray(targets)[idx]
    assert len(idx) == 5 * 2                                 # n_way rows (query labels WITH duplicates) x n_shot
    assert {1, 4} <= set(labs.tolist())
    assert (labs[-6:] == np.repeat([1, 1, 4], 2)).all()      # query classes come last, duplicates kept
    with pytest.raises(AssertionError):
        ld.sample_indices(torch.arange(6))                  # len(qy) must be <= n_way


def test_unknown_names_raise_like_the_reference():
    import nwhead_b200

    with pytest.raises(NotImplementedError):
        nwhead_b200.get_kernel("nope")
    assert nwhead_b200.get_kernel("clip").logit_scale.item() == pytest.approx(np.log(1 / 0.07))
    with pytest.raises(AssertionError):  # as in the reference: the support set must expose .targets
        nwhead_b200.NWNet(torch.nn.Identity(), 3, support_dataset=[ToyDataset([0, 1, 2])])


def test_environment_split_matches_reference_semantics():
    """env_array -> one Subset per environment value (sorted), targets sliced alike; 'irm' builds one sampler per
    environment over ALL of its classes (reference nwhead/support.py:47-56, 84-93)."""
    from nwhead_b200.support import SupportSetEval, SupportSetTrain

    rng = np.random.default_rng(3)
    targets = rng.integers(0, 5, 60).tolist()
    env = np.array([(i // 4) % 3 for i in range(60)])
    ds = ToyDataset(targets)
    st = SupportSetTrain(ds, 5, 'irm', 2, env_array=env)
    assert len(st.env_datasets) == 3 and len(st.train_iter) == 3
    for e, sub in enumerate(st.env_datasets):
        idx = np.flatnonzero(env == e)
        assert np.array_equal(sub.indices, idx) and np.array_equal(sub.targets, np.asarray(targets)[idx])
    np.random.seed(5)
    sx, sy, sm = st.get_support(torch.tensor([0]))
    assert len(set(sm.tolist())) == 1                      # one environment per draw
    e = int(sm[0])
    assert sorted(set(sy.tolist())) == sorted(set(np.asarray(targets)[env == e].tolist()))
    assert len(sy) == 2 * len(set(sy.tolist()))             # n_shot items of every class of that environment
    se = SupportSetEval(ds, 5, 1, 3, env_array=env)
    assert len(se.support_loaders) == 3


def test_kmeans_first_seed_matches_numpy_choice():
    """The first k-means++ centre of every class is `RandomState(0).choice(n, p=uniform)` in scikit-learn
    (_kmeans_plusplus); the product draws the uniform once and maps it through numpy's normalised-cdf search for every
    class size (nwhead_b200.utils._first_seed_offsets)."""
    sizes = [1, 2, 3, 7, 29, 30, 100, 1279, 1280, 1281, 4096, 50001]
    rs = np.random.RandomState(0)
    u0 = rs.random_sample()
    got = U._first_seed_offsets(sizes, u0)
    for n in sizes:
        w = np.ones(n, dtype=np.float32)
        assert got[n] == np.random.RandomState(0).choice(n, p=w / w.sum()), n
    # the draws that follow are plain uniforms: rs.uniform(size=t) continues the same stream
    ref = np.random.RandomState(0)
    ref.choice(5, p=np.ones(5) / 5)
    assert np.array_equal(rs.uniform(size=3), ref.uniform(size=3))


def test_tensor_backward_routing():
    """NWHead's choice between the direct fp32 kernels and the tensor-core forward + backward (host logic)."""
    from nwhead_b200.backward import MIN_PAIRS, MIN_QUERIES, wants_tensor_path

    assert wants_tensor_path(4096, 1280000, 2, "auto")
    assert wants_tensor_path(MIN_QUERIES, MIN_PAIRS // MIN_QUERIES, 2, "auto")
    assert not wants_tensor_path(8, 1280000, 2, "auto")          # few queries: HBM-bound, the direct path reads S once
    assert not wants_tensor_path(4096, 1000, 2, "auto")          # small support
    assert not wants_tensor_path(4096, 1280000, 3, "auto")       # per-query supports are not a shared GEMM
    assert not wants_tensor_path(4096, 1280000, 2, "direct")
    assert wants_tensor_path(2, 30, 2, "tensor") and not wants_tensor_path(2, 30, 3, "tensor")


def test_split_k_choice_fills_waves_and_leaves_no_empty_slice():
    """nwhead_b200.backward.choose_kslices (host logic of nw_dense_products' split-K)."""
    from nwhead_b200.backward import choose_kslices

    assert choose_kslices(40000, 74, 64) == 1          # plenty of units: no split
    assert choose_kslices(1, 74, 3) == 1               # too short to split
    ks = choose_kslices(128, 74, 20000)                # config 3 grad_q: 128 units on 74 CTA pairs
    assert 128 * ks % 74 <= 74 and -(-128 * ks // 74) * (-(-20000 // ks) + 16) < 2 * (20000 + 16)
    for units, kb in [(8, 20000), (32, 20000), (9, 130), (8, 2500), (3, 17)]:
        ks = choose_kslices(units, 74, kb)
        per = -(-kb // ks)
        assert 1 <= ks <= 64 and per * (ks - 1) < kb <= per * ks  # every slice holds at least one k-block

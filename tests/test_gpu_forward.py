"""GPU: the fused tcgen05 forward (through the C ABI) against the oracle and the reference fixtures."""
import numpy as np
import pytest
import torch

from oracle import nw_oracle as O
from gpu_util import assert_head_parity, clustered_features

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_bank(q, s, y, C, kind, precision, scale=1.0):
    from nwhead_b200 import SupportBank

    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), C, kind, precision)
    out = bank.forward(torch.from_numpy(q).to(DEV), scale)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
@pytest.mark.parametrize("kind", ["euclidean", "cosine", "hypersphere_euclidean", "dotproduct"])
@pytest.mark.parametrize("shape", [(8, 10, 30, 64), (37, 20, 29, 512), (130, 7, 300, 96), (300, 1000, 1, 128)])
def test_forward_matches_oracle(cuda_lib, shape, kind, precision):
    B, C, per, d = shape
    # one support row per class (cluster / random mode shape): keep the classes separated so that top-1 is
    # decided by a margin larger than the bf16 rounding error rather than by near-ties
    q, s, y, _ = clustered_features(C, per, d, B, seed=B + d, spread=0.25 if per == 1 else 1.0)
    if kind == "dotproduct":  # keep exp() in range: the reference has no temperature either
        q, s = q * 0.05, s * 0.05
    out = run_bank(q, s, y, C, kind, precision)
    ref = O.nw_forward(q, s, y, C, kind)
    perr, _ = assert_head_parity(out, ref)
    if precision == "bf16x3":
        assert perr < 2e-5, f"3-product split should be near fp32, got {perr:.2e}"


def test_clip_scale(cuda_lib):
    q, s, y, _ = clustered_features(12, 40, 64, 16, seed=3)
    out = run_bank(q, s, y, 12, "clip", "bf16", scale=float(np.exp(O.CLIP_LOGIT_SCALE_INIT)))
    assert_head_parity(out, O.nw_forward(q, s, y, 12, "clip"))


@pytest.mark.parametrize("kind", O.KERNEL_KINDS)
@pytest.mark.parametrize("case", ["mm_medium", "wide"])
def test_reference_fixtures_through_nwhead(cuda_lib, golden_head, case, kind):
    """NWHead.forward drop-in on the reference's own inputs: unsorted labels, duplicates, absent
    classes, a query that coincides with a support row."""
    import nwhead_b200

    g = golden_head
    C = int(g[f"{case}/C"])
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel(kind), C).to(DEV)
    with torch.no_grad():
        out = head(torch.from_numpy(g[f"{case}/q"]).to(DEV), torch.from_numpy(g[f"{case}/s"]).to(DEV),
                   torch.from_numpy(g[f"{case}/y"]).to(DEV))
    ref = g[f"{case}/{kind}/logp"]
    assert_head_parity(out, ref)
    absent = np.setdiff1d(np.arange(C), np.unique(g[f"{case}/y"]))
    got = out.cpu().numpy()[:, absent]
    assert np.array_equal(got, np.full_like(got, np.log(np.float32(1e-12))))  # exactly log(1e-12)


def test_multi_chunk_boundaries_and_ragged_tail(cuda_lib):
    """N not a multiple of the 256-row tile, classes of uneven size cut by tile and chunk boundaries,
    B not a multiple of 128; compared class-LSE by class-LSE with the oracle."""
    from nwhead_b200 import SupportBank, _abi

    rng = np.random.default_rng(5)
    C, d, B = 23, 128, 200
    sizes = rng.integers(1, 900, C)
    sizes[4] = 0  # an absent class in the middle
    y = np.repeat(np.arange(C), sizes).astype(np.int64)
    mu = rng.normal(size=(C, d))
    s = (mu[y] + rng.normal(size=(len(y), d))).astype(np.float32)
    q = (mu[rng.integers(0, C, B)] + rng.normal(size=(B, d))).astype(np.float32)
    plan = _abi.forward_plan(B, len(y))
    assert plan.chunks > 1 and len(y) % 256 != 0
    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), C, "euclidean", "bf16x3")
    lse = bank.class_lse(torch.from_numpy(q).to(DEV)).cpu().numpy()
    ref = O.class_lse(O.pairwise_scores(q, s, "euclidean"), y, C)
    assert np.isneginf(lse[:, 4]).all() and np.isneginf(ref[:, 4]).all()
    fin = np.isfinite(ref)
    assert np.abs(lse[fin] - ref[fin]).max() < 2e-3
    out = bank.forward(torch.from_numpy(q).to(DEV))
    assert_head_parity(out, O.logp_from_class_lse(ref), prob_tol=2e-5)


def test_bitwise_reproducible(cuda_lib):
    q, s, y, _ = clustered_features(50, 100, 256, 300, seed=9)
    a = run_bank(q, s, y, 50, "euclidean", "bf16")
    b = run_bank(q, s, y, 50, "euclidean", "bf16")
    assert torch.equal(a, b)


def test_large_distances_do_not_underflow(cuda_lib):
    """Distances of several hundred: exp(-dist) underflows in fp32 without the running max."""
    q, s, y, _ = clustered_features(10, 40, 512, 32, seed=1)
    q, s = q * 40.0, s * 40.0
    out = run_bank(q, s, y, 10, "euclidean", "bf16x3")
    assert_head_parity(out, O.nw_forward(q, s, y, 10, "euclidean"), prob_tol=1e-3)


def test_label_errors_raise(cuda_lib):
    from nwhead_b200 import SupportBank

    s = torch.randn(40, 16, device=DEV)
    with pytest.raises(RuntimeError, match="num_classes"):
        SupportBank.build(s, torch.full((40,), 7, device=DEV, dtype=torch.int64), 5)
    with pytest.raises(RuntimeError, match="LongTensor"):
        SupportBank.build(s, torch.zeros(40, device=DEV, dtype=torch.int32), 5)


def test_config1_shape(cuda_lib):
    """BASELINE config 1 head shapes: bank 5800 x 512 (29 per class, 200 classes), batch 8."""
    q, s, y, _ = clustered_features(200, 29, 512, 8, seed=1)
    for precision in ("auto", "bf16"):
        out = run_bank(q, s, y, 200, "euclidean", precision)
        assert_head_parity(out, O.nw_forward(q, s, y, 200, "euclidean"))


def test_random_shapes_against_oracle(cuda_lib):
    """Seeded sweep over ragged shapes: B around the 128/256-row tile edges (single CTA and CTA pair), N from a
    handful of rows to several chunks, uneven class sizes with empty classes, d not a multiple of 64."""
    rng = np.random.default_rng(2026)
    for trial in range(24):
        B = int(rng.choice([1, 7, 127, 128, 129, 255, 257, 300, 513]))
        C = int(rng.integers(2, 60))
        sizes = rng.integers(0, int(rng.choice([3, 40, 400])) + 1, C)
        sizes[rng.integers(0, C)] += 26  # at least one non-empty class, N > 25
        d = int(rng.choice([8, 48, 64, 100, 192, 520]))
        kind = str(rng.choice(["euclidean", "cosine", "hypersphere_euclidean"]))
        y = np.repeat(np.arange(C), sizes).astype(np.int64)
        mu = rng.normal(size=(C, d)) * 1.5
        s = (mu[y] + rng.normal(size=(len(y), d))).astype(np.float32)
        q = (mu[rng.integers(0, C, B)] + rng.normal(size=(B, d))).astype(np.float32)
        out = run_bank(q, s, y, C, kind, "bf16x3").cpu().numpy()
        ref = O.nw_forward(q, s, y, C, kind)
        perr = np.abs(np.exp(out) - np.exp(ref)).max()
        assert perr < 5e-5, (trial, B, len(y), d, C, kind, perr)
        empty = np.flatnonzero(sizes == 0)
        assert np.array_equal(out[:, empty], np.full((B, len(empty)), np.log(np.float32(1e-12)), np.float32))


@pytest.mark.parametrize("kind", ["euclidean", "cosine", "dotproduct", "hypersphere_euclidean"])
def test_dense_scores_on_tensor_cores(cuda_lib, kind):
    """SupportBank.scores: the kernel(x, y) matrix through the TMA + tcgen05 mainloop (emit epilogue), on an
    UNSORTED support (columns must come back in the caller's order), ragged N, B across a tile edge."""
    from nwhead_b200 import SupportBank

    rng = np.random.default_rng(8)
    B, N, d, C = 150, 1001, 72, 11
    y = rng.integers(0, C, N).astype(np.int64)
    s = rng.normal(size=(N, d)).astype(np.float32)
    q = rng.normal(size=(B, d)).astype(np.float32)
    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), C, kind, "bf16x3")
    got = bank.scores(torch.from_numpy(q).to(DEV)).cpu().numpy()
    ref = O.pairwise_scores(q, s, kind)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 5e-4 * max(1.0, np.abs(ref).max())
    plain = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), C, kind, "bf16")
    got1 = plain.scores(torch.from_numpy(q).to(DEV)).cpu().numpy()
    assert np.abs(got1 - ref).max() < 0.08 * max(1.0, np.abs(ref).max() / 8)


@pytest.mark.parametrize("case", [(20, 50, 512, 96, "euclidean"), (20, 51, 512, 300, "euclidean"),
                                  (12, 77, 128, 300, "cosine"), (6, 300, 2048, 130, "euclidean")])
def test_support_influence_from_features(cuda_lib, case):
    """Influence computed from features (two tensor-core passes) == the reference formula fed with the exact
    softmax weights (oracle), BASELINE config 5 shapes scaled down; single CTAs and CTA pairs, ragged last tile."""
    from nwhead_b200 import SupportBank

    C, per, d, B, kind = case
    q, s, y, qy = clustered_features(C, per, d, B, seed=21)
    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), C, kind, "bf16x3")
    got = bank.support_influence(torch.from_numpy(q).to(DEV), torch.from_numpy(qy).to(DEV)).cpu().numpy()
    sc = O.pairwise_scores(q, s, kind)
    w = np.exp(sc - sc.max(1, keepdims=True))
    w /= w.sum(1, keepdims=True)
    P = np.zeros((B, C))
    np.add.at(P.T, y, w.T)
    ref = O.support_influence(P, qy, w, y)
    ok = np.isfinite(ref) & (np.abs(ref) > 1e-12)
    assert np.array_equal(np.sign(got[ok]), np.sign(ref[ok]))
    assert np.allclose(got[ok], ref[ok], rtol=5e-3, atol=1e-7)


def test_tensor_core_topk_matches_exact_ranking(cuda_lib):
    """SupportBank.topk (emit scores + bitonic ranking) on a shuffled support: with the 3-product operands the
    ranking equals the exact float64 ranking on data whose score gaps exceed the operand rounding."""
    from nwhead_b200 import SupportBank

    rng = np.random.default_rng(12)
    B, N, d, C, k = 70, 3000, 64, 30, 10
    y = rng.integers(0, C, N).astype(np.int64)
    s = rng.normal(size=(N, d)).astype(np.float32)
    q = rng.normal(size=(B, d)).astype(np.float32)
    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), C, "euclidean", "bf16x3")
    got = bank.topk(torch.from_numpy(q).to(DEV), k, query_chunk=32).cpu().numpy()
    sc = O.pairwise_scores(q, s, "euclidean")
    ref = np.argsort(-sc, axis=1, kind="stable")[:, :k + 1]
    gaps = -np.diff(np.take_along_axis(sc, ref, 1), axis=1)
    clear = (gaps > 1e-3).all(axis=1)  # rows whose first k+1 neighbours are separated by more than the rounding
    assert clear.mean() > 0.5
    assert np.array_equal(got[clear], ref[clear, :k])


def test_cuda_graph_replay_matches_eager(cuda_lib):
    from nwhead_b200 import SupportBank

    q, s, y, _ = clustered_features(200, 29, 512, 8, seed=1)  # BASELINE config 1 head shape
    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), 200, "euclidean", "bf16")
    qd = torch.from_numpy(q).to(DEV)
    eager = bank.forward(qd)
    fwd = bank.graphed(8)
    assert torch.equal(fwd(qd), eager)
    q2 = torch.roll(qd, 3, dims=0)
    assert torch.equal(fwd(q2), torch.roll(eager, 3, dims=0))
    with pytest.raises(ValueError):
        fwd(qd[:4])


def test_empty_and_degenerate_inputs(cuda_lib):
    import nwhead_b200
    from nwhead_b200 import SupportBank

    q, s, y, _ = clustered_features(5, 10, 32, 4, seed=2)
    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), 5, "euclidean", "bf16")
    empty = bank.forward(torch.empty((0, 32), device=DEV))
    assert tuple(empty.shape) == (0, 5)
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 5)
    out = head(torch.empty((0, 32), device=DEV), torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV))
    assert tuple(out.shape) == (0, 5)
    with pytest.raises(ValueError):
        head(torch.from_numpy(q).to(DEV), torch.empty((0, 32), device=DEV), torch.empty((0,), dtype=torch.int64, device=DEV))
    # a single query against a single support row: P = 1 for its class, log(1e-12) elsewhere
    one = head(torch.from_numpy(q[:1]).to(DEV), torch.from_numpy(s[:1]).to(DEV), torch.from_numpy(y[:1]).to(DEV)).cpu().numpy()
    assert abs(one[0, y[0]]) < 1e-6 and np.allclose(np.delete(one[0], y[0]), np.log(np.float32(1e-12)))
    # one query, one class in the bank (tensor-core path, N > 25)
    yy = np.zeros(40, np.int64)
    b1 = SupportBank.build(torch.from_numpy(s[:40]).to(DEV), torch.from_numpy(yy).to(DEV), 3, "cosine", "bf16")
    o1 = b1.forward(torch.from_numpy(q[:1]).to(DEV)).cpu().numpy()
    assert abs(o1[0, 0]) < 1e-6 and np.allclose(o1[0, 1:], np.log(np.float32(1e-12)))


@pytest.mark.parametrize("shape", [(300, 5000, 4, 64), (5000, 40, 80, 32)])
def test_many_classes_and_many_queries(cuda_lib, shape):
    """(B, C, per_class, d): a 5000-class table with 4 supports per class (flush-dense), and 5000 queries (40 query
    groups over few support tiles)."""
    B, C, per, d = shape
    q, s, y, _ = clustered_features(C, per, d, B, seed=C + B, spread=0.3)
    out = run_bank(q, s, y, C, "euclidean", "bf16x3")
    ref = O.nw_forward(q, s, y, C, "euclidean")
    assert np.abs(np.exp(out.cpu().numpy()) - np.exp(ref)).max() < 5e-5


def test_head_caches_the_bank_of_unchanged_support_tensors(cuda_lib):
    import nwhead_b200

    q, s, y, _ = clustered_features(6, 20, 48, 9, seed=5)
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 6)
    sx, sy, qx = torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), torch.from_numpy(q).to(DEV)
    with torch.no_grad():
        a = head(qx, sx, sy)
        bank = head._bank_cache[-1][-1]
        b = head(qx, sx, sy)
        assert head._bank_cache[-1][-1] is bank and torch.equal(a, b)     # same tensors -> cached bank
        sx.mul_(1.5)                                                       # in-place change -> rebuilt
        c = head(qx, sx, sy)
        assert head._bank_cache[-1][-1] is not bank
        assert_head_parity(c, O.nw_forward(q, s * 1.5, y, 6, "euclidean"))
        # a NEW support tensor that lands on the freed memory of an old one (same address, shape and version, as
        # knn mode produces batch after batch) must not be served the old bank
        for scale in (1.0, 3.0, 0.5):
            tmp = torch.from_numpy(s * scale).to(DEV)
            out = head(qx, tmp, sy)
            assert_head_parity(out, O.nw_forward(q, s * scale, y, 6, "euclidean"))
            del tmp


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_exact_topk_via_block_candidates(cuda_lib, precision):
    """SupportBank.topk_exact == dense fp32 ranking (nw_direct_scores + nw_rank_rows), bit for bit, on a shuffled
    support with ragged N; also with a candidate budget so small that some queries take the dense fallback."""
    from nwhead_b200 import SupportBank
    from nwhead_b200.kernel import dense_scores
    from nwhead_b200.utils import rank_rows

    rng = np.random.default_rng(31)
    B, N, d, C, k = 45, 20011, 64, 50, 10
    y = rng.integers(0, C, N).astype(np.int64)
    mu = rng.normal(size=(C, d)) * 2
    s = (mu[y] + rng.normal(size=(N, d))).astype(np.float32)
    q = (mu[rng.integers(0, C, B)] + rng.normal(size=(B, d))).astype(np.float32)
    sx, qx = torch.from_numpy(s).to(DEV), torch.from_numpy(q).to(DEV)
    bank = SupportBank.build(sx, torch.from_numpy(y).to(DEV), C, "euclidean", precision)
    want = rank_rows(dense_scores("euclidean", qx, sx), k)
    got = bank.topk_exact(qx, k, sx, query_chunk=16)
    assert torch.equal(got, want)
    ref = O.topk_neighbors(q, s, k)
    assert np.array_equal(got.cpu().numpy(), ref)            # and equal to the float64 oracle on this data
    assert torch.equal(bank.topk_exact(qx, k, sx, max_blocks=1), want)   # forces the dense fallback for most rows
    assert bank.last_topk_path["dense"] > 0
    # adversarial: thousands of rows within 1e-3 of each other around every query (far below bf16 resolution),
    # spread over all blocks by the label shuffle -> the certificate has to widen the search or fall back
    s2 = s.copy()
    near = rng.choice(N, 4000, replace=False)
    s2[near] = q[rng.integers(0, B, 4000)] + rng.normal(size=(4000, d)).astype(np.float32) * 1e-3
    sx2 = torch.from_numpy(s2).to(DEV)
    bank2 = SupportBank.build(sx2, torch.from_numpy(y).to(DEV), C, "euclidean", precision)
    assert torch.equal(bank2.topk_exact(qx, k, sx2), rank_rows(dense_scores("euclidean", qx, sx2), k))
    res = bank.rounding_residual(sx)                          # measured residuals obey the a-priori bf16 bound
    cn = (sx - bank.center).norm(dim=1)
    assert (res <= cn * (2.0 ** -8 if precision == "bf16" else 2.0 ** -16)).all() and res.max() > 0
    bb, _ = bank.block_best(qx)
    dense = dense_scores("euclidean", qx, sx)
    assert bb.shape == (B, (N + 63) // 64)
    blk_max = torch.full((B, bb.shape[1] * 64), float("-inf"), device=DEV)
    blk_max[:, :N] = dense if bank.perm is None else dense[:, bank.perm]
    assert (bb - blk_max.view(B, -1, 64).amax(2)).abs().max().item() < (0.05 if precision == "bf16" else 1e-3)


def test_small_bank_graph_replay_is_transparent(cuda_lib):
    """NWHead replays a CUDA graph for launch-bound banks once a batch size has been seen twice
    (SupportBank.forward_auto).  Results must be bit-identical to the eager path, independent tensors (no aliasing
    of the graph's static output), and a changed batch size or NW_B200_GRAPHS=0 must keep working."""
    import os

    import nwhead_b200

    q, s, y, _ = clustered_features(10, 30, 64, 24, seed=21)
    bank = nwhead_b200.SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), 10, "euclidean", "bf16")
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 10)
    qt = torch.from_numpy(q).to(DEV)
    want = bank.forward(qt)
    with torch.no_grad():
        outs = [head(qt, bank) for _ in range(4)]          # call 3 onwards replays the captured graph
        assert isinstance(bank._graph_cache[(24, 1.0)], nwhead_b200.bank.GraphedForward)
        other = head(qt.flip(0), bank)                      # same shape, other data: must not disturb earlier results
        small = head(qt[:5], bank)                          # another batch size: eager again, then its own graph
        small2 = head(qt[:5], bank)
        small3 = head(qt[:5], bank)
    for o in outs:
        assert torch.equal(o, want)
    assert outs[2].data_ptr() != outs[3].data_ptr()
    assert torch.equal(other, want.flip(0))
    assert torch.allclose(small, want[:5], atol=1e-5) and torch.equal(small2, small) and torch.equal(small3, small)
    os.environ["NW_B200_GRAPHS"] = "0"
    try:
        with torch.no_grad():
            assert torch.equal(head(qt, bank), want)
    finally:
        del os.environ["NW_B200_GRAPHS"]


@pytest.mark.parametrize("kind", ["euclidean", "cosine"])
@pytest.mark.parametrize("shape", [(300, 9, 400, 128), (40, 9, 400, 512), (300, 9, 400, 1024), (300, 5, 700, 2048)])
def test_tile_metadata_paths_agree(cuda_lib, shape, kind):
    """Interior tiles get their |s|^2 / labels from the TMA producer (bulk copies, 16-byte aligned arrays); edge
    tiles and unaligned arrays are staged by the epilogue sets.  Both must give the same bits, for 1 / 2 / 4
    epilogue sets and for the single-CTA schedule."""
    from nwhead_b200 import SupportBank

    B, C, per, d = shape
    q, s, y, _ = clustered_features(C, per, d, B, seed=17 + d)
    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), C, kind, "bf16")
    qd = torch.from_numpy(q).to(DEV)
    aligned = bank.forward(qd).clone()
    assert bank.sqnorm.data_ptr() % 16 == 0 and bank.labels.data_ptr() % 16 == 0
    n = bank.labels.numel()
    sq = torch.empty(n + 1, dtype=bank.sqnorm.dtype, device=DEV)
    lab = torch.empty(n + 1, dtype=bank.labels.dtype, device=DEV)
    sq[1:].copy_(bank.sqnorm)
    lab[1:].copy_(bank.labels)
    bank.sqnorm, bank.labels = sq[1:], lab[1:]  # same values, 4 bytes off alignment -> staged path for every tile
    assert bank.labels.data_ptr() % 16 == 4
    staged = bank.forward(qd)
    torch.cuda.synchronize()
    assert torch.equal(aligned, staged)
    assert_head_parity(aligned, O.nw_forward(q, s, y, C, kind))


def test_one_support_per_class_shortcut_and_its_near_miss(cuda_lib):
    """N == C with every class present once: the class table is the score matrix (dense-score epilogue).  N == C with
    one class twice and one absent must NOT take that route (absent class: exactly log(1e-12))."""
    from nwhead_b200 import SupportBank

    C, d, B = 300, 192, 70
    q, s, y, _ = clustered_features(C, 1, d, B, seed=5, spread=0.25)
    rng = np.random.default_rng(0)
    order = rng.permutation(C)  # unsorted support: the bank sorts it
    bank = SupportBank.build(torch.from_numpy(s[order]).to(DEV), torch.from_numpy(y[order]).to(DEV), C, "euclidean", "bf16")
    assert bank.identity_classes
    out = bank.forward(torch.from_numpy(q).to(DEV))
    assert_head_parity(out, O.nw_forward(q, s[order], y[order], C, "euclidean"))
    y2 = y.copy()
    y2[7] = 8  # class 7 absent, class 8 twice
    bank2 = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y2).to(DEV), C, "euclidean", "bf16")
    assert not bank2.identity_classes
    out2 = bank2.forward(torch.from_numpy(q).to(DEV))
    ref2 = O.nw_forward(q, s, y2, C, "euclidean")
    assert_head_parity(out2, ref2)
    assert np.allclose(out2[:, 7].cpu().numpy(), np.log(1e-12), atol=1e-5)

"""GPU: tensor-core forward + backward of NWHead.forward against a large shared support
(nwhead_b200/backward.py: nw_backward_coefficients + nw_dense_products) vs the float64 oracle's closed-form
gradients (oracle/nw_oracle.py::nw_backward, which restates the autograd of nwhead/nw.py:266-289)."""
import numpy as np
import pytest
import torch

from oracle import nw_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# gradients from bf16 operands and bf16 coefficients: max-abs error relative to the gradient's max-abs
GRAD_TOL = 3e-2


def _data(B, N, d, C, seed, sorted_labels):
    rng = np.random.default_rng(seed)
    mu = rng.normal(size=(C, d)) * 0.6
    y = rng.integers(0, C, N).astype(np.int64)
    if sorted_labels:
        y.sort()
    s = np.maximum(mu[y] + rng.normal(size=(N, d)) + 0.5, 0).astype(np.float32)
    qy = rng.integers(0, C, B)
    q = np.maximum(mu[qy] + rng.normal(size=(B, d)) + 0.5, 0).astype(np.float32)
    g = rng.normal(size=(B, C)).astype(np.float32)
    return q, s, y, g


def test_dense_products_match_matmul(cuda_lib):
    """a @ b^t on the tensor cores, with and without split-K (skinny problems), ragged row counts and a K that is
    not a multiple of the slice size."""
    from nwhead_b200.backward import dense_products

    g = torch.Generator(device=DEV).manual_seed(3)
    for n_a, n_b, kb in [(300, 130, 3), (70, 64, 75), (520, 700, 21), (9, 2048, 130)]:
        a = torch.randn(kb, n_a, 64, generator=g, device=DEV).bfloat16()
        b = torch.randn(kb, n_b, 64, generator=g, device=DEV).bfloat16()
        got = dense_products(a, b)
        ref = a.permute(1, 0, 2).reshape(n_a, -1).double() @ b.permute(1, 0, 2).reshape(n_b, -1).double().t()
        assert got.shape == (n_a, n_b)
        assert (got.double() - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())


def test_transpose_operand_is_the_transpose(cuda_lib):
    from nwhead_b200.backward import transpose_operand

    x = torch.randn(3, 100, 64, device=DEV).bfloat16()  # 100 rows of 192 elements
    t = transpose_operand(x)                              # 192 rows of 128 (100 + zero padding) elements
    assert t.shape == (2, 192, 64)
    rows = x.permute(1, 0, 2).reshape(100, 192)
    cols = t.permute(1, 0, 2).reshape(192, 128)
    assert torch.equal(cols[:, :100], rows.t()) and not cols[:, 100:].any()


@pytest.mark.parametrize("kind", O.KERNEL_KINDS)
@pytest.mark.parametrize("shape", [(200, 3000, 128, 30, True), (130, 5000, 100, 11, False), (70, 2000, 64, 300, False),
                                   (300, 1100, 192, 7, True)])
def test_tensor_backward_matches_oracle(cuda_lib, shape, kind):
    """Forward within the north-star tolerance, gradients within GRAD_TOL, for every kernel: class-sorted and
    unsorted supports (gradient rows come back in the caller's order), d not a multiple of 64, query counts across
    the 128/256-row tile edges, more classes than rows per class (per-column table lookups)."""
    import nwhead_b200

    B, N, d, C, sorted_labels = shape
    q, s, y, g = _data(B, N, d, C, B + N, sorted_labels)
    if kind == "dotproduct":
        q, s = q * 0.1, s * 0.1
    kern = nwhead_b200.get_kernel(kind).to(DEV)
    head = nwhead_b200.NWHead(kern, C, backward_path="tensor")
    qt = torch.from_numpy(q).to(DEV).requires_grad_(True)
    st = torch.from_numpy(s).to(DEV).requires_grad_(True)
    logp = head(qt, st, torch.from_numpy(y).to(DEV))
    assert logp.grad_fn is not None
    (logp * torch.from_numpy(g).to(DEV)).sum().backward()
    ref = O.nw_forward(q, s, y, C, kind)
    assert np.abs(np.exp(logp.detach().cpu().numpy().astype(np.float64)) - np.exp(ref)).max() < 1e-3
    res = O.nw_backward(q, s, y, C, g, kind)
    for name, got, want in (("grad_q", qt.grad, res[0]), ("grad_s", st.grad, res[1])):
        got = got.cpu().numpy().astype(np.float64)
        assert np.isfinite(got).all()
        err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
        assert err < GRAD_TOL, f"{kind} {name}: max-abs error {err:.3e} of the gradient's max-abs"
    if kind == "clip":
        assert abs(float(kern.logit_scale.grad) - res[2]) < GRAD_TOL * max(1.0, abs(res[2]))


def test_only_the_requested_gradients_are_computed(cuda_lib):
    import nwhead_b200

    q, s, y, g = _data(96, 1500, 64, 12, 5, True)
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 12, backward_path="tensor")
    qt = torch.from_numpy(q).to(DEV).requires_grad_(True)
    st = torch.from_numpy(s).to(DEV)  # no gradient for the support
    (head(qt, st, torch.from_numpy(y).to(DEV)) * torch.from_numpy(g).to(DEV)).sum().backward()
    want = O.nw_backward(q, s, y, 12, g, "euclidean")[0]
    assert np.abs(qt.grad.cpu().numpy() - want).max() < GRAD_TOL * np.abs(want).max()
    assert st.grad is None


def test_support_gradient_alone(cuda_lib):
    """Only the support is trained: the coefficients are computed with the operand roles swapped (rows = supports),
    no W and no grad_q products; unsorted labels, a batch that is not a multiple of 4 (scalar table loads) and one
    that is (16-byte loads)."""
    import nwhead_b200

    for B in (70, 200):
        q, s, y, g = _data(B, 2500, 96, 9, B, False)
        head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 9, backward_path="tensor")
        qt = torch.from_numpy(q).to(DEV)
        st = torch.from_numpy(s).to(DEV).requires_grad_(True)
        (head(qt, st, torch.from_numpy(y).to(DEV)) * torch.from_numpy(g).to(DEV)).sum().backward()
        want = O.nw_backward(q, s, y, 9, g, "euclidean")[1]
        assert np.abs(st.grad.cpu().numpy() - want).max() < GRAD_TOL * np.abs(want).max()


def test_zero_distance_contributes_no_gradient(cuda_lib):
    """A query that coincides with a support row: torch.cdist's backward yields 0 for that pair (and so does the
    direct path); the bf16 recompute sees d2 <= 0 there and must not produce inf / NaN."""
    import nwhead_b200

    q, s, y, g = _data(64, 1024, 64, 8, 11, True)
    q[0] = s[17]
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 8, backward_path="tensor")
    qt = torch.from_numpy(q).to(DEV).requires_grad_(True)
    st = torch.from_numpy(s).to(DEV).requires_grad_(True)
    (head(qt, st, torch.from_numpy(y).to(DEV)) * torch.from_numpy(g).to(DEV)).sum().backward()
    assert torch.isfinite(qt.grad).all() and torch.isfinite(st.grad).all()


def test_auto_routing_and_label_errors(cuda_lib):
    """backward_path='auto': a big batch against a big shared support differentiates on the tensor cores, a small
    problem stays on the direct fp32 kernels; an out-of-range label raises like F.one_hot (nwhead/nw.py:276)."""
    import nwhead_b200

    C, d = 16, 32
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), C)
    g = torch.Generator(device=DEV).manual_seed(1)
    s = torch.randn(1 << 18, d, generator=g, device=DEV)
    y = torch.randint(0, C, (1 << 18,), generator=g, device=DEV)
    q = torch.randn(64, d, generator=g, device=DEV, requires_grad=True)
    out = head(q, s, y)
    assert type(out.grad_fn).__name__ == "NWTensorFunctionBackward"
    out.sum().backward()
    assert torch.isfinite(q.grad).all()
    small = head(q[:8], s[:100], y[:100])
    assert type(small.grad_fn).__name__ != "NWTensorFunctionBackward"
    # the fixed support's bank (and its transposed copy) is built once and reused
    assert head(q, s, y).grad_fn is not None and len(head._bank_cache) == 1
    bad = y.clone()
    bad[5] = C
    strict = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), C, backward_path="tensor")
    with pytest.raises(RuntimeError):
        strict(q, s, bad)
    with pytest.raises(ValueError):
        nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), C, backward_path="nope")


def test_tensor_backward_on_tiny_and_ragged_problems(cuda_lib):
    """Forced tensor path far below its intended sizes: fewer than 8 queries, fewer supports than one tile, one
    support per class (the identity-class bank of cluster mode)."""
    import nwhead_b200

    for B, N, d, C in [(3, 40, 20, 5), (1, 7, 8, 7), (9, 300, 70, 300)]:
        rng = np.random.default_rng(N)
        q = rng.normal(size=(B, d)).astype(np.float32)
        s = rng.normal(size=(N, d)).astype(np.float32)
        y = (np.arange(N) % C).astype(np.int64) if N != C else np.arange(N).astype(np.int64)
        g = rng.normal(size=(B, C)).astype(np.float32)
        head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), C, backward_path="tensor")
        qt = torch.from_numpy(q).to(DEV).requires_grad_(True)
        st = torch.from_numpy(s).to(DEV).requires_grad_(True)
        (head(qt, st, torch.from_numpy(y).to(DEV)) * torch.from_numpy(g).to(DEV)).sum().backward()
        res = O.nw_backward(q, s, y, C, g, "euclidean")
        for got, want in ((qt.grad, res[0]), (st.grad, res[1])):
            err = np.abs(got.cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-30)
            assert err < GRAD_TOL, (B, N, d, C, err)


@pytest.mark.parametrize("kind", [k for k in O.KERNEL_KINDS if k != "dotproduct"])
@pytest.mark.parametrize("case", ["mm_medium", "wide"])
def test_tensor_backward_matches_the_reference_autograd(cuda_lib, golden_head, case, kind):
    """The fixtures of tests/golden/head.npz hold the REFERENCE's own log-probs and autograd gradients
    (oracle/gen_golden.py ran the unmodified nwhead/nw.py:266-289 under torch autograd): the tensor-core forward +
    backward against them, unsorted / duplicate / absent labels and a coincident query/support pair.  (Not the plain
    dot product: the fixtures' unnormalised scores are ~50 in magnitude, which single bf16 products resolve to ~0.1 —
    the forward tests run that kernel with the 3-product split; test_tensor_backward_matches_oracle covers it at a
    scale bf16 resolves.)"""
    import nwhead_b200

    g = golden_head
    C = int(g[f"{case}/C"])
    kern = nwhead_b200.get_kernel(kind).to(DEV)
    head = nwhead_b200.NWHead(kern, C, backward_path="tensor")
    q = torch.from_numpy(g[f"{case}/q"]).to(DEV).requires_grad_(True)
    s = torch.from_numpy(g[f"{case}/s"]).to(DEV).requires_grad_(True)
    y = torch.from_numpy(g[f"{case}/y"]).to(DEV)
    G = torch.from_numpy(g[f"{case}/G"]).to(DEV)
    logp = head(q, s, y)
    (logp * G).sum().backward()
    assert np.abs(np.exp(logp.detach().cpu().numpy()) - np.exp(g[f"{case}/{kind}/logp"])).max() < 1e-3
    gq, gs = g[f"{case}/{kind}/gq"], g[f"{case}/{kind}/gs"]
    scale = max(np.abs(gq).max(), np.abs(gs).max())
    assert np.abs(q.grad.cpu().numpy() - gq).max() / scale < GRAD_TOL
    assert np.abs(s.grad.cpu().numpy() - gs).max() / scale < GRAD_TOL
    if kind == "clip":
        gl = float(g[f"{case}/{kind}/glogit"])
        assert abs(float(kern.logit_scale.grad) - gl) < GRAD_TOL * max(1.0, abs(gl))

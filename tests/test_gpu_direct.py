"""GPU: direct fp32 path (episodic training): forward + closed-form backward against the reference's
autograd (tests/golden/head.npz) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import nw_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("kind", O.KERNEL_KINDS)
@pytest.mark.parametrize("case", ["exact_small", "mm_medium", "wide", "per_query_3d"])
def test_forward_backward_matches_reference(cuda_lib, golden_head, case, kind):
    import nwhead_b200

    g = golden_head
    C = int(g[f"{case}/C"])
    kern = nwhead_b200.get_kernel(kind).to(DEV)
    head = nwhead_b200.NWHead(kern, C)
    q = torch.from_numpy(g[f"{case}/q"]).to(DEV).requires_grad_(True)
    s = torch.from_numpy(g[f"{case}/s"]).to(DEV).requires_grad_(True)
    y = torch.from_numpy(g[f"{case}/y"]).to(DEV)
    G = torch.from_numpy(g[f"{case}/G"]).to(DEV)
    logp = head(q, s, y)
    (logp * G).sum().backward()
    ref = g[f"{case}/{kind}/logp"]
    # fp32 exact-difference path: tolerance 2e-5 on log-probs that are not at the 1e-12 floor
    got = logp.detach().cpu().numpy()
    assert np.abs(np.exp(got) - np.exp(ref)).max() < 2e-5
    gq, gs = g[f"{case}/{kind}/gq"], g[f"{case}/{kind}/gs"]
    scale = max(1.0, np.abs(gq).max(), np.abs(gs).max())
    assert np.abs(q.grad.cpu().numpy() - gq).max() / scale < 2e-4
    assert np.abs(s.grad.cpu().numpy() - gs).max() / scale < 2e-4
    if kind == "clip":
        gl = float(g[f"{case}/{kind}/glogit"])
        assert abs(float(kern.logit_scale.grad) - gl) / max(1.0, abs(gl)) < 2e-4


def test_zero_distance_gradient_is_finite(cuda_lib, golden_head):
    """Query 0 of the 2-D fixtures equals support row 3: cdist backward gives 0 there, not NaN."""
    import nwhead_b200

    g = golden_head
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), int(g["exact_small/C"]))
    q = torch.from_numpy(g["exact_small/q"]).to(DEV).requires_grad_(True)
    s = torch.from_numpy(g["exact_small/s"]).to(DEV).requires_grad_(True)
    assert torch.equal(q[0].detach(), s[3].detach())
    head(q, s, torch.from_numpy(g["exact_small/y"]).to(DEV)).sum().backward()
    assert torch.isfinite(q.grad).all() and torch.isfinite(s.grad).all()


def test_config2_shape(cuda_lib):
    """BASELINE config 2: B=8, n_way=10, n_shot=1, d=512, C=200, NLL loss."""
    import nwhead_b200

    rng = np.random.default_rng(2)
    sy = rng.choice(200, 10, replace=False)
    qy = sy[rng.integers(0, 10, 8)]
    s = np.maximum(rng.normal(size=(10, 512)) + 0.5, 0).astype(np.float32)
    q = np.maximum(rng.normal(size=(8, 512)) + 0.5, 0).astype(np.float32)
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 200)
    qt = torch.from_numpy(q).to(DEV).requires_grad_(True)
    st = torch.from_numpy(s).to(DEV).requires_grad_(True)
    logp = head(qt, st, torch.from_numpy(sy).to(DEV))
    loss = torch.nn.functional.nll_loss(logp, torch.from_numpy(qy).to(DEV))
    loss.backward()
    G = np.zeros((8, 200))
    G[np.arange(8), qy] = -1.0 / 8
    gq, gs = O.nw_backward(q, s, sy, 200, G, "euclidean")
    assert np.abs(logp.detach().cpu().numpy() - O.nw_forward(q, s, sy, 200, "euclidean")).max() < 1e-4
    assert np.abs(qt.grad.cpu().numpy() - gq).max() < 1e-6 + 1e-4 * np.abs(gq).max()
    assert np.abs(st.grad.cpu().numpy() - gs).max() < 1e-6 + 1e-4 * np.abs(gs).max()


def test_kernel_modules_return_dense_scores(cuda_lib, golden_head):
    import nwhead_b200

    g = golden_head
    q, s = g["mm_medium/q"], g["mm_medium/s"]
    for kind in O.KERNEL_KINDS:
        k = nwhead_b200.get_kernel(kind).to(DEV)
        got = k(torch.from_numpy(q).to(DEV), torch.from_numpy(s).to(DEV)).cpu().numpy()
        ref = O.pairwise_scores(q, s, kind)
        assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
        got3 = k(torch.from_numpy(q[:, None]).to(DEV), torch.from_numpy(np.broadcast_to(s, (5,) + s.shape).copy()).to(DEV))
        assert got3.shape == (5, 1, 60) and np.abs(got3.cpu().numpy()[:, 0] - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("kind", ["euclidean", "cosine", "clip", "hypersphere_euclidean", "dotproduct"])
@pytest.mark.parametrize("shape", [(16, 1500, 48, 20, False), (300, 12, 40, 9, False), (3, 1100, 24, 7, True),
                                   (12, 5000, 40, 300, False), (20, 4500, 300, 11, False)])
def test_large_direct_path_against_oracle(cuda_lib, shape, kind):
    """Shapes that leave the fused small-support kernels (N > 1024 supports, or more than 256 queries):
    the generic scores / aggregate / coefficient / gradient kernels, forward and backward vs the float64 oracle.
    N > 1024 shared supports take the split grad_q reduction (several chunks, more than 8 queries, d below and above
    one 256-column tile) and the multi-row grad_s kernel; N > 4096 the shared-memory class bins of the aggregation."""
    import nwhead_b200

    B, N, d, C, batched = shape
    rng = np.random.default_rng(B * 7 + N)
    q = rng.normal(size=(B, d)).astype(np.float32)
    if kind == "dotproduct":
        q *= 0.2
    G = rng.normal(size=(B, C)).astype(np.float32)
    kern = nwhead_b200.get_kernel(kind).to(DEV)
    head = nwhead_b200.NWHead(kern, C)
    if batched:
        s = rng.normal(size=(B, N, d)).astype(np.float32)
        y = rng.integers(0, C, (B, N)).astype(np.int64)
    else:
        s = rng.normal(size=(N, d)).astype(np.float32)
        y = rng.integers(0, C, N).astype(np.int64)
    qt = torch.from_numpy(q).to(DEV).requires_grad_(True)
    st = torch.from_numpy(s).to(DEV).requires_grad_(True)
    logp = head(qt, st, torch.from_numpy(y).to(DEV))
    (logp * torch.from_numpy(G).to(DEV)).sum().backward()
    ref = O.nw_forward(q, s, y, C, kind)
    assert np.abs(np.exp(logp.detach().cpu().numpy()) - np.exp(ref)).max() < 5e-5
    if not batched:
        res = O.nw_backward(q, s, y, C, G, kind)
        scale = max(1.0, np.abs(res[0]).max(), np.abs(res[1]).max())
        assert np.abs(qt.grad.cpu().numpy() - res[0]).max() / scale < 5e-4
        assert np.abs(st.grad.cpu().numpy() - res[1]).max() / scale < 5e-4
        if kind == "clip":
            assert abs(float(kern.logit_scale.grad) - res[2]) / max(1.0, abs(res[2])) < 5e-4
    else:
        # per-query supports: check grad_q against the oracle evaluated query by query
        for b in range(B):
            gq, _ = O.nw_backward(q[b:b + 1], s[b], y[b], C, G[b:b + 1], kind)[:2]
            assert np.abs(qt.grad[b].cpu().numpy() - gq[0]).max() < 5e-4 * max(1.0, np.abs(gq).max())
        assert torch.isfinite(st.grad).all()

"""Small-shape driver for compute-sanitizer (memcheck / racecheck / synccheck) over every variant of the fused
forward kernel and the direct kernels.  Run through tools/run_sanitizer.sh on a GPU box; results are compared with
the float64 oracle so that a sanitizer-clean run is also a correct one."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import nwhead_b200  # noqa: E402
from nwhead_b200 import SupportBank, _abi  # noqa: E402
from oracle import nw_oracle as O  # noqa: E402

DEV = "cuda:0"


def data(n_classes, per, d, b, seed):
    rng = np.random.default_rng(seed)
    y = np.repeat(np.arange(n_classes), per).astype(np.int64)
    mu = rng.normal(size=(n_classes, d)) * 0.6
    s = np.maximum(mu[y] + rng.normal(size=(len(y), d)) + 0.5, 0).astype(np.float32)
    q = np.maximum(mu[rng.integers(0, n_classes, b)] + rng.normal(size=(b, d)) + 0.5, 0).astype(np.float32)
    return q, s, y


def check(name, got, ref, tol=1e-3):
    err = np.abs(np.exp(got) - np.exp(ref)).max()
    print(f"{name}: class-probability error {err:.2e}", flush=True)
    assert err < tol, name


def main():
    lib = _abi.load()
    _abi.check(lib.nw_device_check(), "nw_device_check")
    cases = [
        ("1-CTA tiles, 1 epilogue set (d=2048 rows: kblocks > 16)", 12, 90, 1088, 40, "euclidean"),
        ("1-CTA tiles, 2 epilogue sets (d=128)", 12, 90, 128, 40, "euclidean"),
        ("CTA pair, 2 epilogue sets (d=256), ragged batch", 16, 70, 256, 300, "euclidean"),
        ("CTA pair, 1 epilogue set (d=1088)", 10, 60, 1088, 260, "euclidean"),
        ("CTA pair, LINEAR epilogue (cosine)", 16, 70, 192, 300, "cosine"),
    ]
    for name, c, per, d, b, kind in cases:
        q, s, y = data(c, per, d, b, 1)
        bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), c, kind, "bf16")
        head = nwhead_b200.NWHead(nwhead_b200.get_kernel(kind), c)
        with torch.no_grad():
            got = head(torch.from_numpy(q).to(DEV), bank).cpu().numpy()
        check(name, got, O.nw_forward(q, s, y, c, kind), 2e-3)
    # many chunks (side buffer + merge kernel), multi-table routes (the bank-sharded exchange, here with both
    # tables on this GPU: route 1 = by query row, route 2 = every table)
    c, per, d, b = 24, 700, 128, 300
    q, s, y = data(c, per, d, b, 2)
    bank = SupportBank.build(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), c, "euclidean", "bf16")
    qt = torch.from_numpy(q).to(DEV)
    want = bank.class_lse(qt)
    qb, qs = bank.prepare_queries(qt)
    for rows_per_table in (0, 150):
        t = [torch.full((b, c), float("-inf"), device=DEV) for _ in range(2)]
        ptrs = (ctypes.c_void_p * 2)(*[x.data_ptr() for x in t])
        bank.class_lse_prepared(qb, qs, tables=ptrs, rows_per_table=rows_per_table)
        torch.cuda.synchronize()
        if rows_per_table:
            got = torch.cat((t[0][:150], t[1][150:]))
        else:
            got = t[0]
            assert torch.equal(t[0], t[1])
        assert (got - want).abs().max().item() < 2e-5, rows_per_table
        print(f"multi-table route rows_per_table={rows_per_table}: ok", flush=True)
    # emit modes
    sc = bank.scores(qt, source_order=False).cpu().numpy()
    ref_sc = O.pairwise_scores(q, s, "euclidean")
    print("emit scores: max err", np.abs(sc - ref_sc).max(), flush=True)
    assert np.abs(sc - ref_sc).max() < 0.05
    qy = torch.from_numpy(y[:b].copy()).to(DEV)
    infl = bank.support_influence(qt, qy, source_order=False)
    assert torch.isfinite(infl[infl == infl]).any()
    best, _ = bank.block_best(qt)
    assert best.shape == (b, (len(bank) + 63) // 64)
    idx = bank.topk_exact(qt[:64], 5, torch.from_numpy(s).to(DEV))
    assert np.array_equal(idx.cpu().numpy(), O.topk_neighbors(q[:64], s, 5))
    print("emit influence / block-best / topk_exact: ok", flush=True)
    # direct fp32 path forward + backward (episodic shape and the generic large-support route)
    for n_s in (10, 1500):
        head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), c)
        q8 = qt[:8].clone().requires_grad_(True)
        sx = torch.from_numpy(s[:: max(1, len(s) // n_s)][:n_s].copy()).to(DEV).requires_grad_(True)
        sy = torch.from_numpy(y[:: max(1, len(s) // n_s)][:n_s].copy()).to(DEV)
        out = head(q8, sx, sy)
        out.sum().backward()
        torch.cuda.synchronize()
        assert torch.isfinite(q8.grad).all() and torch.isfinite(sx.grad).all()
    print("direct forward/backward: ok", flush=True)
    # tensor-core backward: coefficient emit in both orientations, split-K products, fused transposed grad_s launch
    head_t = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), c, backward_path="tensor")
    for need_q, need_s in ((True, True), (True, False), (False, True)):
        qa = qt.clone().requires_grad_(need_q)
        sa = torch.from_numpy(s).to(DEV).requires_grad_(need_s)
        head_t(qa, sa, torch.from_numpy(y).to(DEV)).sum().backward()
        torch.cuda.synchronize()
        assert (not need_q or torch.isfinite(qa.grad).all()) and (not need_s or torch.isfinite(sa.grad).all())
    print("tensor-core forward/backward: ok", flush=True)
    print("SANITIZE_DRIVER_OK", flush=True)


if __name__ == "__main__":
    main()

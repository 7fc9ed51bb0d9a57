"""Tensor-core forward + backward of NWHead.forward against a LARGE shared support.

The reference differentiates `NWHead.forward` (nwhead/nw.py:266-289) with plain autograd (train.py:414), whatever
the support size.  The direct fp32 kernels (nw_direct_backward) cover episodic training and are HBM-bound for a few
dozen queries; for a big batch against a big support the work is four dense contractions of 2*B*N*d FLOP each, and
they belong on the tensor cores (closed form: SURVEY.md B.2):

    forward   class log-sum-exp table L (B, C)                      fused forward (nw_forward_class_lse)
    backward  T(b, c) = g/(P + 1e-12) - sum_c' g P/(P + 1e-12)      (B, C) host-side table, P = softmax_c L
              W  (B, N) = dL/dscore [/ distance], bf16              nw_backward_coefficients, orientation 0
              W^t (N, B)                                            nw_backward_coefficients, orientation 1
              grad_q = W S   - rowsum(W) q                          nw_dense_products (split-K over the supports)
              grad_s = W^t Q - colsum(W) s                          nw_dense_products
    (linear scores: grad_q = W S, grad_s = W^t Q; normalised kernels chain through the normalisation.)

Accuracy: scores are recomputed from bf16 operands exactly as in the forward and the coefficients are rounded to
bf16, so gradients carry ~2^-9 relative rounding per term (tests: <= 2e-2 of the gradient's max-abs against the
float64 closed form) — the usual bf16 training trade; the direct path stays the default for small problems.
"""
import torch

from . import _abi
from ._abi import check, load, ptr, stream_of
from .bank import EUCLID_KINDS, NORMALISED_KINDS, SupportBank, logp_from_class_lse

# NWHead.forward takes this path when gradients are needed, the support is shared (2-D) and the problem is at
# least this big (below it the direct fp32 kernels are within a small factor of their HBM bound)
MIN_QUERIES = 64
MIN_PAIRS = 1 << 24


def transpose_operand(x: torch.Tensor) -> torch.Tensor:
    """k-block-major bf16 rows (kb, R, 64) -> the same matrix TRANSPOSED, k-block-major over the rows:
    (ceil(R / 64), kb * 64, 64), out[r / 64][f][r % 64] = x[f / 64][r][f % 64] (zero rows appended):
    nw_transpose_kblocks, one pass over the operand at HBM speed."""
    kb, r, _ = x.shape
    out = torch.empty(((r + 63) // 64, kb * 64, 64), dtype=torch.bfloat16, device=x.device)
    check(load().nw_transpose_kblocks(ptr(x), r, kb, ptr(out), stream_of(x.device)), "nw_transpose_kblocks")
    return out


def bank_transposed(bank: SupportBank) -> torch.Tensor:
    """S^t of a bank, kept with the bank (a fixed support is transposed once, not once per training step)."""
    t = getattr(bank, "_transposed", None)
    if t is None:
        t = bank._transposed = transpose_operand(bank.feats_bf16)
    return t


def backward_table(class_lse: torch.Tensor, grad_logp: torch.Tensor):
    """(row_lse (B,), T (B, C)) from the class log-sum-exp table and dL/dlogp (see the module docstring)."""
    row_lse = torch.logsumexp(class_lse, dim=1)
    p = torch.exp(class_lse - row_lse[:, None])
    g = grad_logp / (p + 1e-12)
    return row_lse.contiguous(), (g - (g * p).sum(1, keepdim=True)).contiguous()


def coefficients(bank: SupportBank, q_bf16, q_sq, row_lse, table, scale: float, orientation: int):
    """orientation 0: W (ceil(N/64), B, 64) and rowsum(W) (B,); orientation 1: W^t (ceil(B/64), N, 64) and
    colsum(W) (N,).  bf16, zero padded; the sums are of the rounded values, in a fixed order."""
    lib = load()
    dev = bank.device
    b, n = q_bf16.shape[1], len(bank)
    epi = _abi.EPI_EUCLID if bank.kind in EUCLID_KINDS else _abi.EPI_LINEAR
    n_rows, n_cols = (b, n) if orientation == 0 else (n, b)
    out = torch.empty(((n_cols + 63) // 64, n_rows, 64), dtype=torch.bfloat16, device=dev)
    sums = torch.empty((n_rows,), dtype=torch.float32, device=dev)
    ws = torch.empty((lib.nw_backward_coefficients_workspace_elems(n_rows, n_cols),), dtype=torch.float32, device=dev)
    if orientation == 0:
        args = (ptr(q_bf16), ptr(q_sq), b, ptr(bank.feats_bf16), ptr(bank.sqnorm), n, bank.row_elems, ptr(row_lse),
                None, None, ptr(bank.labels))
    else:
        table = table.t().contiguous()  # (C, B): the row's class selects a contiguous run over the queries
        args = (ptr(bank.feats_bf16), ptr(bank.sqnorm), n, ptr(q_bf16), ptr(q_sq), b, bank.row_elems, None,
                ptr(bank.labels), ptr(row_lse), None)
    check(lib.nw_backward_coefficients(epi, float(scale), orientation, *args, ptr(table), table.stride(0), ptr(out),
                                       ptr(sums), ptr(ws), ws.numel(), stream_of(dev)), "nw_backward_coefficients")
    return out, sums


def finish(raw: torch.Tensor, rows_bf16, sums, perm, d: int) -> torch.Tensor:
    """grad[dst(r)] = raw[r, :d] - sums[r] * stored_row(r)  (nw_backward_finish; sums None: copy / permute)."""
    n_rows = raw.shape[0]
    out = torch.empty((n_rows, d), dtype=torch.float32, device=raw.device)
    check(load().nw_backward_finish(ptr(raw), raw.stride(0), ptr(rows_bf16), ptr(sums), ptr(perm), n_rows, d, ptr(out),
                                    d, stream_of(raw.device)), "nw_backward_finish")
    return out


def choose_kslices(units: int, workers: int, kb: int) -> int:
    """Split K so that the units fill whole waves of persistent workers: cost = waves x (k-blocks per unit + the
    per-unit pipeline fill and epilogue, ~16 k-blocks' worth).  B=4096 against config 3 is 128 units on 74 CTA pairs
    (two waves, the second 73 % full) unsplit.  The result follows the library's own rounding: no empty slice."""
    kslices = min(range(1, max(1, min(64, kb // 8)) + 1),
                  key=lambda ks: -(-units * ks // workers) * (-(-kb // ks) + 16))
    per = -(-kb // kslices)
    return -(-kb // per)


def dense_products(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a (kb, n_a, 64), b (kb, n_b, 64) bf16 k-block-major -> a @ b^t (n_a, n_b) fp32 on the tensor cores.
    Skinny problems are split along K so that every SM has a unit; the partial products are summed here."""
    lib = load()
    kb, n_a, _ = a.shape
    n_b = b.shape[1]
    assert b.shape[0] == kb and a.dtype == b.dtype == torch.bfloat16
    dev = a.device
    plan = _abi.forward_plan(n_a, n_b)
    units = plan.chunks * plan.q_tiles
    workers = torch.cuda.get_device_properties(dev).multi_processor_count // (2 if plan.cta_pair else 1)
    kslices = choose_kslices(units, workers, kb)
    ld = (n_b + 3) // 4 * 4
    out = torch.empty((kslices, n_a, ld), dtype=torch.float32, device=dev)
    check(lib.nw_dense_products(ptr(a), n_a, ptr(b), n_b, kb * 64, kslices, ptr(out), ld, n_a * ld, stream_of(dev)),
          "nw_dense_products")
    res = out[0] if kslices == 1 else out.sum(0)
    return res[:, :n_b]


def support_gradient(wt: torch.Tensor, q_bf16: torch.Tensor, bank: SupportBank, colsum) -> torch.Tensor:
    """grad wrt the stored support rows, (N, d) in the row order of the tensor the bank was built from, in ONE launch
    (nw_dense_products_transposed): W^t streams once as the kernel's bank operand against Q^t, and the epilogue
    subtracts colsum(W) * stored row (colsum None: linear scores) and scatters the rows back through the bank's
    sort permutation."""
    lib = load()
    dev = bank.device
    n, d = len(bank), bank.d
    qt = transpose_operand(q_bf16)  # (ceil(B/64), row_elems, 64): features as rows, K = the queries
    assert wt.shape[0] == qt.shape[0] and wt.shape[1] == n
    dst = None
    if bank.perm is not None:
        dst = getattr(bank, "_perm_i32", None)
        if dst is None:
            dst = bank._perm_i32 = bank.perm.to(torch.int32)
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    rows_t = bank_transposed(bank) if colsum is not None else None
    check(lib.nw_dense_products_transposed(ptr(qt), qt.shape[1], ptr(wt), n, qt.shape[0] * 64, ptr(colsum), ptr(rows_t),
                                           ptr(dst), d, ptr(out), d, stream_of(dev)), "nw_dense_products_transposed")
    return out


class NWTensorFunction(torch.autograd.Function):
    """NWHead.forward(x, sx, sy) for a 2-D support on the tensor cores, forward and backward."""

    @staticmethod
    def forward(ctx, x, sx, sy, logit_scale, kind, n_classes, bank=None):
        scale = float(logit_scale.detach().exp()) if logit_scale is not None else 1.0
        if bank is None:  # (a fixed support arrives with its cached bank)
            bank = SupportBank.build(sx, sy, n_classes, kind, "bf16")  # raises like F.one_hot on a bad label
        q_bf16, q_sq = bank.prepare_queries(x)
        class_lse = bank.class_lse_prepared(q_bf16, q_sq, scale)
        ctx.bank, ctx.scale, ctx.kind = bank, scale, kind
        ctx.has_scale = logit_scale is not None
        ctx.save_for_backward(x, sx, q_bf16, q_sq, class_lse)
        return logp_from_class_lse(class_lse)

    @staticmethod
    def backward(ctx, grad_out):
        x, sx, q_bf16, q_sq, class_lse = ctx.saved_tensors
        bank, scale, kind = ctx.bank, ctx.scale, ctx.kind
        need_q, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_scale = ctx.has_scale and ctx.needs_input_grad[3]
        d = bank.d
        euclid = kind in EUCLID_KINDS
        normalised = kind in NORMALISED_KINDS
        row_lse, table = backward_table(class_lse, grad_out.float())

        def through_normalisation(g_hat, raw):
            """gradient with respect to the normalised row -> gradient with respect to the raw row"""
            if not normalised:
                return g_hat
            norm = raw.norm(dim=1, keepdim=True).clamp_min(1e-12)
            unit = raw / norm
            return (g_hat - (g_hat * unit).sum(1, keepdim=True) * unit) / norm

        # grad = W S - rowsum(W) q takes q from the STORED operand rows (centred / normalised AND rounded to bf16,
        # like S): otherwise the two terms of a close (query, support) pair do not cancel
        gq = gs = gscale = None
        if need_q or need_scale:
            w, rowsum = coefficients(bank, q_bf16, q_sq, row_lse, table, scale, 0)
            raw = dense_products(w, bank_transposed(bank))
            del w
            g_hat = finish(raw, q_bf16, rowsum if euclid else None, None, d)
            if need_scale:  # d score / d logit_scale = score, and sum_j coef * score = q_hat . grad_q_hat
                gscale = (g_hat * q_bf16.permute(1, 0, 2).reshape(x.shape[0], -1)[:, :d].float()).sum()
            if need_q:
                gq = through_normalisation(g_hat, x)
        if need_s:
            wt, colsum = coefficients(bank, q_bf16, q_sq, row_lse, table, scale, 1)
            gs = through_normalisation(support_gradient(wt, q_bf16, bank, colsum if euclid else None), sx)
            del wt
        return gq, gs, None, gscale, None, None, None


def wants_tensor_path(n_query: int, n_support: int, support_dims: int, mode: str) -> bool:
    if mode == "direct" or support_dims != 2:
        return False
    if mode == "tensor":
        return True
    return n_query >= MIN_QUERIES and n_query * n_support >= MIN_PAIRS

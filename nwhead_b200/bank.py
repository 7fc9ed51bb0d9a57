"""Device-resident support bank and the tensor-core forward built on it.

Replaces the CPU fp32 bank of the reference (nwhead/nw.py:118-125, 213-243; nwhead/support.py:113-120)
and the per-call whole-bank host->device copy in NWNet.predict (nwhead/nw.py:156).

HBM layout (all contiguous, 16-byte aligned):
    feats_bf16 (row_elems/64, N, 64) bf16   K-BLOCK-MAJOR; row_elems = precision*d rounded up to 64.  Each k-block
                                     of a row is one 128-B TMA swizzle row, each (row tile, k-block) box is contiguous
    sqnorm     (N,)          fp32    squared norms of the ROUNDED rows (euclidean kinds)
    labels     (N,)          int32   class-sorted
    offsets    (C+1,)        int32   first row of every class
    perm       (N,)          int64   bank row -> row of the tensor the bank was built from (None = identity)
    center     (d,)          fp32    column mean removed before rounding (euclidean only)
"""
from __future__ import annotations

import os
import weakref
from typing import Optional

import torch

from . import _abi
from ._abi import KIND, check, load, ptr, stream_of

NORMALISED_KINDS = ("hypersphere_euclidean", "cosine", "clip")
EUCLID_KINDS = ("euclidean", "hypersphere_euclidean")
AUTO_X3_MAX_ELEMS = 1 << 26  # 'auto' uses the 3-product split up to 64 Mi bank elements


def resolve_precision(precision: str, n: int, d: int) -> int:
    precision = os.environ.get("NW_B200_PRECISION", precision)
    if precision == "auto":
        return _abi.PREC_BF16X3 if n * d <= AUTO_X3_MAX_ELEMS else _abi.PREC_BF16
    if precision == "bf16":
        return _abi.PREC_BF16
    if precision == "bf16x3":
        return _abi.PREC_BF16X3
    raise ValueError(f"unknown precision {precision!r} (use 'auto', 'bf16' or 'bf16x3')")


def rows_to_bf16(rows: torch.Tensor, *, perm, center, normalize: bool, layout: int, precision: int):
    """nw_rows_to_bf16: fp32 (R, d) -> (bf16 k-block-major (row_elems/64, R, 64), sqnorm (R,))."""
    lib = load()
    assert rows.dim() == 2 and rows.dtype == torch.float32 and rows.stride(1) == 1
    n, d = rows.shape
    row_elems = lib.nw_row_elems(d, precision)
    out = torch.empty((row_elems // 64, n, 64), dtype=torch.bfloat16, device=rows.device)
    sq = torch.empty((n,), dtype=torch.float32, device=rows.device)
    check(
        lib.nw_rows_to_bf16(ptr(rows), n, d, rows.stride(0), ptr(perm), ptr(center), int(normalize), layout,
                            precision, ptr(out), row_elems, ptr(sq), stream_of(rows.device)),
        "nw_rows_to_bf16",
    )
    return out, sq


class SupportBank:
    """A class-sorted support set resident in HBM in the layout the fused forward consumes."""

    def __init__(self, feats_bf16, sqnorm, labels_i32, offsets, perm, center, kind, precision, d, n_classes):
        self.feats_bf16 = feats_bf16
        self.sqnorm = sqnorm
        self.labels = labels_i32
        self.offsets = offsets
        self.perm = perm
        self.center = center
        self.kind = kind
        self.precision = precision
        self.d = d
        self.n_classes = n_classes
        # row j is the one support of class j (cluster / random mode banks, nwhead/support.py:117-129 with one
        # centroid or sample per class): the class log-sum-exp IS the score matrix, see class_lse_prepared
        self.identity_classes = False

    def __len__(self):
        return self.feats_bf16.shape[1]

    @property
    def device(self):
        return self.feats_bf16.device

    @property
    def row_elems(self):
        return self.feats_bf16.shape[0] * 64

    def rows_as_matrix(self, rows=None) -> torch.Tensor:
        """(n, row_elems) row-major fp32 view of the stored bf16 rows (inspection / tests)."""
        t = self.feats_bf16 if rows is None else self.feats_bf16[:, rows]
        return t.permute(1, 0, 2).reshape(t.shape[1], -1).float()

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.feats_bf16, self.sqnorm, self.labels, self.offsets))

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def build(feats: torch.Tensor, labels: torch.Tensor, n_classes: int, kind: str = "euclidean",
              precision: str = "auto", center: Optional[torch.Tensor] = None, use_center: bool = True,
              class_range=None, labels_validated: bool = False) -> "SupportBank":
        """feats (N, d) fp32 CUDA, labels (N,) int64 CUDA (any order).  Raises like F.one_hot
        (nwhead/nw.py:276) when a label is outside [0, n_classes).
        labels_validated=True: the labels are known to be inside [0, n_classes) (e.g. rows gathered from a bank that
        was validated when it was built — knn mode builds such a support for every batch): the support is class-
        sorted unconditionally and NO host synchronisation takes place (the check and the "already sorted?" probe are
        what needs one: an invalid label must never reach the fused forward's class-indexed stores)."""
        if kind not in KIND:
            raise NotImplementedError(kind)
        dev = _abi.require_cuda(feats, labels)
        lib = load()
        if feats.dtype in (torch.float16, torch.bfloat16):  # features computed under autocast
            feats = feats.float()
        if feats.dtype != torch.float32:
            raise TypeError(f"support features must be float32, got {feats.dtype}")  # reference: fp64 raises
        if labels.dtype != torch.int64:
            raise RuntimeError("one_hot is only applicable to index tensor of type LongTensor.")
        feats = feats.detach()
        if feats.stride(1) != 1:
            feats = feats.contiguous()
        labels = labels.detach().contiguous()
        n, d = feats.shape
        st = stream_of(dev)
        prec = resolve_precision(precision, n, d)

        labels_i32 = torch.empty((n,), dtype=torch.int32, device=dev)
        status = torch.empty((2,), dtype=torch.int32, device=dev)
        if labels_validated:
            bad, descents = 0, 1
        else:
            check(lib.nw_labels_to_i32(ptr(labels), None, n, n_classes, ptr(labels_i32), ptr(status), st),
                  "nw_labels_to_i32")
            bad, descents = status.tolist()
        if bad:
            raise RuntimeError("Class values must be smaller than num_classes.")
        perm = None
        if descents:  # unsorted support: class-sort it (stable).  torch.sort is plumbing, not the hot path.
            perm = torch.sort(labels, stable=True).indices.contiguous()
            check(lib.nw_labels_to_i32(ptr(labels), ptr(perm), n, n_classes, ptr(labels_i32), ptr(status), st),
                  "nw_labels_to_i32")
        offsets = torch.empty((n_classes + 1,), dtype=torch.int32, device=dev)
        check(lib.nw_class_offsets(ptr(labels_i32), n, n_classes, ptr(offsets), st), "nw_class_offsets")

        if kind == "euclidean" and use_center and center is None:
            center = torch.empty((d,), dtype=torch.float32, device=dev)
            ws_bytes = lib.nw_column_mean_workspace_bytes(d)
            ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=dev)
            check(lib.nw_column_mean(ptr(feats), n, d, feats.stride(0), ptr(center), ptr(ws), ws_bytes, st),
                  "nw_column_mean")
        if kind != "euclidean":
            center = None
        feats_bf16, sqnorm = rows_to_bf16(feats, perm=perm, center=center, normalize=kind in NORMALISED_KINDS,
                                          layout=_abi.ROWS_BANK, precision=prec)
        bank = SupportBank(feats_bf16, sqnorm, labels_i32, offsets, perm, center, kind, prec, d, n_classes)
        if n == n_classes and not labels_validated:  # this path has synchronised already (label check above)
            bank.identity_classes = bool(torch.equal(
                offsets, torch.arange(n_classes + 1, dtype=torch.int32, device=dev)))
        return bank

    # ------------------------------------------------------------------------------------------
    def subset(self, bank_rows: torch.Tensor) -> "SupportBank":
        """Bank restricted to the given bank rows (must keep the class-sorted order): used by
        mode='random' (nwhead/support.py:126-129, 139) and by class-aligned sharding."""
        lib = load()
        idx = bank_rows.to(self.device, torch.int64)
        labels = self.labels.index_select(0, idx).contiguous()
        offsets = torch.empty((self.n_classes + 1,), dtype=torch.int32, device=self.device)
        check(lib.nw_class_offsets(ptr(labels), labels.numel(), self.n_classes, ptr(offsets), stream_of(self.device)),
              "nw_class_offsets")
        perm = idx if self.perm is None else self.perm.index_select(0, idx)
        return SupportBank(self.feats_bf16.index_select(1, idx).contiguous(), self.sqnorm.index_select(0, idx).contiguous(),
                           labels, offsets, perm, self.center, self.kind, self.precision, self.d, self.n_classes)

    def class_shard(self, rank: int, world: int) -> "SupportBank":
        """Contiguous class-aligned shard: rank r owns classes [r*C/R, (r+1)*C/R)."""
        c_lo = (rank * self.n_classes) // world
        c_hi = ((rank + 1) * self.n_classes) // world
        off = self.offsets[[c_lo, c_hi]].tolist()
        r0, r1 = off
        if r1 <= r0:
            raise ValueError(f"rank {rank} owns no support rows (classes [{c_lo},{c_hi}))")
        sl = slice(r0, r1)
        lib = load()
        labels = self.labels[sl].contiguous()
        offsets = torch.empty((self.n_classes + 1,), dtype=torch.int32, device=self.device)
        check(lib.nw_class_offsets(ptr(labels), labels.numel(), self.n_classes, ptr(offsets), stream_of(self.device)),
              "nw_class_offsets")
        perm = (torch.arange(r0, r1, device=self.device) if self.perm is None else self.perm[sl]).contiguous()
        return SupportBank(self.feats_bf16[:, sl].contiguous(), self.sqnorm[sl].contiguous(), labels, offsets, perm,
                           self.center, self.kind, self.precision, self.d, self.n_classes)

    # ------------------------------------------------------------------------------------------
    _FIELDS = ("feats_bf16", "sqnorm", "labels", "offsets", "perm", "center")

    def state_dict(self) -> dict:
        """Everything needed to serve from this bank again without the fp32 features or the featurizer
        (the reference rebuilds its bank on every precompute(), nwhead/nw.py:118-125)."""
        d = {k: getattr(self, k) for k in self._FIELDS}
        d.update(kind=self.kind, precision=self.precision, d=self.d, n_classes=self.n_classes, layout_version=1)
        return d

    @staticmethod
    def from_state_dict(state: dict, device=None) -> "SupportBank":
        if state.get("layout_version") != 1:
            raise ValueError("unknown SupportBank layout version")
        t = {k: (None if state[k] is None else state[k].to(device or state[k].device).contiguous())
             for k in SupportBank._FIELDS}
        return SupportBank(t["feats_bf16"], t["sqnorm"], t["labels"], t["offsets"], t["perm"], t["center"],
                           state["kind"], state["precision"], state["d"], state["n_classes"])

    def save(self, path: str) -> None:
        torch.save({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in self.state_dict().items()}, path)

    @staticmethod
    def load(path: str, device="cuda:0") -> "SupportBank":
        return SupportBank.from_state_dict(torch.load(path, map_location="cpu"), device)

    # ------------------------------------------------------------------------------------------
    def prepare_queries(self, q: torch.Tensor):
        """fp32 queries -> (bf16 rows in the query layout, squared norms), centred / normalised like the bank."""
        if q.dtype in (torch.float16, torch.bfloat16):
            q = q.float()
        if q.dtype != torch.float32:
            raise TypeError(f"query features must be float32, got {q.dtype}")
        q = q.detach()
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"queries must be (B, {self.d}), got {tuple(q.shape)}")
        if q.stride(1) != 1:
            q = q.contiguous()
        return rows_to_bf16(q, perm=None, center=self.center, normalize=self.kind in NORMALISED_KINDS,
                            layout=_abi.ROWS_QUERY, precision=self.precision)

    def class_lse(self, q: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """(B, C) per-class log-sum-exp of the scores against this bank (−inf for absent classes)."""
        _abi.require_cuda(q, self.feats_bf16)
        if q.shape[0] == 0:  # empty batch: nothing to launch
            return torch.empty((0, self.n_classes), dtype=torch.float32, device=self.device)
        b = q.shape[0]
        if b < self.MIN_QUERY_ROWS:
            # the TMA loads of a query tile with fewer than 8 real rows (one 128-byte-swizzle atom) are slower:
            # the fused forward took 1.06 ms at B=1 against 0.74 ms at B=8 on the config-3 bank.  Zero rows are free.
            q = torch.cat((q.detach().float().reshape(b, -1), q.new_zeros((self.MIN_QUERY_ROWS - b, self.d),
                                                                           dtype=torch.float32)))
        q_bf16, q_sq = self.prepare_queries(q)
        return self.class_lse_prepared(q_bf16, q_sq, scale)[:b]

    MIN_QUERY_ROWS = 8

    def class_lse_prepared(self, q_bf16: torch.Tensor, q_sq: torch.Tensor, scale: float = 1.0, tables=None,
                           rows_per_table: int = 0):
        """tables=None: returns a fresh (B, C) table.  tables=ctypes array of device pointers (one table per
        rank, in rank order): results are stored into all of them, or with rows_per_table > 0 only into the
        table that owns the row (nw_forward_class_lse_peers); returns None."""
        lib = load()
        b = q_bf16.shape[1]
        n = len(self)
        plan = _abi.forward_plan(b, n)
        dev = self.device
        # chunk-boundary partials + room for the further epilogue sets' tables (short GEMMs: 2 sets up to d = 1024,
        # 4 up to d = 512)
        extra = (lib.nw_forward_epilogue_sets(self.row_elems) - 1) * b * self.n_classes if tables is None else 0
        side = torch.empty((max(int(plan.side_elems) + extra, 1),), dtype=torch.float32, device=dev)
        epi = _abi.EPI_EUCLID if self.kind in EUCLID_KINDS else _abi.EPI_LINEAR
        if tables is not None:
            check(
                lib.nw_forward_class_lse_peers(epi, float(scale), ptr(q_bf16), ptr(q_sq), b, ptr(self.feats_bf16),
                                               ptr(self.sqnorm), ptr(self.labels), n, self.row_elems,
                                               self.n_classes, tables, len(tables), int(rows_per_table), ptr(side),
                                               side.numel(), stream_of(dev)),
                "nw_forward_class_lse_peers",
            )
            return None
        out = torch.empty((b, self.n_classes), dtype=torch.float32, device=dev)
        if self.identity_classes:
            # one support per class, in class order: L[b, c] = score(b, c).  The dense-score epilogue writes the
            # table with coalesced 16-byte stores (config-4 predict: 57 us in the class-indexed epilogue, which
            # stores one value per class end and thread)
            check(
                lib.nw_forward_emit(epi, float(scale), ptr(q_bf16), ptr(q_sq), b, ptr(self.feats_bf16),
                                    ptr(self.sqnorm), None, n, self.row_elems, _abi.EMIT_SCORES, None, None, None,
                                    ptr(out), self.n_classes, stream_of(dev)),
                "nw_forward_emit",
            )
            return out
        check(
            lib.nw_forward_class_lse(epi, float(scale), ptr(q_bf16), ptr(q_sq), b, ptr(self.feats_bf16),
                                     ptr(self.sqnorm), ptr(self.labels), n, self.row_elems, self.n_classes,
                                     ptr(out), ptr(side), side.numel(), stream_of(dev)),
            "nw_forward_class_lse",
        )
        return out

    def _emit(self, q, scale, kind, row_lse=None, p_query=None, qlabel=None, source_order=False):
        lib = load()
        _abi.require_cuda(q, self.feats_bf16)
        q_bf16, q_sq = self.prepare_queries(q)
        b, n, dev = q.shape[0], len(self), self.device
        ld = (n + 3) // 4 * 4  # 16-byte aligned rows
        out = torch.empty((b, ld), dtype=torch.float32, device=dev)
        epi = _abi.EPI_EUCLID if self.kind in EUCLID_KINDS else _abi.EPI_LINEAR
        check(
            lib.nw_forward_emit(epi, float(scale), ptr(q_bf16), ptr(q_sq), b, ptr(self.feats_bf16), ptr(self.sqnorm),
                                ptr(self.labels), n, self.row_elems, kind, ptr(row_lse), ptr(p_query), ptr(qlabel),
                                ptr(out), ld, stream_of(dev)),
            "nw_forward_emit",
        )
        out = out[:, :n]
        if source_order and self.perm is not None:  # column j of the result <-> row j of the tensor the bank was built from
            res = torch.empty((b, n), dtype=torch.float32, device=dev)
            res[:, self.perm] = out
            return res
        return out

    def scores(self, q: torch.Tensor, scale: float = 1.0, source_order: bool = True) -> torch.Tensor:
        """Dense (B, N) similarity matrix on the tensor cores (bf16-operand accuracy; use precision='bf16x3'
        banks for near-fp32 scores).  Columns follow the bank's class-sorted rows, or with source_order the
        rows of the tensor the bank was built from."""
        return self._emit(q, scale, _abi.EMIT_SCORES, source_order=source_order)

    def topk(self, q: torch.Tensor, k: int, scale: float = 1.0, source_order: bool = True,
             query_chunk: int = 512) -> torch.Tensor:
        """Indices (B, k) of the k best-scoring support rows per query, best first: tensor-core dense scores
        (nw_forward_emit) + the bitonic ranking (nw_rank_rows), chunked over queries to bound memory.  The
        ranking has the accuracy of the bank's operands: near-fp32 for 'bf16x3' banks, ~1e-2 in score for
        'bf16' (NWNet.get_neighbors keeps the exact fp32 path by default)."""
        from .utils import rank_rows

        out = []
        for i in range(0, q.shape[0], query_chunk):
            sc = self._emit(q[i:i + query_chunk], scale, _abi.EMIT_SCORES)
            idx = rank_rows(sc.contiguous(), k)
            if source_order and self.perm is not None:
                idx = self.perm[idx]
            out.append(idx)
        return torch.cat(out, dim=0)

    def block_best(self, q: torch.Tensor, scale: float = 1.0):
        """(B, ceil(N/64)) best score of every query inside every block of 64 consecutive bank rows
        (nw_forward_emit / NW_EMIT_BLOCK_BEST): the candidate search of topk_exact.  Also returns the (B,)
        squared norms of the prepared (centred, rounded) queries."""
        lib = load()
        _abi.require_cuda(q, self.feats_bf16)
        q_bf16, q_sq = self.prepare_queries(q)
        b, n, dev = q.shape[0], len(self), self.device
        nblk = (n + 255) // 256 * 4
        ld = (b + 3) // 4 * 4
        out = torch.empty((nblk, ld), dtype=torch.float32, device=dev)
        epi = _abi.EPI_EUCLID if self.kind in EUCLID_KINDS else _abi.EPI_LINEAR
        check(
            lib.nw_forward_emit(epi, float(scale), ptr(q_bf16), ptr(q_sq), b, ptr(self.feats_bf16), ptr(self.sqnorm),
                                None, n, self.row_elems, _abi.EMIT_BLOCK_BEST, None, None, None, ptr(out), ld,
                                stream_of(dev)),
            "nw_forward_emit",
        )
        return out[:(n + 63) // 64, :b].t().contiguous(), q_sq

    def rounding_residual(self, rows: torch.Tensor) -> torch.Tensor:
        """(n,) norm of what this bank's operand rounding discards from each fp32 row (nw_rounding_residual)."""
        lib = load()
        rows = rows.detach().float().reshape(rows.shape[0], -1)
        if rows.stride(1) != 1:
            rows = rows.contiguous()
        out = torch.empty((rows.shape[0],), dtype=torch.float32, device=rows.device)
        check(lib.nw_rounding_residual(ptr(rows), rows.shape[0], self.d, rows.stride(0), ptr(self.center),
                                       self.precision, ptr(out), stream_of(rows.device)), "nw_rounding_residual")
        return out.sqrt()

    def topk_exact(self, q: torch.Tensor, k: int, source_feats: torch.Tensor, max_blocks: int = 64,
                   query_chunk: int = 2048) -> torch.Tensor:
        """EXACT k nearest supports (euclidean banks), indices into `source_feats` (the fp32 tensor the bank was
        built from), nearest first — the same ranking as the dense fp32 path (ties by ascending index), without
        the (B, N) score matrix:

          1. tensor-core pass (nw_forward_emit / NW_EMIT_BLOCK_BEST): best reduced-precision score beta_j =
             -distance of every query in every block j of 64 bank rows; the blocks of each query are ranked
             (nw_rank_rows);
          2. nw_topk_refine, ONE launch: the rows of the m best blocks of every query are gathered from
             `source_feats` and scored with the exact fp32 differences of the dense path (bit for bit), ranked by
             (score, source index) -> tau_c, the exact k-th best candidate score;
          3. certificate, in the same kernel: no row outside the candidates can score >= tau_c.  Such a row has a
             reduced-precision score <= beta_(m+1), and the pass is off by a bounded amount: the distance between
             two rounded rows differs from the true one by at most eta = |q - q~| + max_j |s_j - s~_j| (triangle
             inequality; the residual norms are measured, nw_rounding_residual), and the fp32 accumulation moves the
             squared distance by at most e2 = 2^-18 (|q|^2 + max|s|^2).  So its true score is at most
             U = -(sqrt(beta_(m+1)^2 - e2) - eta); U < tau_c certifies the query.

        Every query sizes its own candidate budget m <= `max_blocks` (at most 64) inside the kernel from the same
        bounds (blocks that could still reach the k-th best block score); queries it cannot certify take the dense
        exact path — so the result is exact for every input, and the only host synchronisation is one read of the
        count of uncertified queries per chunk (`last_topk_path` counts both routes)."""
        from .kernel import dense_scores
        from .utils import rank_rows

        if self.kind != "euclidean":
            raise NotImplementedError("topk_exact is provided for the euclidean kernel")
        lib = load()
        dev, n, d = self.device, len(self), self.d
        if source_feats.shape[0] != n or source_feats.shape[1:].numel() != d:
            raise ValueError("source_feats must be the (N, d) tensor the bank was built from")
        k = min(int(k), n)
        src_all = source_feats.detach().float().reshape(n, d)
        if not src_all.is_contiguous():
            src_all = src_all.contiguous()
        cached = getattr(self, "_resid_of", None)   # (weak reference to the tensor object, its version, max residual)
        if source_feats.is_inference():             # no version counter to validate a cache entry against
            cached = (None, None, self.rounding_residual(src_all).max().reshape(1))
        elif cached is None or cached[0]() is not source_feats or cached[1] != source_feats._version:
            cached = (weakref.ref(source_feats), source_feats._version,
                      self.rounding_residual(src_all).max().reshape(1))
            self._resid_of = cached
        resid_max = cached[2]
        if getattr(self, "_smax_sq", None) is None:
            self._smax_sq = self.sqnorm.max().reshape(1)
        nblk = (n + 63) // 64
        m_cap = max(1, min(int(max_blocks), 64, nblk))
        out = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
        self.last_topk_path = {"blocks": 0, "dense": 0}
        for i0 in range(0, q.shape[0], query_chunk):
            qc = q[i0:i0 + query_chunk].detach().float().reshape(-1, d).contiguous()
            b = qc.shape[0]
            done = torch.zeros((b,), dtype=torch.int32, device=dev)
            n_left = b
            if m_cap * 64 >= k:
                best, q_sq = self.block_best(qc)                     # (b, nblk) scores = -distance
                width = min(nblk, max(m_cap + 1, k))
                order = rank_rows(best, width)                       # blocks by best score, descending
                sorted_best = best.gather(1, order)
                resid_q = self.rounding_residual(qc)
                pending = torch.zeros((1,), dtype=torch.int32, device=dev)
                check(lib.nw_topk_refine(ptr(qc), b, d, ptr(src_all), n, ptr(self.perm), ptr(order), ptr(sorted_best),
                                         width, m_cap, nblk, k, ptr(q_sq), ptr(resid_q), ptr(self._smax_sq),
                                         ptr(resid_max), self.precision, ptr(done), ptr(out[i0:i0 + b]), ptr(pending),
                                         stream_of(dev)), "nw_topk_refine")
                n_left = int(pending.item())                         # the one host synchronisation of the chunk
            self.last_topk_path["blocks"] += b - n_left
            if n_left:  # too many near-ties for the candidate budget: dense exact scores for those queries
                rows = (done == 0).nonzero().flatten()
                out[i0 + rows] = rank_rows(dense_scores("euclidean", qc[rows], src_all), k)
                self.last_topk_path["dense"] += n_left
        return out

    def support_influence(self, q: torch.Tensor, qlabel: torch.Tensor, scale: float = 1.0,
                          source_order: bool = True) -> torch.Tensor:
        """support_influence (reference util/metric.py:23-50) computed from FEATURES in two tensor-core passes:
        pass 1 = the fused forward (class log-sum-exp -> softmax mass of each query's class and the softmax
        normaliser), pass 2 = scores again with the influence formula in the epilogue.  The (B, N) weight matrix
        the reference takes as an input is never materialised."""
        lib = load()
        lse = self.class_lse(q, scale)
        b = lse.shape[0]
        qy = qlabel.to(self.device, torch.int32).contiguous()
        z = torch.empty((b,), dtype=torch.float32, device=self.device)
        p = torch.empty((b,), dtype=torch.float32, device=self.device)
        check(lib.nw_row_stats(ptr(lse), b, self.n_classes, ptr(qy), ptr(z), ptr(p), stream_of(self.device)),
              "nw_row_stats")
        return self._emit(q, scale, _abi.EMIT_INFLUENCE, row_lse=z, p_query=p, qlabel=qy, source_order=source_order)

    def graphed(self, batch: int, scale: float = 1.0) -> "GraphedForward":
        """CUDA-graph replay of forward() for a fixed batch size (see GraphedForward)."""
        return GraphedForward(self, batch, scale)

    def forward(self, q: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """log(softmax-weighted label aggregation + 1e-12): NWHead.forward (nwhead/nw.py:266-289)."""
        return logp_from_class_lse(self.class_lse(q, scale))

    # Small banks (cluster centroids, CUB-sized full banks) make predict launch-bound: five kernel launches and four
    # ctypes calls for a few microseconds of GPU work.  forward_auto replays the step as ONE CUDA graph once the same
    # (batch, scale) has been seen twice on this bank; the result is copied out of the graph's static buffer, so it
    # is an ordinary tensor.  NW_B200_GRAPHS=0 switches it off.
    GRAPH_MAX_BANK_ELEMS = 1 << 24
    GRAPH_MAX_BATCH = 8192
    GRAPH_CACHE = 4

    def forward_auto(self, q: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        b = q.shape[0]
        if (len(self) * self.row_elems > self.GRAPH_MAX_BANK_ELEMS or not 0 < b <= self.GRAPH_MAX_BATCH
                or q.dtype != torch.float32 or q.dim() != 2 or q.shape[1] != self.d
                or self.device.index != torch.cuda.current_device()
                or torch.cuda.is_current_stream_capturing() or os.environ.get("NW_B200_GRAPHS", "1") == "0"):
            return self.forward(q, scale)
        cache = self.__dict__.setdefault("_graph_cache", {})
        key = (b, float(scale))
        entry = cache.get(key)
        if entry is None:
            if len(cache) >= self.GRAPH_CACHE:
                cache.pop(next(iter(cache)))
            cache[key] = 1                       # seen once: capture on the next call with this shape
            return self.forward(q, scale)
        if entry == 1:
            entry = cache[key] = GraphedForward(self, b, scale)
        return entry(q).clone()


class GraphedForward:
    """SupportBank.forward for a fixed batch size captured in a CUDA graph: query prep, -inf fill, fused forward,
    chunk merge and finalise replay as ONE graph launch (small-batch predict is launch-latency bound: five
    kernel launches + four ctypes calls per step otherwise).  The result tensor is reused by the next call."""

    def __init__(self, bank: "SupportBank", batch: int, scale: float = 1.0):
        self.bank, self.batch = bank, batch
        dev = bank.device
        self.q = torch.zeros((batch, bank.d), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside capture (kernel attributes, allocator pools)
                bank.forward(self.q, scale)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = bank.forward(self.q, scale)

    def __call__(self, q: torch.Tensor) -> torch.Tensor:
        if q.shape != self.q.shape:
            raise ValueError(f"graphed forward was captured for {tuple(self.q.shape)}, got {tuple(q.shape)}")
        self.q.copy_(q, non_blocking=True)
        self.graph.replay()
        return self.out


def logp_from_class_lse(class_lse: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = load()
    b, c = class_lse.shape
    if b == 0:
        return class_lse.clone() if out is None else out
    if not class_lse.is_contiguous():
        class_lse = class_lse.contiguous()
    if out is None:
        out = torch.empty_like(class_lse)
    check(lib.nw_logp_from_class_lse(ptr(class_lse), b, c, ptr(out), stream_of(class_lse.device)),
          "nw_logp_from_class_lse")
    return out


def class_lse_merge_(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a <- log(exp(a) + exp(b)) elementwise (generic row-sharded merge)."""
    check(load().nw_class_lse_merge(ptr(a), ptr(b), a.numel(), stream_of(a.device)), "nw_class_lse_merge")
    return a

"""Full-mode predict with the support bank sharded across the GPUs of one box (new; the reference has no
multi-GPU code — SURVEY.md 2.3, 8e).

Partitioning: class-aligned contiguous ranges of the class-sorted bank; rank r owns classes
[r*C/R, (r+1)*C/R).  Queries are replicated.  Each rank runs the fused forward on its shard and emits the
(B, C) class log-sum-exp table with -inf for classes it does not own; because every class is owned by
exactly one rank, ONE all-reduce(MAX) over NVLink merges the partial softmax state exactly.  Every rank
then finalises log(P + 1e-12) locally.

`merge_class_lse` is backend-agnostic (NCCL on GPUs, gloo in the CPU tests of the host logic).
"""
import ctypes

import torch
import torch.distributed as dist


def class_range(rank: int, world: int, n_classes: int):
    """Classes owned by `rank`: [lo, hi)."""
    return (rank * n_classes) // world, ((rank + 1) * n_classes) // world


def merge_class_lse(partial: torch.Tensor, group=None, class_aligned: bool = True) -> torch.Tensor:
    """In-place exact merge of per-rank class-LSE tables.

    class_aligned=True : every (b, c) entry is finite on at most one rank -> all-reduce(MAX) is exact.
    class_aligned=False: generic row sharding -> log-sum-exp merge as max + log(sum exp(x - max)):
                         one MAX and one SUM all-reduce."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    if class_aligned:
        dist.all_reduce(partial, op=dist.ReduceOp.MAX, group=group)
        return partial
    mx = partial.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    safe = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)
    e = torch.exp(partial - safe)
    dist.all_reduce(e, op=dist.ReduceOp.SUM, group=group)
    partial.copy_(safe + torch.log(e))
    return partial


class PeerTables:
    """Double-buffered (max_batch, C) class-LSE tables in symmetric memory (one per rank, peer-mapped over
    NVLink).  The fused forward stores every class-LSE entry its shard owns into ALL ranks' tables from the
    epilogue (nw_forward_class_lse_peers), so the exchange overlaps the MMAs and no all-reduce is launched; a
    signal-pad barrier on the stream separates the writes from the local finalise.

    Hazards: step i writes table i%2 everywhere, then barrier_i, then every rank reads only its own table i%2.
    Table i%2 is written again at step i+2, which a rank can only start after barrier_{i+1} — and every rank
    enqueues barrier_{i+1} after its finalise of step i.  Entries of classes no shard owns keep the -inf the
    tables are created with; every owned entry is overwritten on each step."""

    def __init__(self, max_batch: int, n_classes: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.max_batch, self.n_classes = max_batch, n_classes
        self.tables, self.handles, self.ptr_arrays = [], [], []
        for _ in range(2):
            t = symm_mem.empty((max_batch, n_classes), dtype=torch.float32, device=device)
            t.fill_(float("-inf"))
            hdl = symm_mem.rendezvous(t, self.group.group_name)
            ptrs = list(hdl.buffer_ptrs)  # one table per rank, in rank order
            self.tables.append(t)
            self.handles.append(hdl)
            self.ptr_arrays.append((ctypes.c_void_p * len(ptrs))(*ptrs))
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every table is -inf before any peer may store into it
        self.step = 0

    def next(self):
        i = self.step % 2
        self.step += 1
        return self.tables[i], self.handles[i], self.ptr_arrays[i], i


class PeerQueries:
    """Double-buffered query buffers in symmetric memory: (row_elems/64, max_batch, 64) bf16 + (max_batch,) fp32
    squared norms per rank.  `replicate(q_rows)` converts THIS rank's rows of the batch and stores them into every
    rank's buffer (nw_rows_to_bf16_peers: NVLink stores from the conversion kernel), then a signal barrier on the
    current stream; afterwards every rank holds the whole prepared batch.  Replaces an fp32 all-gather of the
    queries (twice the NVLink bytes) followed by world-fold redundant conversion.

    Hazards: step i writes buffer i%2 on every rank, barrier, then each rank's forward reads its own copy.  Buffer
    i%2 is written again at step i+2; a rank only issues that after it has RETIRED step i (FullModePredictor.result),
    which needs the exchange after the forward of step i — and every rank takes part in that exchange only once its
    own forward of step i has finished reading the buffer."""

    CHANNEL0 = 2  # signal-pad channels 0/1 belong to PeerTables

    def __init__(self, shard, max_batch: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _abi

        self.group = group if group is not None else dist.group.WORLD
        self.shard, self.max_batch = shard, max_batch
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        kb = shard.row_elems // 64
        self.bufs = []
        for _ in range(2):
            qb = symm_mem.empty((kb, max_batch, 64), dtype=torch.bfloat16, device=shard.device)
            qs = symm_mem.empty((max_batch,), dtype=torch.float32, device=shard.device)
            qb.zero_()
            qs.zero_()
            hb = symm_mem.rendezvous(qb, self.group.group_name)
            hs = symm_mem.rendezvous(qs, self.group.group_name)
            pb = (ctypes.c_void_p * self.world)(*list(hb.buffer_ptrs))
            ps = (ctypes.c_void_p * self.world)(*list(hs.buffer_ptrs))
            self.bufs.append((qb, qs, hb, pb, ps))
        torch.cuda.synchronize(shard.device)
        dist.barrier(self.group)
        self.step = 0
        self._abi = _abi

    def replicate(self, q_rows: torch.Tensor, batch: int):
        """q_rows: this rank's (batch / world, d) fp32 rows.  Returns the (row_elems/64, batch, 64) bf16 batch and
        its (batch,) squared norms, complete on every rank once the stream reaches the barrier."""
        abi, bank = self._abi, self.shard
        rows = q_rows.shape[0]
        if rows * self.world != batch or batch != self.max_batch:
            raise ValueError(f"expected {self.max_batch // self.world} rows per rank (batch {self.max_batch}), got {rows}")
        if q_rows.dtype != torch.float32 or q_rows.stride(1) != 1:
            q_rows = q_rows.float().contiguous()
        i = self.step % 2
        self.step += 1
        qb, qs, hb, pb, ps = self.bufs[i]
        from .bank import NORMALISED_KINDS

        abi.check(abi.load().nw_rows_to_bf16_peers(
            abi.ptr(q_rows), rows, bank.d, q_rows.stride(0), abi.ptr(bank.center), int(bank.kind in NORMALISED_KINDS),
            bank.precision, pb, ps, self.world, batch, self.rank * rows, bank.row_elems, abi.stream_of(bank.device)),
            "nw_rows_to_bf16_peers")
        hb.barrier(channel=self.CHANNEL0 + i)
        return qb, qs


class ShardedBank:
    """This rank's class-aligned shard of a support bank + the merged forward.

    exchange='nccl'  : local table + ONE all-reduce(MAX) (works everywhere).
    exchange='peer'  : in-kernel NVLink peer stores into every rank's table + a signal barrier (PeerTables)."""

    def __init__(self, shard, group=None, exchange="nccl", max_batch=0):
        if exchange not in ("nccl", "peer"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.peer = None
        if self.world > 1:
            # Validate the shard plan COLLECTIVELY: a rank without support rows (C < world, empty class ranges)
            # must make every rank raise, not just itself — the others would wait in the next collective forever.
            rows = torch.tensor([0 if shard is None else len(shard)], dtype=torch.int64,
                                device=None if shard is None else shard.device)
            dist.all_reduce(rows, op=dist.ReduceOp.MIN, group=group)
            if int(rows.item()) == 0:
                raise ValueError("a rank owns no support rows: use fewer ranks than non-empty classes")
        if exchange == "peer" and self.world > 1:
            if max_batch <= 0:
                raise ValueError("exchange='peer' needs max_batch > 0 (rows of the symmetric class-LSE tables)")
            self.peer = PeerTables(max_batch, shard.n_classes, shard.device, group)

    @staticmethod
    def from_full(bank, group=None, exchange="nccl", max_batch=0):
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return ShardedBank(bank, group, exchange, max_batch)
        try:
            shard = bank.class_shard(rank, world)
        except ValueError:
            shard = None  # this rank would own nothing: ShardedBank raises on EVERY rank
        if shard is None:  # take part in the collective validation, which raises
            rows = torch.zeros((1,), dtype=torch.int64, device=bank.device)
            dist.all_reduce(rows, op=dist.ReduceOp.MIN, group=group)
            raise ValueError("a rank owns no support rows: use fewer ranks than non-empty classes")
        return ShardedBank(shard, group, exchange, max_batch)

    def class_lse(self, q, scale: float = 1.0):
        """(B, C) class log-sum-exp over the WHOLE bank, identical on every rank.  With the peer exchange the
        result is a view of a double-buffered symmetric table: consume it before the call after next."""
        if self.peer is None:
            return merge_class_lse(self.shard.class_lse(q, scale), self.group)
        b = q.shape[0]
        if b > self.peer.max_batch:
            raise ValueError(f"batch {b} exceeds the peer tables' max_batch {self.peer.max_batch}")
        table, hdl, ptrs, ch = self.peer.next()
        q_bf16, q_sq = self.shard.prepare_queries(q)
        self.shard.class_lse_prepared(q_bf16, q_sq, scale, tables=ptrs)
        hdl.barrier(channel=ch)  # all ranks' peer stores have landed (kernel completion + signal exchange)
        return table[:b]

    def class_lse_rows(self, q, scale: float = 1.0):
        """This rank's rows [rank*B/R, (rank+1)*B/R) of the merged table, shape (B/R, C).  With the peer exchange
        each class-LSE entry crosses NVLink once (all-to-all) instead of being replicated to every rank."""
        b = q.shape[0]
        rank = dist.get_rank(self.group) if self.world > 1 else 0
        if b % self.world:
            raise ValueError(f"batch {b} is not a multiple of the world size {self.world}")
        rows = b // self.world
        if self.peer is None:
            return self.class_lse(q, scale)[rank * rows:(rank + 1) * rows]
        if b > self.peer.max_batch:
            raise ValueError(f"batch {b} exceeds the peer tables' max_batch {self.peer.max_batch}")
        table, hdl, ptrs, ch = self.peer.next()
        q_bf16, q_sq = self.shard.prepare_queries(q)
        self.shard.class_lse_prepared(q_bf16, q_sq, scale, tables=ptrs, rows_per_table=rows)
        hdl.barrier(channel=ch)
        return table[rank * rows:(rank + 1) * rows]

    def class_lse_rows_prepared(self, q_bf16, q_sq, scale: float = 1.0):
        """class_lse_rows for a batch that is already converted and replicated (PeerQueries.replicate)."""
        b = q_bf16.shape[1]
        rank = dist.get_rank(self.group) if self.world > 1 else 0
        rows = b // self.world
        if self.peer is None:
            lse = merge_class_lse(self.shard.class_lse_prepared(q_bf16, q_sq, scale), self.group)
            return lse[rank * rows:(rank + 1) * rows]
        table, hdl, ptrs, ch = self.peer.next()
        self.shard.class_lse_prepared(q_bf16, q_sq, scale, tables=ptrs, rows_per_table=rows)
        hdl.barrier(channel=ch)
        return table[rank * rows:(rank + 1) * rows]

    def forward(self, q, scale: float = 1.0):
        from .bank import logp_from_class_lse

        return logp_from_class_lse(self.class_lse(q, scale))

    def forward_rows(self, q, scale: float = 1.0):
        """log-probs of this rank's rows only (each rank finalises B/R rows)."""
        from .bank import logp_from_class_lse

        return logp_from_class_lse(self.class_lse_rows(q, scale))

"""Pipelined full-mode predict from host memory (the end-to-end serving path).

`FullModePredictor` takes query features that live in (pinned) HOST memory and returns log-probabilities
in host memory.  Host<->device copies run on their own CUDA streams and are double-buffered, so the copy of
batch i+1 and the read-back of batch i-1 overlap the fused forward of batch i.

With a process group (one rank per GPU, class-aligned bank shards — nwhead_b200/dist.py) the host I/O is
sharded too: rank r uploads only rows [r*B/R, (r+1)*B/R) of the batch, converts them to the bf16 query layout and
stores them into every rank's query buffer over NVLink from the conversion kernel itself (peer exchange; with the
NCCL exchange: an all-gather of the fp32 rows), the ranks run the fused forward on their bank shard, exchange the
class log-sum-exp entries (in-kernel peer stores, or ONE all-reduce(MAX)), and each rank finalises and returns only
its own rows.  PCIe traffic per rank drops by R while every query still sees the whole bank.
"""
import torch
import torch.distributed as dist

from .bank import logp_from_class_lse


class _Slot:
    def __init__(self, rows, rows_total, d, c, device, sharded):
        self.q_slice = torch.empty((rows, d), dtype=torch.float32, device=device)
        self.q_full = torch.empty((rows_total, d), dtype=torch.float32, device=device) if sharded else self.q_slice
        self.logp = torch.empty((rows, c), dtype=torch.float32, device=device)
        self.out_host = torch.empty((rows, c), dtype=torch.float32).pin_memory()
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.busy = False
        self.ticket = -1


class FullModePredictor:
    """predict(mode='full') for feature batches in host memory.

    bank : SupportBank (whole bank, world size 1), this rank's class-aligned shard, or a dist.ShardedBank
           (which also selects the exchange: NCCL all-reduce or in-kernel NVLink peer stores).
    rows : rows of the query batch THIS rank uploads / returns per call (B, or B / world_size).
    """

    def __init__(self, bank, rows: int, group=None, depth: int = 2, scale: float = 1.0):
        from .dist import ShardedBank

        self.sharded = bank if isinstance(bank, ShardedBank) else ShardedBank(bank, group)
        bank = self.sharded.shard
        self.bank, self.group, self.scale = bank, group, scale
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rows = rows
        dev = bank.device
        self.copy_stream = torch.cuda.Stream(dev)
        self.out_stream = torch.cuda.Stream(dev)
        # Bank sharded over several GPUs with the peer exchange: the query conversion kernel replicates this rank's
        # rows into every rank's query buffer over NVLink (dist.PeerQueries); otherwise the fp32 rows are
        # all-gathered with NCCL and every rank converts the whole batch.
        self.peer_queries = None
        if self.world > 1 and self.sharded.peer is not None and depth <= 2:
            from .dist import PeerQueries

            self.peer_queries = PeerQueries(bank, rows * self.world, group)
        self.slots = [_Slot(rows, rows * self.world, bank.d, bank.n_classes, dev,
                            self.world > 1 and self.peer_queries is None) for _ in range(depth)]
        self.next = 0

    def submit(self, q_host: torch.Tensor) -> int:
        """Enqueue one batch (this rank's rows, pinned fp32 host tensor).  Returns a ticket for result()."""
        if q_host.is_cuda or q_host.shape != (self.rows, self.bank.d) or q_host.dtype != torch.float32:
            raise ValueError(f"expected a host float32 tensor of shape ({self.rows}, {self.bank.d})")
        ticket = self.next
        slot = self.slots[ticket % len(self.slots)]
        if slot.busy:
            raise RuntimeError("pipeline full: call result() on the oldest ticket first")
        slot.busy = True
        slot.ticket = ticket
        self.next += 1
        compute = torch.cuda.current_stream(self.bank.device)
        # The slot's previous use was retired by result() (d2h_done implies its compute finished), so the upload
        # may start right away and overlap the forward of the batch submitted before this one.
        prepared = None
        with torch.cuda.stream(self.copy_stream):
            slot.q_slice.copy_(q_host, non_blocking=True)
            # replicate the queries over NVLink on the copy stream too (overlaps the forward of the previous batch)
            if self.peer_queries is not None:
                prepared = self.peer_queries.replicate(slot.q_slice, self.rows * self.world)
            elif self.world > 1:
                dist.all_gather_into_tensor(slot.q_full, slot.q_slice, group=self.group)
            slot.h2d_done.record(self.copy_stream)
        compute.wait_event(slot.h2d_done)
        if prepared is not None:
            mine = self.sharded.class_lse_rows_prepared(*prepared, self.scale)
        else:
            mine = self.sharded.class_lse_rows(slot.q_full, self.scale)  # this rank's rows of the merged table
        logp_from_class_lse(mine, out=slot.logp)
        slot.compute_done.record(compute)
        self.out_stream.wait_event(slot.compute_done)
        with torch.cuda.stream(self.out_stream):
            slot.out_host.copy_(slot.logp, non_blocking=True)
            slot.d2h_done.record(self.out_stream)
        return ticket

    def result(self, ticket: int) -> torch.Tensor:
        """Blocks until the batch of `ticket` is back in host memory; returns the pinned (rows, C) tensor
        (valid until the slot is reused `depth` submits later)."""
        slot = self.slots[ticket % len(self.slots)]
        if not slot.busy or slot.ticket != ticket:
            raise ValueError(f"ticket {ticket} is not outstanding (never issued, or its result was already taken)")
        slot.d2h_done.synchronize()
        slot.busy = False
        return slot.out_host

    def __call__(self, q_host: torch.Tensor) -> torch.Tensor:
        return self.result(self.submit(q_host))

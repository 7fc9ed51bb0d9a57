"""Full-mode predict with the support bank sharded across the GPUs of one box (new; the reference has no
multi-GPU code — SURVEY.md 2.3, 8e).

Partitioning: class-aligned contiguous ranges of the class-sorted bank; rank r owns classes
[r*C/R, (r+1)*C/R).  Queries are replicated.  Each rank runs the fused forward on its shard and emits the
(B, C) class log-sum-exp table with -inf for classes it does not own; because every class is owned by
exactly one rank, ONE all-reduce(MAX) over NVLink merges the partial softmax state exactly.  Every rank
then finalises log(P + 1e-12) locally.

`merge_class_lse` is backend-agnostic (NCCL on GPUs, gloo in the CPU tests of the host logic).
"""
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


class ShardedBank:
    """This rank's class-aligned shard of a support bank + the merged forward."""

    def __init__(self, shard, group=None):
        self.shard = shard
        self.group = group

    @staticmethod
    def from_full(bank, group=None):
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        return ShardedBank(bank.class_shard(rank, world) if world > 1 else bank, group)

    def class_lse(self, q, scale: float = 1.0):
        return merge_class_lse(self.shard.class_lse(q, scale), self.group)

    def forward(self, q, scale: float = 1.0):
        from .bank import logp_from_class_lse

        return logp_from_class_lse(self.class_lse(q, scale))

"""CPU, world_size 2 over gloo: the bank-shard merge of nwhead_b200.dist (host logic of the multi-GPU
path).  Each rank computes the class-LSE table of ITS shard with the oracle; the merged table and the
final log-probs must equal the single-process answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nw_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, class_aligned, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nwhead_b200.dist import class_range, merge_class_lse

    rng = np.random.default_rng(0)
    C, per, d, B = 11, 9, 16, 6
    y = np.repeat(np.arange(C), per)
    s = rng.normal(size=(C * per, d))
    q = rng.normal(size=(B, d))
    sc = O.pairwise_scores(q, s, "euclidean")
    if class_aligned:
        lo, hi = class_range(rank, world, C)
        sel = (y >= lo) & (y < hi)
    else:
        sel = (np.arange(len(y)) % world) == rank  # rows of every class on every rank
    part = torch.from_numpy(O.class_lse(sc[:, sel], y[sel], C))
    merged = merge_class_lse(part, class_aligned=class_aligned).numpy()
    full = O.class_lse(sc, y, C)
    ok = np.allclose(merged, full, atol=1e-12) and np.allclose(O.logp_from_class_lse(merged),
                                                               O.nw_forward(q, s, y, C), atol=1e-10)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("class_aligned", [True, False])
def test_two_rank_merge_is_exact(class_aligned):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), class_aligned, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_class_ranges_partition_all_classes():
    from nwhead_b200.dist import class_range

    for C in (1000, 200, 7):
        for world in (1, 2, 4, 8):
            spans = [class_range(r, world, C) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))


class _FakeShard:
    """Stands in for a SupportBank shard in the host-side plan validation (no GPU, no kernels)."""

    def __init__(self, rows):
        self.rows, self.device, self.n_classes = rows, torch.device("cpu"), 4

    def __len__(self):
        return self.rows


def _plan_worker(rank, world, port, rows, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nwhead_b200.dist import ShardedBank

    try:
        ShardedBank(_FakeShard(rows[rank]), exchange="nccl")
        ret[rank] = "ok"
    except ValueError as e:
        ret[rank] = "ValueError" if "owns no support rows" in str(e) else repr(e)
    try:
        ShardedBank(_FakeShard(5), exchange="peer", max_batch=0)
        ret[f"peer{rank}"] = "ok"
    except ValueError as e:
        ret[f"peer{rank}"] = "ValueError" if "max_batch" in str(e) else repr(e)
    dist.destroy_process_group()


@pytest.mark.parametrize("rows,want", [((5, 7), "ok"), ((5, 0), "ValueError")])
def test_shard_plan_is_validated_on_every_rank(rows, want):
    """a rank that owns no support rows must make EVERY rank raise (the others would wait in the next
    collective forever), and the peer exchange must refuse max_batch = 0 instead of allocating empty tables."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_plan_worker, args=(world, _free_port(), rows, ret), nprocs=world, join=True)
    assert ret[0] == want and ret[1] == want
    assert ret["peer0"] == "ValueError" and ret["peer1"] == "ValueError"

"""GPU: the pipelined host-memory predictor (nwhead_b200.FullModePredictor) and, when two GPUs are
visible, the NCCL bank-sharded path against the single-GPU answer."""
import os
import socket

import numpy as np
import pytest
import torch

from gpu_util import clustered_features

pytestmark = pytest.mark.gpu


def test_predictor_matches_direct_call(cuda_lib):
    import nwhead_b200

    dev = "cuda:0"
    q, s, y, _ = clustered_features(40, 64, 256, 384, seed=11)
    bank = nwhead_b200.SupportBank.build(torch.from_numpy(s).to(dev), torch.from_numpy(y).to(dev), 40, "euclidean", "bf16")
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 40)
    want = head(torch.from_numpy(q).to(dev), bank).cpu()
    pred = nwhead_b200.FullModePredictor(bank, rows=384)
    qh = torch.from_numpy(q).pin_memory()
    assert torch.equal(pred(qh), want)
    # pipelined: several batches in flight, results come back in order and stay bitwise identical
    batches = [torch.from_numpy(np.roll(q, k, axis=0).copy()).pin_memory() for k in range(5)]
    tickets, outs = [], []
    for b in batches:
        tickets.append(pred.submit(b))
        if len(tickets) == 2:
            outs.append(pred.result(tickets.pop(0)).clone())
    while tickets:
        outs.append(pred.result(tickets.pop(0)).clone())
    for k, o in enumerate(outs):
        assert torch.equal(o, torch.roll(want, k, dims=0))
    with pytest.raises(RuntimeError, match="pipeline full"):
        pred.submit(qh), pred.submit(qh), pred.submit(qh)
    with pytest.raises(ValueError):
        nwhead_b200.FullModePredictor(bank, rows=384).submit(qh[:10])


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    import nwhead_b200
    from nwhead_b200.dist import ShardedBank

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    q, s, y, _ = clustered_features(30, 100, 128, 256, seed=3)
    full = nwhead_b200.SupportBank.build(torch.from_numpy(s).to(dev), torch.from_numpy(y).to(dev), 30, "euclidean", "bf16")
    want = full.forward(torch.from_numpy(q).to(dev))
    ok = True
    rows = 256 // world
    for exchange in ("nccl", "peer"):  # one all-reduce(MAX) vs in-kernel NVLink peer stores + signal barrier
        sharded = ShardedBank.from_full(full, exchange=exchange, max_batch=256)
        for _ in range(3):  # several steps: exercises the double-buffered peer tables
            got = sharded.forward(torch.from_numpy(q).to(dev))
            ok = ok and (got - want).abs().max().item() < 2e-5
            part = sharded.forward_rows(torch.from_numpy(q).to(dev))  # all-to-all: only this rank's rows
            ok = ok and (part - want[rank * rows:(rank + 1) * rows]).abs().max().item() < 2e-5
        pred = nwhead_b200.FullModePredictor(sharded, rows=rows)
        for _ in range(3):
            mine = pred(torch.from_numpy(q[rank * rows:(rank + 1) * rows]).pin_memory())
            ok = ok and (mine.to(dev) - want[rank * rows:(rank + 1) * rows]).abs().max().item() < 2e-5
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_gpu_bank_shards_match_single_gpu(cuda_lib):
    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_second_gpu_while_first_is_current(cuda_lib):
    """the reference works on any device through `.to(x.device)` and never calls set_device.  Everything
    here lives on cuda:1 while cuda:0 stays the current device; the library must launch on cuda:1 (on a side stream
    too) and give the bits of the cuda:0 run."""
    import nwhead_b200

    assert torch.cuda.current_device() == 0
    q, s, y, _ = clustered_features(12, 50, 128, 200, seed=5)
    outs = {}
    for dev in ("cuda:0", "cuda:1"):
        qt, st, yt = torch.from_numpy(q).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(y).to(dev)
        head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 12)
        bank = nwhead_b200.SupportBank.build(st, yt, 12, "euclidean", "bf16")
        with torch.no_grad():
            a = head(qt, bank)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                b = head(qt, bank)
            side.synchronize()
        q8 = qt[:8].clone().requires_grad_(True)
        s10 = st[:10].clone().requires_grad_(True)
        out = head(q8, s10, yt[:10])
        out.sum().backward()
        infl = nwhead_b200.support_influence(torch.softmax(a[:4], 1), torch.nn.functional.one_hot(yt[:4], 12).float(),
                                             torch.softmax(torch.randn(4, 600, generator=torch.Generator().manual_seed(1)),
                                                           1).to(dev),
                                             torch.nn.functional.one_hot(yt, 12).float())
        assert torch.cuda.current_device() == 0
        assert a.device == torch.device(dev) and torch.equal(a, b)
        outs[dev] = [t.detach().cpu() for t in (a, out, q8.grad, s10.grad, infl)]
    for x0, x1 in zip(outs["cuda:0"], outs["cuda:1"]):
        assert torch.equal(x0, x1)

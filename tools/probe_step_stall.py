"""Where does the host block while submitting the SECOND resident step after a synchronize?  (bench.py's per-step
trace: step 1 of every region is submitted 6-82 ms after step 0, later steps 0.1 ms apart.)  Developer probe."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bench  # noqa: E402
from nwhead_b200 import SupportBank, _abi  # noqa: E402
from nwhead_b200._abi import check, load, ptr, stream_of  # noqa: E402
from nwhead_b200.bank import logp_from_class_lse, rows_to_bf16  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
mu = bench.class_means(1000, 2048, dev)
feats, labels = bench.synth_bank(mu, 1280, dev)
bank = SupportBank.build(feats, labels, 1000, "euclidean", "bf16")
q, _ = bench.synth_queries(mu, 4096, dev)
lib = load()
n, b = len(bank), 4096


def step(stamps, ev=None, fresh=False):
    t = [time.perf_counter()]
    q_bf16, q_sq = bank.prepare_queries(q); t.append(time.perf_counter())
    if fresh:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    if ev is not None:
        ev[0].record()
    t.append(time.perf_counter())
    plan = _abi.forward_plan(b, n)
    side = torch.empty((int(plan.side_elems),), dtype=torch.float32, device=dev)
    out = torch.empty((b, 1000), dtype=torch.float32, device=dev); t.append(time.perf_counter())
    check(lib.nw_forward_class_lse(0, 1.0, ptr(q_bf16), ptr(q_sq), b, ptr(bank.feats_bf16), ptr(bank.sqnorm),
                                   ptr(bank.labels), n, bank.row_elems, 1000, ptr(out), ptr(side), side.numel(),
                                   stream_of(dev)), "fwd"); t.append(time.perf_counter())
    if ev is not None:
        ev[1].record()
    t.append(time.perf_counter())
    lp = logp_from_class_lse(out); t.append(time.perf_counter())
    stamps.append([(y - x) * 1e3 for x, y in zip(t, t[1:])])
    return lp


names = ["prepare_queries", "record e0", "plan+empty", "nw_forward_class_lse", "record e1", "logp"]
for mode in ("no events", "pre-created timing events", "fresh timing events", "pre-created timing events", "fresh timing events"):
    for _ in range(8):
        step([])
    pool = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
    for a, c in pool:
        a.record(); c.record()
    torch.cuda.synchronize()
    stamps = []
    t0 = time.perf_counter()
    for i in range(6):
        lp = step(stamps, pool[i] if mode.startswith("pre") else None, mode.startswith("fresh"))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{mode}: submit of 6 steps took {(t1 - t0) * 1e3:.1f} ms")
    for i, st in enumerate(stamps[:3]):
        print(f"   step {i}: " + ", ".join(f"{nm} {v:.2f}" for nm, v in zip(names, st)))

"""Kernel-only timing of nw_forward_class_lse for small query batches (developer probe)."""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from nwhead_b200 import SupportBank, _abi

dev = torch.device("cuda:0")
for spec in sys.argv[1:]:
    B, N, d, C = (int(v) for v in spec.split(","))
    feats = torch.relu(torch.randn(N, d, device=dev) + 0.5)
    labels = (torch.arange(N, device=dev) * C // N)
    bank = SupportBank.build(feats, labels, C, "euclidean", "bf16")
    q = torch.relu(torch.randn(B, d, device=dev) + 0.5)
    qb, qs = bank.prepare_queries(q)
    plan = _abi.forward_plan(B, N)
    for _ in range(5):
        bank.class_lse_prepared(qb, qs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    a.record()
    for _ in range(iters):
        bank.class_lse_prepared(qb, qs)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / iters * 1e3
    print(f"B={B} N={N} d={d}: {us:.1f} us/call, tiles/cta={plan.tiles_per_chunk} grid={plan.grid} pair={plan.cta_pair} "
          f"bank={N*d*2/1e6:.0f}MB  -> {N*d*2/us/1e3:.0f} GB/s", flush=True)

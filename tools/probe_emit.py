"""Kernel-only timing of nw_forward_emit (developer probe)."""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from nwhead_b200 import SupportBank, _abi
from nwhead_b200._abi import check, load, ptr, stream_of

dev = torch.device("cuda:0")
lib = load()
for spec in sys.argv[1:]:
    B, N, d, C = (int(v) for v in spec.split(","))
    feats = torch.relu(torch.randn(N, d, device=dev) + 0.5)
    labels = (torch.arange(N, device=dev) * C // N)
    bank = SupportBank.build(feats, labels, C, "euclidean", "bf16")
    q = torch.relu(torch.randn(B, d, device=dev) + 0.5)
    qb, qs = bank.prepare_queries(q)
    ld = (N + 3) // 4 * 4
    out = torch.empty((B, ld), device=dev)
    plan = _abi.forward_plan(B, N)

    import os
    KIND = int(os.environ.get("EMIT_KIND", "0"))

    def run():
        check(lib.nw_forward_emit(0, 1.0, ptr(qb), ptr(qs), B, ptr(bank.feats_bf16), ptr(bank.sqnorm), ptr(bank.labels), N,
                                  bank.row_elems, KIND, None, None, None, ptr(out), ld, stream_of(dev)), "emit")

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"emit B={B} N={N} d={d}: {ms:.3f} ms  out {B*N*4/ms/1e6:.0f} GB/s  {2.0*B*N*d/ms/1e9:.0f} TFLOP/s  "
          f"plan chunks={plan.chunks} tpc={plan.tiles_per_chunk} grid={plan.grid}", flush=True)
    lse_ms = None
    for _ in range(3):
        bank.class_lse_prepared(qb, qs)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        bank.class_lse_prepared(qb, qs)
    b.record()
    torch.cuda.synchronize()
    print(f"   class_lse same shape: {a.elapsed_time(b)/10:.3f} ms", flush=True)

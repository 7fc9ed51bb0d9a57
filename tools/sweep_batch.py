"""Query-batch sweep of full-mode predict on the config-3 bank (B = 256 / 1024 / 4096 / 16384, and
the small batches where the pass is HBM-bound): SupportBank.forward end to end on the device (query preparation +
fused forward + finalisation), CUDA-event timed.  Writes gpurun_out/batch_sweep.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from nwhead_b200 import SupportBank


def main():
    dev = torch.device("cuda:0")
    n, d, c = 1280000, 2048, 1000
    peaks = bench.measured_peaks()
    mu = bench.class_means(c, d, dev)
    feats, labels = bench.synth_bank(mu, n // c, dev)
    bank = SupportBank.build(feats, labels, c, "euclidean", "bf16")
    del feats
    out = []
    batches = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else (1, 8, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384)
    for b in batches:
        q, qy = bench.synth_queries(mu, b, dev)
        for _ in range(3):
            logp = bank.forward(q)
        iters = 30 if b <= 1024 else 10
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(iters):
            logp = bank.forward(q)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / iters
        row = {"B": b, "ms": round(ms, 4), "queries_per_s": round(b / ms * 1e3, 1),
               "tflops": round(2.0 * b * n * d / ms / 1e9, 1),
               "bank_GBs": round((n * d * 2 + n * 8) / ms / 1e6, 1),
               "top1_vs_generating_class": float((logp.argmax(1) == qy).float().mean())}
        out.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"workload": f"N={n} d={d} C={c} bf16 bank, SupportBank.forward on 1 B200", "peaks": peaks, "rows": out},
              open("gpurun_out/batch_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()

"""SupportBank.forward eager vs GraphedForward (one CUDA-graph launch per step) where predict is launch-bound:
the config-1 bank (5800 x 512, C=200, B=8) and small batches on the config-3 bank.  python tools/probe_graph.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from nwhead_b200 import SupportBank
from nwhead_b200.bank import GraphedForward


def timed(fn, iters):
    for _ in range(5):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters * 1e3


def main():
    dev = torch.device("cuda:0")
    for n, d, c, batches in ((5800, 512, 200, (8, 64)), (1280000, 2048, 1000, (1, 8, 128))):
        mu = bench.class_means(c, d, dev)
        feats, labels = bench.synth_bank(mu, n // c, dev)
        bank = SupportBank.build(feats, labels, c, "euclidean", "bf16")
        del feats
        for b in batches:
            q, _ = bench.synth_queries(mu, b, dev)
            graphed = GraphedForward(bank, b)
            assert torch.equal(graphed(q), bank.forward(q))
            row = {"N": n, "d": d, "C": c, "B": b, "eager_us": round(timed(lambda: bank.forward(q), 200), 1),
                   "graphed_us": round(timed(lambda: graphed(q), 200), 1)}
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

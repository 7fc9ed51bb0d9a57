"""One launch each of the round-1 late additions at full size, for an ncu capture:
K1 with the NW_EMIT_BLOCK_BEST epilogue (N=1.28M, d=2048, B=256), nw_kmeans_assign (k=3, config 4) and
nw_rounding_residual.  python tools/probe_new_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from nwhead_b200 import SupportBank
from nwhead_b200.utils import _kmeans_assign


def main():
    dev = torch.device("cuda:0")
    n, d, c = 1280000, 2048, 1000
    mu = bench.class_means(c, d, dev)
    feats, labels = bench.synth_bank(mu, n // c, dev)
    q, _ = bench.synth_queries(mu, 256, dev)
    bank = SupportBank.build(feats, labels, c, "euclidean", "bf16")
    for _ in range(2):
        bb, _ = bank.block_best(q)
        res = bank.rounding_residual(feats)
        cent = feats[::427][: c * 3].contiguous()
        assign, dist = _kmeans_assign(feats, labels.to(torch.int32), cent, 3)
    torch.cuda.synchronize()
    print("ok", float(bb.max()), float(res.max()), int(assign.max()), float(dist.mean()))


if __name__ == "__main__":
    main()

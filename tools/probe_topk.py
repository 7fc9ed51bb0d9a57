"""Exact top-k at scale: SupportBank.topk_exact (tensor-core block search + fp32 re-rank) against the dense fp32
ranking (nw_direct_scores + nw_rank_rows) on config-3-shaped synthetic data.  python tools/probe_topk.py [N d B k]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nwhead_b200 import SupportBank
from nwhead_b200.kernel import dense_scores
from nwhead_b200.utils import rank_rows
import bench


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) / reps * 1e3


def main():
    n, d, b, k = (int(a) for a in (sys.argv[1:5] + ["1280000", "2048", "256", "20"][len(sys.argv) - 1:]))
    dev = torch.device("cuda:0")
    n_classes = 1000
    mu = bench.class_means(n_classes, d, dev)
    feats, labels = bench.synth_bank(mu, n // n_classes, dev)
    q, _ = bench.synth_queries(mu, b, dev)
    for prec in ("bf16", "bf16x3"):
        bank = SupportBank.build(feats, labels, n_classes, "euclidean", prec)
        got, t_fast = timed(lambda: bank.topk_exact(q, k, feats))
        want, t_dense = timed(lambda: rank_rows(dense_scores("euclidean", q, feats), k), reps=1)
        bb, t_bb = timed(lambda: bank.block_best(q))
        if prec == "bf16":
            _, t_sc = timed(lambda: dense_scores("euclidean", q, feats), reps=1)
            print(f"dense scores alone {t_sc:.1f} ms = {2e-9 * b * n * d / t_sc:.1f} TFLOP/s (sub+fma per element)", flush=True)
        print(f"{prec}: N={n} d={d} B={b} k={k}  topk_exact {t_fast:.1f} ms (block_best {t_bb:.1f} ms)  "
              f"dense {t_dense:.1f} ms  equal={torch.equal(got, want)} paths={bank.last_topk_path}", flush=True)
        del bank


if __name__ == "__main__":
    main()

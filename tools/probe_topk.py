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
        # where the rest goes: block ranking, gathers, residuals, the refinement kernel
        from nwhead_b200._abi import check, load, ptr, stream_of
        lib = load()
        best, q_sq = bb
        nblk = best.shape[1]
        width = min(nblk, max(65, k))
        order, t_rank = timed(lambda: rank_rows(best, width))
        sorted_best, t_gather = timed(lambda: best.gather(1, order))
        resid_q, t_resid = timed(lambda: bank.rounding_residual(q))
        done = torch.zeros(b, dtype=torch.int32, device=dev)
        out = torch.empty((b, k), dtype=torch.int64, device=dev)
        pending = torch.zeros(1, dtype=torch.int32, device=dev)

        def refine():
            done.zero_()
            check(lib.nw_topk_refine(ptr(q), b, d, ptr(feats), n, ptr(bank.perm), ptr(order), ptr(sorted_best), width,
                                     min(64, nblk), nblk, k, ptr(q_sq), ptr(resid_q), ptr(bank._smax_sq),
                                     ptr(bank._resid_of[2]), bank.precision, ptr(done), ptr(out), ptr(pending),
                                     stream_of(dev)), "nw_topk_refine")

        _, t_refine = timed(refine)
        print(f"    breakdown: block ranking {t_rank:.2f} ms, gather {t_gather:.2f} ms, query residual {t_resid:.2f} ms, "
              f"refine kernel {t_refine:.2f} ms", flush=True)
        del bank


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — NW full-mode predict throughput (BASELINE.json config 3) on 1..8 B200.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU op sequence on the host cores

Workload ("step"): one batch of B=4096 synthetic ResNet-50-shaped queries (d=2048) against the
class-sorted 1.28M x 2048 support bank with C=1000 classes, euclidean kernel, full mode:
query prep -> fused tcgen05 forward (class log-sum-exp) -> [all-reduce MAX across ranks] -> log-probs.
With N GPUs the SAME bank is sharded class-aligned across ranks (strong scaling).

Prints ONE JSON line (rank 0).  `value` = queries/s with inputs resident in HBM; `e2e` = the same
through the public serving API (nwhead_b200.FullModePredictor) with queries in pinned host memory and the
(B, C) log-probs read back to the host every step (copies double-buffered against the compute).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NW full-mode queries/s (N=1.28M,d=2048) at 1/2/4/8 B200; % tensor peak"
CLASS_BLOCK = 5  # classes generated per RNG block (lets any rank rebuild exactly its own classes)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--n-support", type=int, default=1280000)
    ap.add_argument("--dim", type=int, default=2048)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: in-kernel NVLink peer stores + signal barrier, or one NCCL all-reduce(MAX)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"burst": p.get("bf16_tflops"), "sustained": p.get("bf16_tflops_sustained"), "hbm": p.get("hbm_gbs"),
                "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------
# synthetic data: S = relu(mu[y] + randn + 0.5), y = j // per_class
# ----------------------------------------------------------------------------------------------
def class_means(n_classes, d, dev):
    g = torch.Generator(device=dev).manual_seed(1234)
    return torch.randn(n_classes, d, generator=g, device=dev) * 0.6


def synth_shard(mu, c_lo, c_hi, per_class, dev):
    d = mu.shape[1]
    n = (c_hi - c_lo) * per_class
    feats = torch.empty((n, d), dtype=torch.float32, device=dev)
    row = 0
    for blk in range(c_lo // CLASS_BLOCK, (c_hi + CLASS_BLOCK - 1) // CLASS_BLOCK):
        b_lo, b_hi = blk * CLASS_BLOCK, min((blk + 1) * CLASS_BLOCK, mu.shape[0])
        g = torch.Generator(device=dev).manual_seed(100000 + blk)
        lab = torch.arange(b_lo, b_hi, device=dev).repeat_interleave(per_class)
        x = torch.relu(mu[lab] + torch.randn(len(lab), d, generator=g, device=dev) + 0.5)
        keep = (lab >= c_lo) & (lab < c_hi)
        x = x[keep]
        feats[row:row + len(x)] = x
        row += len(x)
    assert row == n
    labels = torch.arange(c_lo, c_hi, device=dev).repeat_interleave(per_class)
    return feats, labels


def synth_queries(mu, batch, dev):
    g = torch.Generator(device=dev).manual_seed(4321)
    qy = torch.randint(0, mu.shape[0], (batch,), generator=g, device=dev)
    return torch.relu(mu[qy] + torch.randn(batch, mu.shape[1], generator=g, device=dev) + 0.5), qy


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, c in self.rows:
            if len(c) < 7 or not any(a <= t <= b for a, b in windows):
                continue
            try:
                sm.append(float(c[0]))
                mx = max(mx, float(c[1]))
            except ValueError:
                continue
            for name, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_baseline(n_support, d, n_classes, seconds_budget=20.0):
    """The reference's CPU op sequence (oracle/torch_port.py) on a bounded sample of config 3:
    single queries against a class-balanced 1/8 sub-bank, scaled linearly in N (flagged)."""
    from oracle import torch_port as TP

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = max(n_classes, n_support // 8)
    per = max(1, n_s // n_classes)
    n_s = per * n_classes
    g = torch.Generator().manual_seed(7)
    y = torch.arange(n_classes).repeat_interleave(per)
    mu = torch.randn(n_classes, d, generator=g) * 0.6
    s = torch.relu(mu[y] + torch.randn(n_s, d, generator=g) + 0.5)
    q = torch.relu(mu[:4] + torch.randn(4, d, generator=g) + 0.5)
    TP.port_nw_forward(q[:1], s, y, n_classes, "euclidean")  # warm-up
    times, t_start = [], time.time()
    while len(times) < 5 or (time.time() - t_start < seconds_budget and len(times) < 24):
        t0 = time.perf_counter()
        TP.port_nw_forward(q[len(times) % 4:len(times) % 4 + 1], s, y, n_classes, "euclidean")
        times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    qps = (1.0 / t) * (n_s / n_support)
    return {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"B=1 query x N={n_s} supports (1/{n_support // n_s} of the bank), median of {len(times)} runs, "
                      f"scaled linearly in N to {n_support} (extrapolated); torch {torch.__version__} CPU, "
                      f"{torch.get_num_threads()} threads"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (fp32 torch op sequence of
    nwhead/nw.py:266-289 + kernel.py:13-15, restated in oracle/torch_port.py because the Python reference
    cannot travel to the GPU box).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import torch_port as TP

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_classes, d = args.classes, args.dim
    per = max(1, (args.n_support // 8) // n_classes)
    n_s = per * n_classes
    g = torch.Generator().manual_seed(7)
    y = torch.arange(n_classes).repeat_interleave(per)
    mu = torch.randn(n_classes, d, generator=g) * 0.6
    s = torch.relu(mu[y] + torch.randn(n_s, d, generator=g) + 0.5)
    q = torch.relu(mu[:8] + torch.randn(8, d, generator=g) + 0.5)
    for _ in range(max(1, min(args.warmup, 3))):
        TP.port_nw_forward(q[:1], s, y, n_classes, "euclidean")
    t0 = time.perf_counter()
    for i in range(args.steps):
        TP.port_nw_forward(q[i % 8:i % 8 + 1], s, y, n_classes, "euclidean")
    el = time.perf_counter() - t0
    qps = (args.steps / el) * (n_s / args.n_support)
    sample = (f"each step = B=1 query x N={n_s} supports (1/{args.n_support // n_s} of the bank); q/s scaled linearly "
              f"in N to {args.n_support} (extrapolated; the reference materialises (B,N,d) and (B,N,C))")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"NWHead full-mode inference N={args.n_support} d={d} C={n_classes} euclidean",
                   "batch": 1, "note": sample},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    import nwhead_b200
    from nwhead_b200 import SupportBank, _abi
    from nwhead_b200.bank import logp_from_class_lse
    from nwhead_b200.dist import ShardedBank, class_range

    _abi.check(_abi.load().nw_device_check(), "nw_device_check")
    B, N, d, C = args.batch, args.n_support, args.dim, args.classes
    per_class = N // C
    assert per_class * C == N, "n_support must be a multiple of classes"

    # ---- build this rank's class-aligned shard of the bank (same global data for every world size)
    mu = class_means(C, d, dev)
    c_lo, c_hi = class_range(rank, world, C)
    feats, labels = synth_shard(mu, c_lo, c_hi, per_class, dev)
    center = feats.sum(0, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(center)
    center = (center / N).float().contiguous()  # global centre: every shard rounds exactly like the 1-GPU bank
    bank = SupportBank.build(feats, labels, C, "euclidean", args.precision, center=center)
    del feats
    try:
        sharded = ShardedBank(bank, exchange=args.exchange, max_batch=B)
    except Exception as e:  # symmetric memory unavailable on this box: fall back to the NCCL all-reduce
        if rank == 0:
            print(f"bench: peer exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
        sharded = ShardedBank(bank, exchange="nccl")
    q_dev, qy = synth_queries(mu, B, dev)
    q_host = q_dev.cpu().pin_memory()
    torch.cuda.synchronize()
    plan = _abi.forward_plan(B, len(bank))

    from nwhead_b200.dist import merge_class_lse as sharded_merge

    rows = B // world
    assert rows * world == B, "batch must be a multiple of the number of GPUs"

    def step_resident():
        qb, qs = bank.prepare_queries(q_dev)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        # every rank finalises its own B/R rows of the merged table (the results stay sharded across ranks)
        if sharded.peer is None:
            lse = bank.class_lse_prepared(qb, qs)
            e1.record()
            lse = sharded_merge(lse)[rank * rows:(rank + 1) * rows]
        else:
            table, hdl, ptrs, ch = sharded.peer.next()
            bank.class_lse_prepared(qb, qs, tables=ptrs, rows_per_table=rows)
            e1.record()
            hdl.barrier(channel=ch)
            lse = table[rank * rows:(rank + 1) * rows]
        return logp_from_class_lse(lse), (e0, e1)

    # end-to-end arm: the public serving API.  Every step uploads the step's queries from pinned host memory
    # and reads the step's (rows, C) log-probs back; rank r moves rows [r*B/R, (r+1)*B/R) over PCIe and the
    # ranks all-gather the queries over NVLink.  Copies are double-buffered against the fused forward.
    predictor = nwhead_b200.FullModePredictor(sharded, rows)
    q_host_slice = q_host[rank * rows:(rank + 1) * rows].clone().pin_memory()

    def run_e2e(n_steps):
        tickets, last = [], None
        for _ in range(n_steps):
            tickets.append(predictor.submit(q_host_slice))
            if len(tickets) == 2:
                last = predictor.result(tickets.pop(0))
        while tickets:
            last = predictor.result(tickets.pop(0))
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    windows = []

    # ---- device-resident throughput
    for _ in range(max(args.warmup, 3)):
        logp, _ = step_resident()
    barrier()
    kernel_events = []
    w0 = time.time()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        logp, ev = step_resident()
        kernel_events.append(ev)
    s1.record()
    barrier()
    windows.append((w0, time.time()))
    total_ms = max_over_ranks(s0.elapsed_time(s1))
    kern_ms = max_over_ranks(statistics.mean(a.elapsed_time(b) for a, b in kernel_events))
    top1 = (logp.argmax(1) == qy[rank * rows:(rank + 1) * rows]).float().mean().item()
    psum = logp.exp().sum(1).mean().item()

    # ---- end to end through the public API with host buffers
    run_e2e(max(args.warmup, 3))
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    out_last = run_e2e(args.steps)
    torch.cuda.synchronize()
    e2e_local_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    windows.append((w0, time.time()))
    e2e_ms = max_over_ranks(e2e_local_ms)
    e2e_top1 = (out_last.argmax(1).to(dev) == qy[rank * rows:(rank + 1) * rows]).float().mean().item()
    time.sleep(0.15)
    sampler.stop()

    if rank == 0:
        peaks = measured_peaks()
        ms_step = total_ms / args.steps
        flops_launch = 2.0 * B * len(bank) * d  # per rank, per launch (2*N*d per query)
        achieved = flops_launch / (kern_ms * 1e-3) / 1e12
        peak_kind = "sustained" if total_ms > 1000.0 else "burst"
        peak = peaks[peak_kind]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": B * args.steps / (total_ms * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"NWHead full-mode inference: support N={N} d={d} C={C}, query batch B={B}, euclidean, "
                            f"bank sharded class-aligned over {world} GPU(s)",
                "precision": args.precision, "parallelism": f"bank-shard x{world}" if world > 1 else "single",
                "exchange": ("none" if world == 1 else "in-kernel NVLink peer stores (all-to-all by query row) + signal barrier"
                             if sharded.peer is not None else "one NCCL all-reduce(MAX)"),
                "l2": "inputs larger than L2: the bf16 bank shard streamed every step is "
                      f"{bank.feats_bf16.numel() * 2 / 1e9:.2f} GB",
                "plan": {"chunks": plan.chunks, "tiles_per_chunk": plan.tiles_per_chunk, "grid": plan.grid},
                "check": {"top1_vs_generating_class": top1, "mean_prob_sum": psum},
            },
            "e2e": {"value": B * args.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": q_host.numel() * 4, "d2h_bytes_per_step": B * C * 4,
                    "api": "nwhead_b200.FullModePredictor.submit/result (pinned host in, pinned host out, depth 2)",
                    "top1_vs_generating_class": e2e_top1},
            "gpu_launches": args.steps * (4 + (1 if plan.chunks > 1 else 0)),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "nw_forward_kernel<EUCLID> (+ its -inf fill and chunk-boundary merge launches)",
                         "kernel_ms": kern_ms, "flops_per_launch": flops_launch,
                         "peak_kind": f"{peak_kind} cuBLAS bf16, {peaks['source']}"},
            "clocks": sampler.summary(windows),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port_baseline(N, d, C)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

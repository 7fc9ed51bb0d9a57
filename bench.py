#!/usr/bin/env python
"""bench.py — NW full-mode predict throughput (BASELINE.json config 3) on 1..8 B200.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU op sequence on the host cores

Workload ("step"): one batch of B=4096 synthetic ResNet-50-shaped queries (d=2048; half of them placed between
two classes so that the posteriors are not all one-hot) against the class-sorted 1.28M x 2048 support bank with
C=1000 classes, euclidean kernel, full mode:
query prep -> fused tcgen05 forward (class log-sum-exp) -> [exchange across ranks] -> log-probs.
With N GPUs the SAME bank is sharded class-aligned across ranks (strong scaling).

Prints ONE JSON line (rank 0).
  value      queries/s with inputs resident in HBM, exactly --steps steps (the contract's number)
  e2e        the same through the public serving API (nwhead_b200.FullModePredictor): queries in pinned host
             memory, (B, C) log-probs read back to the host every step
  sustained  the resident step looped for >= --sustained-seconds at steady-state power, with the SM clock
             measured INSIDE the kernel (clock64 / globaltimer) and the tensor-pipe busy fraction at that clock
  roofline   the fused forward kernel against the measured cuBLAS bf16 peak (MEASURED_PEAKS.json)
  check      correctness gates evaluated outside the timed regions (the process exits 3 if one fails):
               N = 1: class probabilities / top-1 of the timed batch against a float64 restatement on the GPU
               N > 1: the merged log-probs of BOTH arms against the unsharded bank on rank 0
  alt        N > 1: the zero-communication alternative (bank replicated, queries sharded), SURVEY.md 8e
  aux        N = 1: BASELINE configs 2, 4, 5 (bounded), each next to the reference's op sequence on the host cores
  cpu_baseline  the reference's op sequence for THIS config on the host cores (N = 1)
"""
import argparse
import json
import math
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
CLASS_BLOCK = 5            # classes generated per RNG block
FLOP_PER_CLK_PER_SM = 8192  # dense bf16 tcgen05: 128x256x16 MMA per SM = 2*128*256*16 flop over 128 clk
LOGP_TOL_VS_UNSHARDED = 2e-5
PROB_TOL_VS_FP64 = 1e-3     # north_star tolerance


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
    ap.add_argument("--sustained-seconds", type=float, default=3.0)
    ap.add_argument("--warmup-seconds", type=float, default=1.0,
                    help="keep warming up (beyond --warmup steps) until this much wall time has passed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the bounded config 2/4/5 measurements (N=1)")
    ap.add_argument("--no-alt", action="store_true", help="skip the query-sharded / replicated-bank arm (N>1)")
    ap.add_argument("--fp64-queries", type=int, default=1024, help="queries checked against float64 (N=1)")
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


def synth_bank(mu, per_class, dev):
    """The whole class-sorted bank; generated class block by class block so that it does not depend on anything
    but the seeds (every world size sees the same 1.28M rows)."""
    n_classes, d = mu.shape
    feats = torch.empty((n_classes * per_class, d), dtype=torch.float32, device=dev)
    for blk in range((n_classes + CLASS_BLOCK - 1) // CLASS_BLOCK):
        b_lo, b_hi = blk * CLASS_BLOCK, min((blk + 1) * CLASS_BLOCK, n_classes)
        g = torch.Generator(device=dev).manual_seed(100000 + blk)
        lab = torch.arange(b_lo, b_hi, device=dev).repeat_interleave(per_class)
        feats[b_lo * per_class:b_hi * per_class] = torch.relu(
            mu[lab] + torch.randn(len(lab), d, generator=g, device=dev) + 0.5)
    labels = torch.arange(n_classes, device=dev).repeat_interleave(per_class)
    return feats, labels


def synth_queries(mu, batch, dev):
    """Even rows: drawn around one class mean.  Odd rows: 52 % / 48 % between two class means, so their
    posteriors are mixed (a broken shard merge cannot pass as top-1 = 100 % on one-hot answers)."""
    g = torch.Generator(device=dev).manual_seed(4321)
    c = mu.shape[0]
    qy = torch.randint(0, c, (batch,), generator=g, device=dev)
    other = torch.randint(0, c, (batch,), generator=g, device=dev)
    mix = torch.where(torch.arange(batch, device=dev) % 2 == 0, 0.0, 0.48).to(torch.float32)
    q = torch.relu((1 - mix)[:, None] * mu[qy] + mix[:, None] * mu[other]
                   + torch.randn(batch, mu.shape[1], generator=g, device=dev) + 0.5)
    return q, qy


# ----------------------------------------------------------------------------------------------
# checker (NOT the product): float64 restatement of nwhead/nw.py:266-289 + nwhead/kernel.py:13-15 with torch on
# the GPU, batched over queries and chunked over the bank.  |q|^2 + |s|^2 - 2 q.s in float64 carries ~1e-13 of
# cancellation error at these norms, far below the fp32 reference's own rounding.
# ----------------------------------------------------------------------------------------------
def fp64_class_probs(q, feats, labels, n_classes, chunk=16384):
    qd = q.double()
    qn = (qd * qd).sum(1, keepdim=True)
    m = torch.full((q.shape[0], 1), float("-inf"), dtype=torch.float64, device=q.device)
    w = torch.zeros((q.shape[0], n_classes), dtype=torch.float64, device=q.device)
    for i in range(0, feats.shape[0], chunk):
        s = feats[i:i + chunk].double()
        d2 = qn + (s * s).sum(1)[None, :] - 2.0 * (qd @ s.t())
        sc = -d2.clamp_min_(0).sqrt_()
        m_new = torch.maximum(m, sc.max(dim=1, keepdim=True).values)
        w *= (m - m_new).exp()
        w.index_add_(1, labels[i:i + chunk], (sc - m_new).exp_())
        m = m_new
    return w / w.sum(1, keepdim=True)


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled in-process through NVML every 5 ms DURING the timed regions
    (a 100 ms nvidia-smi loop never saw the 33 ms timed region of an 8-GPU run); nvidia-smi is the fallback."""
    REASON_BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                   0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}
    SMI_FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                  "clocks_event_reasons.sw_power_cap")

    def __init__(self, torch_index, period_s=0.005):
        self.index, self.period, self.rows = torch_index, period_s, []
        self.nvml, self.handle, self.proc, self.thread, self.stop_flag = None, None, None, None, False
        self.sm_max = None
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[torch_index]) if vis and vis.split(",")[torch_index].isdigit() else torch_index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.SMI_FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump_smi, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            t = time.time()
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                bits = int(reasons_fn(self.handle))
                self.rows.append((t, sm, pw, [name for bit, name in self.REASON_BITS.items() if bits & bit]))
            except Exception:
                pass
            time.sleep(self.period)

    def _pump_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                self.sm_max = float(c[1])
                self.rows.append((time.time(), float(c[0]), float(c[2]),
                                  [nm for nm, v in zip(names, c[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()

    def summary(self, windows):
        sel = [r for r in self.rows if any(a <= r[0] <= b for a, b in windows)]
        reasons = sorted({name for r in sel for name in r[3]})
        return {"sm_mhz": statistics.median(r[1] for r in sel) if sel else None, "sm_max_mhz": self.sm_max,
                "power_w": statistics.median(r[2] for r in sel) if sel else None, "reasons": reasons,
                "samples": len(sel), "source": "nvml in-process, 5 ms" if self.nvml is not None else "nvidia-smi -lms 20"}


class ClockProbe:
    """SM clock measured inside the fused forward kernel (nw_forward_set_clock_probe): per CTA, SM cycles and
    nanoseconds of its epilogue role, accumulated over the launches of a region."""

    def __init__(self, abi, dev, n_ctas):
        self.abi, self.n = abi, n_ctas
        self.buf = torch.zeros((2 * n_ctas,), dtype=torch.int64, device=dev)

    def __enter__(self):
        self.buf.zero_()
        torch.cuda.synchronize()
        if os.environ.get("NW_BENCH_CLOCK_PROBE", "on") != "off":
            self.abi.check(self.abi.load().nw_forward_set_clock_probe(self.abi.ptr(self.buf), self.n), "set_clock_probe")
        return self

    def __exit__(self, *exc):
        self.abi.load().nw_forward_set_clock_probe(None, 0)

    def mhz(self):
        torch.cuda.synchronize()
        v = self.buf.view(-1, 2).double().cpu()
        ok = v[:, 1] > 0
        if not bool(ok.any()):
            return None
        f = (v[ok, 0] / v[ok, 1] * 1e3)
        return {"median": float(f.median()), "min": float(f.min()), "max": float(f.max()), "ctas": int(ok.sum())}


# ----------------------------------------------------------------------------------------------
# the reference's CPU op sequence for config 3 (oracle/torch_port.py — checker / baseline infrastructure)
# ----------------------------------------------------------------------------------------------
def host_bank_fits(n_support, d, n_classes):
    """The reference at B=1 needs the fp32 bank + the int64 and fp32 one-hot matrices + cdist's operands."""
    need = n_support * d * 4 * 2.2 + n_support * n_classes * 12 * 1.2
    try:
        import psutil

        return psutil.virtual_memory().available > need + (8 << 30), need
    except Exception:
        return False, need


def host_bank(n_support, d, n_classes, full):
    """Class-sorted synthetic bank in host memory, same distribution as the device bank.  One noise block of 1/8
    of the rows is drawn and shared by 8 groups of classes (each row still gets its own class mean): drawing
    2.6 G normals serially on the host would take longer than the measurement itself."""
    per = max(1, (n_support if full else n_support // 8) // n_classes)
    n_s = per * n_classes
    g = torch.Generator().manual_seed(7)
    mu = torch.randn(n_classes, d, generator=g) * 0.6
    y = torch.arange(n_classes).repeat_interleave(per)
    s = torch.empty((n_s, d), dtype=torch.float32)
    blk = -(-n_s // 8) if full else n_s
    noise = torch.randn(blk, d, generator=g) + 0.5
    for i in range(0, n_s, blk):
        j = min(i + blk, n_s)
        torch.add(mu[y[i:j]], noise[:j - i], out=s[i:j])
    s.relu_()
    q = torch.relu(mu[:8] + torch.randn(8, d, generator=g) + 0.5)
    return q, s, y, n_s


def cpu_port_baseline(n_support, d, n_classes, seconds_budget=20.0):
    """B=1 queries through the reference's op sequence on the host cores.  Against the FULL bank when host memory
    allows, else against a class-balanced 1/8 sub-bank scaled linearly in N (flagged)."""
    from oracle import torch_port as TP

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full, need = host_bank_fits(n_support, d, n_classes)
    q, s, y, n_s = host_bank(n_support, d, n_classes, full)
    TP.port_nw_forward(q[:1], s, y, n_classes, "euclidean")  # warm-up
    times, t_start = [], time.time()
    while len(times) < 3 or (time.time() - t_start < seconds_budget and len(times) < 24):
        t0 = time.perf_counter()
        TP.port_nw_forward(q[len(times) % 8:len(times) % 8 + 1], s, y, n_classes, "euclidean")
        times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    qps = (1.0 / t) * (n_s / n_support)
    how = (f"the FULL bank N={n_s} (same config, not extrapolated)" if n_s == n_support else
           f"N={n_s} supports (1/{n_support // n_s} of the bank; host memory short of {need / 2**30:.0f} GiB), "
           f"scaled linearly in N to {n_support} (extrapolated)")
    return {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "same_config": n_s == n_support,
            "sample": f"B=1 query x {how}, median of {len(times)} runs; torch {torch.__version__} CPU, "
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
    full, need = host_bank_fits(args.n_support, d, n_classes)
    q, s, y, n_s = host_bank(args.n_support, d, n_classes, full)
    for _ in range(max(1, min(args.warmup, 3))):
        TP.port_nw_forward(q[:1], s, y, n_classes, "euclidean")
    t0 = time.perf_counter()
    for i in range(args.steps):
        TP.port_nw_forward(q[i % 8:i % 8 + 1], s, y, n_classes, "euclidean")
    el = time.perf_counter() - t0
    qps = (args.steps / el) * (n_s / args.n_support)
    if n_s == args.n_support:
        sample = f"each step = B=1 query x the FULL bank N={n_s} (same config, not extrapolated)"
    else:
        sample = (f"each step = B=1 query x N={n_s} supports (1/{args.n_support // n_s} of the bank); q/s scaled linearly "
                  f"in N to {args.n_support} (extrapolated; host memory short of {need / 2**30:.0f} GiB)")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"NWHead full-mode inference N={args.n_support} d={d} C={n_classes} euclidean",
                   "batch": 1, "note": sample + "; the reference materialises (B,N,d) and (B,N,C), so B=1"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                         "same_config": n_s == args.n_support},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def gpu_ms(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def aux_configs(feats, labels, bank, mu, dev, peaks):
    """BASELINE configs 4, 5 and 2 (bounded: a few seconds in total), each beside the reference's op sequence on the
    host cores in the same run.  HBM-bound kernels stream inputs far larger than L2 (10.5 GB, 4 GB)."""
    import nwhead_b200
    from nwhead_b200 import SupportBank
    from nwhead_b200.metric import support_influence_from_labels
    from nwhead_b200.utils import class_centroids
    from oracle import torch_port as TP

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, d = feats.shape
    c = bank.n_classes
    out = {}
    # ---- config 4: per-class centroid reduction over the fp32 bank + cluster-mode predict
    ms = gpu_ms(lambda: class_centroids(feats, None, bank.offsets, c), 5)
    alg = n * d * 4 + n * 4 + c * d * 4
    cent, cy = class_centroids(feats, None, bank.offsets, c)
    cbank = SupportBank.build(cent, cy, c, "euclidean", "bf16")
    q, _ = synth_queries(mu, 4096, dev)
    ms_pred = gpu_ms(lambda: cbank.forward(q), 10)
    sample_c = min(c, 50)
    per = n // c
    hs, hy = feats[:sample_c * per].cpu(), labels[:sample_c * per].cpu()
    t0 = time.perf_counter()
    TP.port_compute_clusters(hs, hy, 1)
    cpu_s = (time.perf_counter() - t0) * (c / sample_c)
    out["cfg4_cluster"] = {
        "workload": f"class centroids over N={n} d={d} C={c} fp32 (nw_class_centroids) + predict B=4096 vs {c} centroids",
        "centroid_ms": ms, "algorithmic_GB": alg / 1e9, "achieved_GBs": alg / ms / 1e6, "peak_GBs": peaks["hbm"],
        "frac_of_hbm_peak": alg / ms / 1e6 / peaks["hbm"], "predict_ms": ms_pred,
        "predict_queries_per_s": 4096 / ms_pred * 1e3,
        "cpu_baseline": {"centroid_s": cpu_s, "cores": cores, "kind": "port",
                         "sample": f"the reference's per-class sklearn KMeans(1, random_state=0) loop "
                                   f"(nwhead/utils.py:227-231) on {sample_c} of {c} classes, scaled linearly"}}
    del cent, cbank, hs, hy
    # ---- config 5: support_influence with the weights given (the reference's signature)
    b5, n5, c5 = 10000, 50000, 200
    g = torch.Generator(device=dev).manual_seed(5)
    w = torch.softmax(torch.randn(b5, n5, generator=g, device=dev), dim=-1)
    sy = (torch.arange(n5, device=dev) // (n5 // c5)).to(torch.int32)
    p = torch.zeros(b5, c5, device=dev).index_add_(1, sy.long(), w)
    qy = torch.randint(0, c5, (b5,), generator=g, device=dev).to(torch.int32)
    ms = gpu_ms(lambda: support_influence_from_labels(p, qy, w, sy), 5)
    alg = b5 * n5 * 8 + b5 * c5 * 4 + n5 * 4
    hb = 64
    pc, wc = p[:hb].cpu(), w[:hb].cpu()
    qoh = torch.nn.functional.one_hot(qy[:hb].long().cpu(), c5).float()
    soh = torch.nn.functional.one_hot(sy.long().cpu(), c5).float()
    t0 = time.perf_counter()
    TP.port_support_influence(pc, qoh, wc, soh)
    cpu_pairs = hb * n5 / (time.perf_counter() - t0)
    out["cfg5_influence"] = {
        "workload": f"support_influence B={b5} N={n5} C={c5}, weights given (nw_support_influence)", "ms": ms,
        "pairs_per_s": b5 * n5 / ms * 1e3, "algorithmic_GB": alg / 1e9, "achieved_GBs": alg / ms / 1e6,
        "peak_GBs": peaks["hbm"], "frac_of_hbm_peak": alg / ms / 1e6 / peaks["hbm"],
        "cpu_baseline": {"pairs_per_s": cpu_pairs, "cores": cores, "kind": "port",
                         "sample": f"the reference's per-query loop (util/metric.py:37-50) on {hb} of {b5} queries"}}
    del w, p, pc, wc, qoh, soh
    # ---- config 2: episodic head forward + backward (latency-bound: 82 kFLOP)
    g = torch.Generator(device=dev).manual_seed(2)
    sy2 = torch.randperm(200, generator=g, device=dev)[:10]
    qy2 = sy2[torch.randint(0, 10, (8,), generator=g, device=dev)]
    s0 = torch.relu(torch.randn(10, 512, generator=g, device=dev) + 0.5)
    q0 = torch.relu(torch.randn(8, 512, generator=g, device=dev) + 0.5)
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 200)

    def step(fwd, q_in, s_in, sy_in, qy_in):
        qq = q_in.clone().requires_grad_(True)
        ss = s_in.clone().requires_grad_(True)
        torch.nn.functional.nll_loss(fwd(qq, ss, sy_in), qy_in).backward()

    def wall_us(fn, iters=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e6

    class _NoopHead(torch.autograd.Function):  # what torch's own machinery costs around a two-input custom op
        @staticmethod
        def forward(ctx, a, b_):
            ctx.save_for_backward(a, b_)
            return out_static

        @staticmethod
        def backward(ctx, g_):
            a, b_ = ctx.saved_tensors
            return ga_static, gb_static

    out_static = torch.zeros(8, 200, device=dev)
    ga_static, gb_static = torch.zeros_like(q0), torch.zeros_like(s0)
    floor = wall_us(lambda: step(lambda a, b_, c_: _NoopHead.apply(a, b_), q0, s0, sy2, qy2))
    ours = wall_us(lambda: step(head, q0, s0, sy2, qy2))
    eager = wall_us(lambda: step(lambda a, b_, c_: TP.port_nw_forward(a, b_, c_, 200, "euclidean"), q0, s0, sy2, qy2))
    qc, sc, syc, qyc = q0.cpu(), s0.cpu(), sy2.cpu(), qy2.cpu()
    t0 = time.perf_counter()
    for _ in range(200):
        step(lambda a, b_, c_: TP.port_nw_forward(a, b_, c_, 200, "euclidean"), qc, sc, syc, qyc)
    cpu_us = (time.perf_counter() - t0) / 200 * 1e6
    out["cfg2_episodic_head"] = {
        "workload": "NW head forward+backward B=8 N=10 d=512 C=200 (direct fp32 kernels), wall time per step in a "
                    "tight loop incl. autograd, 2 clones and nll_loss",
        "wall_us": ours, "torch_gpu_unfused_wall_us": eager, "autograd_glue_floor_wall_us": floor,
        "note": "latency-bound: a tensor-peak fraction is meaningless at 82 kFLOP.  The floor is the same loop with "
                "a custom autograd op that launches nothing (2 clones, nll_loss forward/backward, the autograd "
                "engine, 2 gradient accumulations): host time no head implementation can remove",
        "cpu_baseline": {"wall_us": cpu_us, "cores": cores, "kind": "port",
                         "sample": "the same step through the reference's op sequence on CPU tensors, 200 iterations"}}
    # ---- large-support training step: forward + backward (grad_q) of B=1024 queries against the whole bank on the
    # tensor cores (nwhead_b200/backward.py; the reference differentiates NWHead.forward with autograd at any N)
    try:
        b_t = 1024
        qt, qyt = synth_queries(mu, b_t, dev)
        head_t = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), c, backward_path="tensor")

        def train_step():
            qq = qt.clone().requires_grad_(True)
            torch.nn.functional.nll_loss(head_t(qq, feats, labels), qyt).backward()
            return qq.grad

        ms_t = gpu_ms(train_step, 5, warm=2)
        flop_t = 3 * 2.0 * b_t * n * d  # forward, score recompute, W @ S
        nb = 160000
        hs, hy = feats[:nb].cpu(), labels[:nb].cpu()
        hq = qt[:1].cpu()
        t0 = time.perf_counter()
        for _ in range(2):
            qq = hq.clone().requires_grad_(True)
            torch.nn.functional.nll_loss(TP.port_nw_forward(qq, hs, hy, c, "euclidean"), qyt[:1].cpu()).backward()
        cpu_s = (time.perf_counter() - t0) / 2
        out["large_support_backward"] = {
            "workload": f"NWHead forward + backward (grad wrt queries) B={b_t} vs the fixed N={n} d={d} support, tensor-core "
                        "path (class-LSE forward, coefficient recompute, split-K W @ S)",
            "ms": ms_t, "tflops": flop_t / ms_t / 1e9, "contractions": 3, "peak_tflops": peaks["burst"],
            "frac_of_tensor_peak": flop_t / ms_t / 1e9 / peaks["burst"] if peaks["burst"] else None,
            "cpu_baseline": {"s_per_query_scaled": cpu_s * (n / nb), "cores": cores, "kind": "port",
                             "sample": f"the reference's op sequence + torch autograd, B=1 query x N={nb} supports "
                                       f"(1/{n // nb} of the bank), forward+backward, scaled linearly in N"}}
        del hs, hy
        # config 3 itself as a training step: 4096 queries, gradients for the queries AND all 1.28M support rows
        q4, qy4 = synth_queries(mu, 4096, dev)
        feats.requires_grad_(True)

        def train_step_both():
            qq = q4.clone().requires_grad_(True)
            feats.grad = None
            torch.nn.functional.nll_loss(head_t(qq, feats, labels), qy4).backward()

        try:
            ms_b = gpu_ms(train_step_both, 3, warm=1)
            out["large_support_backward"]["both_gradients"] = {
                "workload": f"B=4096 vs N={n}: forward + grad_q + grad_s (the support is rebuilt and transposed every step)",
                "ms": ms_b, "tflops": 5 * 2.0 * 4096 * n * d / ms_b / 1e9, "contractions": 5}
        finally:
            feats.requires_grad_(False)
            feats.grad = None
        del head_t
    except Exception as e:
        out.setdefault("large_support_backward", {})["error"] = f"{type(e).__name__}: {e}"[:300]
    torch.cuda.empty_cache()
    return out


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
    from nwhead_b200.dist import ShardedBank
    from nwhead_b200.dist import merge_class_lse as sharded_merge

    _abi.check(_abi.load().nw_device_check(), "nw_device_check")
    B, N, d, C = args.batch, args.n_support, args.dim, args.classes
    per_class = N // C
    assert per_class * C == N, "n_support must be a multiple of classes"
    rows = B // world
    assert rows * world == B, "batch must be a multiple of the number of GPUs"

    # ---- the bank.  Every rank generates the same 1.28M rows and builds the UNSHARDED bf16 bank (5.24 GB): it is
    # the reference result for the N > 1 exactness check and the operand of the query-sharded alternative.  The
    # rank's class-aligned shard is a slice of it (same centring vector => same rounding as the 1-GPU bank).
    mu = class_means(C, d, dev)
    feats, labels = synth_bank(mu, per_class, dev)
    full_bank = SupportBank.build(feats, labels, C, "euclidean", args.precision)
    if world > 1:
        del feats
        bank = full_bank.class_shard(rank, world)
        torch.cuda.empty_cache()
    else:
        bank = full_bank
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

    def event_pairs(n):
        """Timing events, created AND recorded once before they are needed: the first record of a torch event is
        what creates the CUDA event, and creating one while the persistent forward kernel occupies every SM blocked
        the host for 2-80 ms on the second step of every timed region (profiles/r2_bench_warmup_trace.txt)."""
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in pairs:
            a.record()
            b.record()
        return pairs

    stamps = []  # NW_BENCH_TRACE: host time of the sub-calls of every resident step

    def step_resident(ev=None):
        t_a = time.perf_counter()
        qb, qs = bank.prepare_queries(q_dev)
        t_b = time.perf_counter()
        e0, e1 = ev if ev is not None else event_pairs(1)[0]
        e0.record()
        # every rank finalises its own B/R rows of the merged table (the results stay sharded across ranks)
        if sharded.peer is None:
            t_c = time.perf_counter()
            lse = bank.class_lse_prepared(qb, qs)
            t_d = time.perf_counter()
            e1.record()
            lse = sharded_merge(lse)[rank * rows:(rank + 1) * rows]
            if trace_on:
                out = logp_from_class_lse(lse)
                stamps.append([(y - x) * 1e3 for x, y in ((t_a, t_b), (t_b, t_c), (t_c, t_d), (t_d, time.perf_counter()))])
                return out, (e0, e1)
        else:
            table, hdl, ptrs, ch = sharded.peer.next()
            bank.class_lse_prepared(qb, qs, tables=ptrs, rows_per_table=rows)
            e1.record()
            hdl.barrier(channel=ch)
            lse = table[rank * rows:(rank + 1) * rows]
        return logp_from_class_lse(lse), (e0, e1)

    # end-to-end arm: the public serving API.  Every step uploads the step's queries from pinned host memory
    # and reads the step's (rows, C) log-probs back; rank r moves rows [r*B/R, (r+1)*B/R) over PCIe and the
    # ranks replicate the queries over NVLink.  Copies are double-buffered against the fused forward.
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

    def gather_rows(t):
        """(rows, C) on every rank -> (B, C) on rank 0 (check only, outside the timed regions)."""
        t = t.to(dev).contiguous()
        if world == 1:
            return t
        out = torch.empty((B, C), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t)
        return out

    def timed_resident(n_steps):
        """n_steps resident steps between barriers, CUDA events on the launching stream, max over ranks."""
        events = event_pairs(n_steps)
        (s0, s1), = event_pairs(1)
        # Two untimed steps whose results are held together with whatever the caller still holds: the caching
        # allocator then owns every block the loop below will ask for.  Without them the SECOND step of a region
        # needed one block more than any step before it, and that cudaMalloc — issued while the persistent forward
        # kernel occupied every SM — blocked the host for 2-150 ms: one 70 ms "step" in a 20-step window
        # (profiles/r2_bench_warmup_trace.txt).  Timing events are pre-created for the same reason.
        held = [step_resident(events[i])[0] for i in range(min(2, n_steps))]
        del held
        for _ in range(8):  # back to the operating point after that stall (the GPU idled through it)
            step_resident(events[0])
        barrier()
        w0 = time.time()
        s0.record()
        host_t = []
        for i in range(n_steps):
            out, _ = step_resident(events[i])
            host_t.append(time.perf_counter())
        s1.record()
        barrier()
        win = (w0, time.time())
        total = max_over_ranks(s0.elapsed_time(s1))
        per_step = [a.elapsed_time(b) for a, b in events]
        kern = max_over_ranks(statistics.mean(per_step))
        trace.append({"steps": n_steps, "window": win, "kernel_ms_per_step": per_step,
                      "host_submit_ms": [(t - host_t[0]) * 1e3 for t in host_t[:64]]})
        return out, total, kern, win

    trace_on = bool(os.environ.get("NW_BENCH_TRACE"))
    trace = []  # per-step kernel times of every timed region (NW_BENCH_TRACE=<file> dumps them with the NVML samples)
    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("NW_BENCH_SAMPLER", "on") != "off":  # (developer switch: is the sampler intrusive?)
        sampler.start()
    probe = ClockProbe(_abi, dev, plan.grid)

    # ---- device-resident throughput: the contract's number, exactly --steps steps
    # Warm-up: at least --warmup (>= 3) steps AND at least --warmup-seconds of back-to-back steps, so that the
    # timed region starts at the GPU's operating point (memory and SM clocks up, power at its cap) and not
    # somewhere on the ramp from idle; the count actually run is what the line reports as "warmup".
    warm_events, warm_t0, n_warm = [], time.perf_counter(), 0
    while True:
        for _ in range(max(args.warmup, 3) if n_warm == 0 else 4):
            logp, ev = step_resident()
            warm_events.append(ev)
            n_warm += 1
        torch.cuda.synchronize()
        more = torch.tensor([1 if time.perf_counter() - warm_t0 < args.warmup_seconds else 0], device=dev)
        if world > 1:  # every rank must run the same number of steps (the exchange is collective)
            dist.all_reduce(more, op=dist.ReduceOp.MAX)
        if not int(more.item()):
            break
    trace.append({"steps": n_warm, "window": None, "what": "warm-up",
                  "kernel_ms_per_step": [a.elapsed_time(b) for a, b in warm_events]})
    with probe:
        logp, total_ms, kern_ms, win_value = timed_resident(args.steps)
        mhz_value = probe.mhz()
    logp_resident = logp.clone()

    # ---- the same step at steady-state power: >= --sustained-seconds, SM clock measured in the kernel
    sus_steps = max(args.steps, int(math.ceil(args.sustained_seconds * 1e3 / (total_ms / args.steps))))
    with probe:
        _, sus_ms, sus_kern_ms, win_sus = timed_resident(sus_steps)
        mhz_sus = probe.mhz()

    # ---- end to end through the public API with host buffers
    run_e2e(max(args.warmup, 3))
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    out_last = run_e2e(args.steps)
    torch.cuda.synchronize()
    e2e_local_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    win_e2e = (w0, time.time())
    e2e_ms = max_over_ranks(e2e_local_ms)
    logp_e2e = out_last.clone()

    # ---- N > 1: the zero-communication alternative — bank replicated on every GPU, queries sharded
    alt = None
    if world > 1 and not args.no_alt:
        q_rows = q_dev[rank * rows:(rank + 1) * rows].contiguous()
        for _ in range(3):
            full_bank.forward(q_rows)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            logp_alt = full_bank.forward(q_rows)
        a1.record()
        barrier()
        alt_ms = max_over_ranks(a0.elapsed_time(a1))
        alt = {"query_sharded_replicated_bank": {
            "value": B * args.steps / (alt_ms * 1e-3), "unit": "queries/s", "ms_per_step": alt_ms / args.steps,
            "note": f"every GPU holds the whole {full_bank.feats_bf16.numel() * 2 / 1e9:.2f} GB bf16 bank and answers "
                    f"B/{world} = {rows} queries per step; no exchange.  Scales N only while the bank fits one GPU"}}
    time.sleep(0.05)
    sampler.stop()

    # ---- correctness gates (outside every timed region)
    check, failed = {}, []
    got_res, got_e2e = gather_rows(logp_resident), gather_rows(logp_e2e)
    if rank == 0:
        top1 = (got_res.argmax(1) == qy).float().mean().item()
        check["top1_vs_generating_class"] = top1
        check["mean_prob_sum"] = got_res.exp().sum(1).mean().item()
        check["e2e_equals_resident"] = bool(torch.equal(got_res, got_e2e))
        if world > 1:
            ref = full_bank.forward(q_dev)  # the same batch against the UNSHARDED bank, one GPU, no exchange
            for name, got in (("resident", got_res), ("e2e", got_e2e)):
                err = (got - ref).abs().max().item()
                perr = (got.exp() - ref.exp()).abs().max().item()
                check[f"max_abs_logp_vs_unsharded_{name}"] = err
                check[f"max_abs_prob_vs_unsharded_{name}"] = perr
                check[f"top1_agreement_vs_unsharded_{name}"] = (got.argmax(1) == ref.argmax(1)).float().mean().item()
                if not err <= LOGP_TOL_VS_UNSHARDED:
                    failed.append(f"{name}: max |logp - unsharded| = {err:.3e} > {LOGP_TOL_VS_UNSHARDED}")
            check["max_abs_logp_vs_unsharded"] = max(check["max_abs_logp_vs_unsharded_resident"],
                                                     check["max_abs_logp_vs_unsharded_e2e"])
            check["queries_checked"] = B
            pm = ref.exp().max(1).values
            check["queries_with_mixed_posterior"] = int(((pm > 0.1) & (pm < 0.9)).sum().item())
            if alt is not None:
                check["max_abs_logp_vs_unsharded_alt"] = (logp_alt - ref[:rows]).abs().max().item()
        else:
            nq = min(args.fp64_queries, B)
            ref_p = fp64_class_probs(q_dev[:nq], feats, labels, C)
            got_p = got_res[:nq].double().exp()
            perr = (got_p - ref_p).abs().max().item()
            agree = (got_p.argmax(1) == ref_p.argmax(1)).float().mean().item()
            pm = ref_p.max(1).values
            check.update({"prob_err_vs_fp64": perr, "top1_agreement_vs_fp64": agree, "queries_checked": nq,
                          "queries_with_mixed_posterior": int(((pm > 0.1) & (pm < 0.9)).sum().item())})
            if not perr <= PROB_TOL_VS_FP64:
                failed.append(f"class-probability error vs float64 {perr:.3e} > {PROB_TOL_VS_FP64}")
            if not agree >= 0.999:
                failed.append(f"top-1 agreement vs float64 {agree:.4f} < 0.999")
        check["passed"] = not failed

    if rank == 0:
        peaks = measured_peaks()
        ms_step = total_ms / args.steps
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        flops_launch = 2.0 * B * len(bank) * d  # per rank, per launch (2*N*d per query)
        achieved = flops_launch / (kern_ms * 1e-3) / 1e12
        peak_kind = "sustained" if total_ms > 1000.0 else "burst"
        peak = peaks[peak_kind]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch_by_gpus", {}).get(str(world))

        def pipe_frac(k_ms, mhz):
            if not mhz:
                return None
            return flops_launch / (k_ms * 1e-3 * mhz["median"] * 1e6 * n_sm * FLOP_PER_CLK_PER_SM)

        sus_tflops = flops_launch / (sus_kern_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": B * args.steps / (total_ms * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"NWHead full-mode inference: support N={N} d={d} C={C}, query batch B={B} (half of the "
                            f"queries between two classes), euclidean, bank sharded class-aligned over {world} GPU(s)",
                "precision": args.precision, "parallelism": f"bank-shard x{world}" if world > 1 else "single",
                "exchange": ("none" if world == 1 else "in-kernel NVLink peer stores (all-to-all by query row) + signal barrier"
                             if sharded.peer is not None else "one NCCL all-reduce(MAX)"),
                "l2": "inputs larger than L2: the bf16 bank shard streamed every step is "
                      f"{bank.feats_bf16.numel() * 2 / 1e9:.2f} GB",
                "plan": {"chunks": plan.chunks, "tiles_per_chunk": plan.tiles_per_chunk, "grid": plan.grid},
            },
            "check": check,
            "e2e": {"value": B * args.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": q_host.numel() * 4, "d2h_bytes_per_step": B * C * 4,
                    "api": "nwhead_b200.FullModePredictor.submit/result (pinned host in, pinned host out, depth 2)"},
            "gpu_launches": args.steps * (4 + (1 if plan.chunks > 1 else 0)),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "nw_forward_kernel<EUCLID> (+ its -inf fill and chunk-boundary merge launches)",
                         "kernel_ms": kern_ms, "kernel_ms_min_max": [min(trace[1]["kernel_ms_per_step"]),
                                                                     max(trace[1]["kernel_ms_per_step"])],
                         "flops_per_launch": flops_launch,
                         "peak_kind": f"{peak_kind} cuBLAS bf16, {peaks['source']}",
                         "sm_mhz_in_kernel": mhz_value,
                         "tensor_pipe_busy_at_that_clock": pipe_frac(kern_ms, mhz_value)},
            "sustained": {
                "seconds": sus_ms * 1e-3, "steps": sus_steps, "value": B * sus_steps / (sus_ms * 1e-3),
                "unit": "queries/s", "ms_per_step": sus_ms / sus_steps, "kernel_ms": sus_kern_ms,
                "tflops_per_gpu": sus_tflops, "peak": peaks["sustained"],
                "frac_of_sustained_peak": sus_tflops / peaks["sustained"] if peaks["sustained"] else None,
                "frac_of_burst_peak": sus_tflops / peaks["burst"] if peaks["burst"] else None,
                "sm_mhz_in_kernel": mhz_sus, "tensor_pipe_busy_at_that_clock": pipe_frac(sus_kern_ms, mhz_sus),
                "clocks": sampler.summary([win_sus]),
                "how": "the resident step looped back to back; SM clock = clock64 delta / globaltimer delta per CTA, "
                       "summed over the launches; tensor-pipe busy = FLOP / (kernel time x that clock x SMs x 8192)"},
            "clocks": sampler.summary([win_value, win_e2e]),
        }
        if alt is not None:
            line["alt"] = alt
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port_baseline(N, d, C)
        if world == 1 and not args.no_aux:
            try:
                line["aux"] = aux_configs(feats, labels, bank, mu, dev, peaks)
            except Exception as e:  # the secondary measurements must never cost the headline line
                line["aux"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(line), flush=True)
        if os.environ.get("NW_BENCH_TRACE"):
            with open(os.environ["NW_BENCH_TRACE"], "w") as f:
                json.dump({"regions": trace, "stamps": stamps, "stamp_names": ["prepare_queries", "events", "class_lse_prepared", "record+logp"], "nvml": [[r[0], r[1], r[2], r[3]] for r in sampler.rows]}, f)
        for msg in failed:
            print(f"bench: CHECK FAILED — {msg}", file=sys.stderr, flush=True)
    fail_flag = torch.tensor([1 if failed else 0], device=dev)
    if world > 1:
        dist.broadcast(fail_flag, 0)
        dist.destroy_process_group()
    if int(fail_flag.item()):
        sys.exit(3)


if __name__ == "__main__":
    main()

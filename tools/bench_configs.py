"""Secondary measurements for BASELINE.json configs 2, 4, 5 and the bank build (developer tool; the graded
line is bench.py).  Writes one JSON object per measurement to stdout and gpurun_out/bench_configs.json.

Timing: CUDA events on the launching stream, 3 warm-up + 10 timed iterations, inputs larger than L2 for the
HBM-bound kernels (bank 10.5 GB, influence 4 GB).  Peaks from MEASURED_PEAKS.json.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nwhead_b200  # noqa: E402
from nwhead_b200 import SupportBank, _abi  # noqa: E402
from nwhead_b200.metric import support_influence_from_labels  # noqa: E402
from nwhead_b200.utils import class_centroids  # noqa: E402
from oracle import torch_port as TP  # noqa: E402

DEV = torch.device("cuda:0")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
OUT = []


def timed(fn, iters=10, warm=3):
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


def emit(**kw):
    OUT.append(kw)
    print(json.dumps(kw), flush=True)


def synth(n, d, c, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    per = n // c
    mu = torch.randn(c, d, generator=g, device=DEV) * 0.6
    feats = torch.empty(n, d, device=DEV)
    for i in range(0, n, 1 << 16):
        j = min(i + (1 << 16), n)
        lab = (torch.arange(i, j, device=DEV) // per).clamp_max(c - 1)
        feats[i:j] = torch.relu(mu[lab] + torch.randn(j - i, d, generator=g, device=DEV) + 0.5)
    return feats, (torch.arange(n, device=DEV) // per).clamp_max(c - 1), mu


def cfg2_episodic():
    """B=8, n_way=10, n_shot=1, d=512, C=200: fused direct fwd+bwd vs the reference's op sequence on the GPU."""
    g = torch.Generator(device=DEV).manual_seed(2)
    sy = torch.randperm(200, generator=g, device=DEV)[:10]
    qy = sy[torch.randint(0, 10, (8,), generator=g, device=DEV)]
    s0 = torch.relu(torch.randn(10, 512, generator=g, device=DEV) + 0.5)
    q0 = torch.relu(torch.randn(8, 512, generator=g, device=DEV) + 0.5)
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 200)

    def ours():
        q = q0.clone().requires_grad_(True)
        s = s0.clone().requires_grad_(True)
        torch.nn.functional.nll_loss(head(q, s, sy), qy).backward()

    def eager():
        q = q0.clone().requires_grad_(True)
        s = s0.clone().requires_grad_(True)
        torch.nn.functional.nll_loss(TP.port_nw_forward(q, s, sy, 200, "euclidean"), qy).backward()

    t_ours, t_eager = timed(ours, 50, 10), timed(eager, 50, 10)
    emit(config="cfg2 episodic head fwd+bwd B=8 N=10 d=512 C=200", ours_us=t_ours * 1e3, torch_gpu_unfused_us=t_eager * 1e3,
         note="latency-bound (82 kFLOP): tensor-peak fraction is meaningless at this size; includes autograd + "
              "2 clones + nll_loss in both arms")


def cfg4_cluster(feats, labels, n_classes):
    n, d = feats.shape
    bank = SupportBank.build(feats, labels, n_classes, "euclidean", "bf16")
    ms = timed(lambda: class_centroids(feats, None, bank.offsets, n_classes))
    bytes_alg = n * d * 4 + n * 4 + n_classes * d * 4
    emit(config=f"cfg4 class centroids N={n} d={d} C={n_classes} (fp32 in)", ms=ms, algorithmic_GB=bytes_alg / 1e9,
         achieved_GBs=bytes_alg / ms / 1e6, peak_GBs=PEAKS["hbm_gbs"], frac=bytes_alg / ms / 1e6 / PEAKS["hbm_gbs"])
    cent, cy = class_centroids(feats, None, bank.offsets, n_classes)
    cbank = SupportBank.build(cent, cy, n_classes, "euclidean", "bf16")
    g = torch.Generator(device=DEV).manual_seed(4321)
    q = torch.relu(torch.randn(4096, d, generator=g, device=DEV) + 0.5)
    ms = timed(lambda: cbank.forward(q))
    emit(config=f"cfg4 cluster-mode predict B=4096 vs {len(cbank)} centroids d={d}", ms=ms, queries_per_s=4096 / ms * 1e3,
         tflops=2.0 * 4096 * len(cbank) * d / ms / 1e9)
    # n_shot_cluster = 3: per-class k-means for all 1000 classes at once (the reference runs 1000 sklearn fits)
    from nwhead_b200.utils import _kmeans_assign, kmeans_centroids
    lab32 = labels.to(torch.int32)
    t0 = time.perf_counter()
    c3, _ = kmeans_centroids(feats, lab32, None, bank.offsets, n_classes, 3)
    torch.cuda.synchronize()
    total_ms = (time.perf_counter() - t0) * 1e3
    ms = timed(lambda: _kmeans_assign(feats, lab32, c3, 3))
    emit(config=f"cfg4 k-means k=3 N={n} d={d} C={n_classes}", total_ms=total_ms, assign_pass_ms=ms,
         assign_GBs=(n * d * 4 + n * 12) / ms / 1e6, peak_GBs=PEAKS["hbm_gbs"],
         assign_frac=(n * d * 4 + n * 12) / ms / 1e6 / PEAKS["hbm_gbs"])
    return bank


def bank_build(feats, labels, n_classes):
    n, d = feats.shape
    ms = timed(lambda: SupportBank.build(feats, labels, n_classes, "euclidean", "bf16"), iters=5, warm=2)
    # mean pass reads N*d*4, conversion pass reads N*d*4 and writes N*d*2 (+ norms, labels)
    bytes_alg = 2 * n * d * 4 + n * d * 2 + n * 16
    emit(config=f"K0 bank build (labels + centre + bf16 rows + norms) N={n} d={d}", ms=ms, algorithmic_GB=bytes_alg / 1e9,
         achieved_GBs=bytes_alg / ms / 1e6, peak_GBs=PEAKS["hbm_gbs"], frac=bytes_alg / ms / 1e6 / PEAKS["hbm_gbs"])
    qs = torch.relu(torch.randn(4096, d, device=DEV) + 0.5)
    bank = SupportBank.build(feats[:4096], labels[:4096], n_classes, "euclidean", "bf16")
    ms = timed(lambda: bank.prepare_queries(qs))
    emit(config=f"query prep B=4096 d={d}", ms=ms)


def cfg5_influence():
    B, N, C = 10000, 50000, 200
    g = torch.Generator(device=DEV).manual_seed(5)
    w = torch.softmax(torch.randn(B, N, generator=g, device=DEV), dim=-1)
    sy = (torch.arange(N, device=DEV) // (N // C)).to(torch.int32)
    P = torch.zeros(B, C, device=DEV).index_add_(1, sy.long(), w)
    qy = torch.randint(0, C, (B,), generator=g, device=DEV).to(torch.int32)
    ms = timed(lambda: support_influence_from_labels(P, qy, w, sy))
    bytes_alg = B * N * 8 + B * C * 4 + N * 4
    emit(config=f"cfg5 support_influence B={B} N={N} C={C} (weights given)", ms=ms, pairs_per_s=B * N / ms * 1e3,
         algorithmic_GB=bytes_alg / 1e9, achieved_GBs=bytes_alg / ms / 1e6, peak_GBs=PEAKS["hbm_gbs"],
         frac=bytes_alg / ms / 1e6 / PEAKS["hbm_gbs"])
    # the reference's Python loop on the host cores, bounded sample
    Pc, wc = P[:64].cpu(), w[:64].cpu()
    qoh = torch.nn.functional.one_hot(qy[:64].long().cpu(), C).float()
    soh = torch.nn.functional.one_hot(sy.long().cpu(), C).float()
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    TP.port_support_influence(Pc, qoh, wc, soh)
    dt = time.perf_counter() - t0
    emit(config="cfg5 reference loop on host (port), 64 queries x 50000 supports", pairs_per_s=64 * N / dt,
         cores=os.cpu_count(), kind="port")


def cfg5_from_features():
    """support_influence computed FROM FEATURES (d=512) on the tensor cores: fused forward + emit pass."""
    B, N, C, d = 10000, 50000, 200, 512
    feats, labels, mu = synth(N, d, C, 55)
    g = torch.Generator(device=DEV).manual_seed(56)
    qy = torch.randint(0, C, (B,), generator=g, device=DEV)
    q = torch.relu(mu[qy] + torch.randn(B, d, generator=g, device=DEV) + 0.5)
    for prec in ("bf16x3", "bf16"):
        bank = SupportBank.build(feats, labels, C, "euclidean", prec)
        ms = timed(lambda: bank.support_influence(q, qy, source_order=False), iters=5, warm=2)
        k = 3 if prec == "bf16x3" else 1
        flops = 2 * 2.0 * B * N * d * k  # two GEMM passes
        emit(config=f"cfg5 support_influence from features B={B} N={N} d={d} C={C} ({prec})", ms=ms,
             pairs_per_s=B * N / ms * 1e3, tflops=flops / ms / 1e9, out_GB=B * N * 4 / 1e9,
             note="class-LSE pass + emit pass; the (B,N) weight matrix is never written; 4 B/pair of HBM output")
        ms = timed(lambda: bank.scores(q, source_order=False), iters=5, warm=2)
        emit(config=f"dense scores B={B} N={N} d={d} ({prec}) via nw_forward_emit", ms=ms,
             tflops=2.0 * B * N * d * k / ms / 1e9, out_GBs=B * N * 4 / ms / 1e6)


def cfg3_torch_gpu_unfused(feats, labels, n_classes):
    """The reference's own op sequence on the B200 through stock PyTorch (cuBLAS cdist + softmax + one-hot bmm).
    It materialises (B,N,d), so only B=1 fits comfortably."""
    q = torch.relu(torch.randn(1, feats.shape[1], device=DEV) + 0.5)
    try:
        ms = timed(lambda: TP.port_nw_forward(q, feats, labels, n_classes, "euclidean"), iters=5, warm=2)
        emit(config=f"cfg3 torch-gpu-unfused (reference op sequence on the B200) B=1 N={len(feats)}", ms=ms,
             queries_per_s=1e3 / ms)
    except RuntimeError as e:  # OOM
        emit(config="cfg3 torch-gpu-unfused", error=str(e)[:120])


class SynthImages(torch.utils.data.Dataset):
    """CUB-shaped synthetic dataset: 5994 images of 3x224x224, labels i % 200."""

    def __init__(self, n=5994, n_classes=200):
        self.targets = [i % n_classes for i in range(n)]

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(int(i))
        return torch.randn(3, 224, 224, generator=g), self.targets[i]


def cfg1_cfg2_with_resnet18():
    """Whole NWNet calls with a ResNet-18 backbone (torchvision, fc removed -> 512 features): config 1
    precompute + predict(mode) at batch 8, and the config 2 episodic training step (B=8, n_way=10, n_shot=1)."""
    try:
        import torchvision
    except ImportError:
        emit(config="cfg1/cfg2 with ResNet-18", error="torchvision not available")
        return
    import numpy as np

    feat = torchvision.models.resnet18(weights=None)
    feat.fc = torch.nn.Identity()
    ds = SynthImages()
    net = nwhead_b200.NWNet(feat, 200, support_dataset=ds, feat_dim=512, kernel_type="euclidean", n_shot=1, n_way=10,
                            device="cuda:0").to(DEV)
    net.eval()
    t0 = time.perf_counter()
    with torch.no_grad():
        net.precompute()
    torch.cuda.synchronize()
    emit(config="cfg1 NWNet.precompute() ResNet-18, 5994 synthetic 224x224 images, 200 classes", seconds=time.perf_counter() - t0,
         bank_rows=len(net.support_eval.full_bank), note="includes the host-side synthetic image generation")
    x = torch.randn(8, 3, 224, 224, device=DEV)
    with torch.no_grad():
        f = net.featurizer(x)
        t_feat = timed(lambda: net.featurizer(x), 20, 5)
        times = {mode: timed(lambda: net.predict(x, mode=mode), 20, 5) for mode in ("full", "cluster", "random")}
        t_feat = min(t_feat, timed(lambda: net.featurizer(x), 20, 5))  # (the first timing runs on a cold GPU)
        for mode in ("full", "cluster", "random"):
            ms = times[mode]
            emit(config=f"cfg1 NWNet.predict(mode='{mode}') batch 8, ResNet-18 on the B200", ms=ms, queries_per_s=8e3 / ms,
                 featurizer_ms=t_feat, head_ms=ms - t_feat)
        bank = net.support_eval.full_bank
        ms = timed(lambda: bank.forward(f), 50, 10)
        emit(config="cfg1 head only: bank 5800x512, batch 8 (fused forward + finalise)", ms=ms)
        fwd = bank.graphed(8)
        ms = timed(lambda: fwd(f), 50, 10)
        emit(config="cfg1 head only, CUDA-graph replay (SupportBank.graphed)", ms=ms)
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9, nesterov=True)
    y = torch.tensor([3, 17, 17, 42, 99, 150, 199, 0], device=DEV)
    np.random.seed(0)
    sup = net.support_train.get_support(y)  # sample once: the timing excludes host-side image generation
    sup = tuple(t.to(DEV) for t in sup)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.nll_loss(net(x, y, support_data=sup), y)
        loss.backward()
        opt.step()

    ms = timed(step, 20, 5)
    emit(config="cfg2 whole episodic training step (ResNet-18 fwd+bwd on 18 images, NW head fwd+bwd, SGD) B=8 n_way=10",
         ms=ms, note="backbone-dominated; head GPU time is 23 us of it (profiles/r1_episodic_launches_ours.txt)")


def main():
    _abi.check(_abi.load().nw_device_check(), "nw_device_check")
    if "--cfg4" in sys.argv:
        feats, labels, _ = synth(1280000, 2048, 1000, 1234)
        cfg4_cluster(feats, labels, 1000)
        return
    if "--resnet" in sys.argv:
        cfg1_cfg2_with_resnet18()
    cfg2_episodic()
    cfg5_influence()
    cfg5_from_features()
    torch.cuda.empty_cache()
    feats, labels, _ = synth(1280000, 2048, 1000, 1234)
    bank_build(feats, labels, 1000)
    cfg4_cluster(feats, labels, 1000)
    cfg3_torch_gpu_unfused(feats, labels, 1000)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w") as f:
        json.dump(OUT, f, indent=1)


if __name__ == "__main__":
    main()

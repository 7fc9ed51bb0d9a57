"""Quick device-side timing probe of the fused forward (developer tool, not the bench contract)."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from nwhead_b200 import SupportBank, _abi  # noqa: E402
from nwhead_b200.bank import logp_from_class_lse  # noqa: E402


def synth_bank(n, d, c, dev, seed=1234):
    g = torch.Generator(device=dev).manual_seed(seed)
    per = n // c
    mu = torch.randn(c, d, generator=g, device=dev) * 0.6
    feats = torch.empty(n, d, device=dev)
    step = 1 << 16
    for i in range(0, n, step):
        j = min(i + step, n)
        lab = torch.arange(i, j, device=dev) // per
        feats[i:j] = torch.relu(mu[lab.clamp_max(c - 1)] + torch.randn(j - i, d, generator=g, device=dev) + 0.5)
    labels = (torch.arange(n, device=dev) // per).clamp_max(c - 1)
    return feats, labels, mu


def main():
    dev = torch.device("cuda:0")
    shapes = [(4096, 163840, 2048, 1000), (4096, 1280000, 2048, 1000)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for B, N, d, C in shapes:
        feats, labels, mu = synth_bank(N, d, C, dev)
        t0 = time.time()
        bank = SupportBank.build(feats, labels, C, "euclidean", "bf16")
        torch.cuda.synchronize()
        t_build = time.time() - t0
        del feats
        g = torch.Generator(device=dev).manual_seed(4321)
        qy = torch.randint(0, C, (B,), generator=g, device=dev)
        q = torch.relu(mu[qy] + torch.randn(B, d, generator=g, device=dev) + 0.5)
        qb, qs = bank.prepare_queries(q)
        plan = _abi.forward_plan(B, N)
        for _ in range(2):
            lse = bank.class_lse_prepared(qb, qs)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        iters = 5
        # SM clock measured inside the kernel (clock64 / globaltimer per CTA, summed over the launches)
        probe = torch.zeros((2 * plan.grid,), dtype=torch.int64, device=dev)
        _abi.check(_abi.load().nw_forward_set_clock_probe(_abi.ptr(probe), plan.grid), "set_clock_probe")
        ev[0].record()
        for _ in range(iters):
            lse = bank.class_lse_prepared(qb, qs)
        ev[1].record()
        torch.cuda.synchronize()
        _abi.load().nw_forward_set_clock_probe(None, 0)
        pv = probe.view(-1, 2).double()
        mhz = float((pv[:, 0] / pv[:, 1].clamp_min(1) * 1e3).median())
        ms = ev[0].elapsed_time(ev[1]) / iters
        logp = logp_from_class_lse(lse)
        acc = (logp.argmax(1) == qy).float().mean().item()
        tf = 2.0 * B * N * d / (ms * 1e-3) / 1e12
        print(f"B={B} N={N} d={d} C={C}: {ms:.3f} ms/batch  {B / (ms * 1e-3):.0f} q/s  {tf:.1f} TFLOP/s  "
              f"SM {mhz:.0f} MHz -> {2.0 * B * N * d / (ms * 1e-3 * mhz * 1e6 * 148 * 8192):.3f} of the tensor pipe  "
              f"plan(chunks={plan.chunks}, tpc={plan.tiles_per_chunk}, grid={plan.grid})  build {t_build:.2f}s  "
              f"top1-vs-label {acc:.3f}  psum {logp.exp().sum(1).mean().item():.6f}", flush=True)
        del bank


if __name__ == "__main__":
    main()

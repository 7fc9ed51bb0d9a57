"""Tensor-core forward + backward of the NW head against a large shared support (nwhead_b200/backward.py): time per
phase with CUDA events, TFLOP/s against the 4 (grad_q only: 3) contractions of 2*B*N*d FLOP a backward consists of,
and the direct fp32 path on the same inputs where it finishes in reasonable time.  Developer probe.

    python tools/probe_tensor_backward.py [B,N,d,C[,grad_s] ...]
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import nwhead_b200  # noqa: E402
from nwhead_b200 import SupportBank  # noqa: E402
from nwhead_b200 import backward as BW  # noqa: E402
from nwhead_b200.bank import logp_from_class_lse  # noqa: E402

dev = torch.device("cuda:0")


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


class ClockProbe:
    """SM clock measured inside K1 (nw_forward_set_clock_probe): cycles / nanoseconds per CTA over the launches."""

    def __init__(self):
        from nwhead_b200 import _abi
        self.abi, self.n = _abi, 148
        self.buf = torch.zeros((2 * self.n,), dtype=torch.int64, device=dev)

    def __enter__(self):
        self.buf.zero_()
        torch.cuda.synchronize()
        self.abi.check(self.abi.load().nw_forward_set_clock_probe(self.abi.ptr(self.buf), self.n), "probe")
        return self

    def __exit__(self, *exc):
        torch.cuda.synchronize()
        self.abi.load().nw_forward_set_clock_probe(None, 0)
        v = self.buf.view(-1, 2).double().cpu()
        ok = v[:, 1] > 0
        self.mhz = float((v[ok, 0] / v[ok, 1] * 1e3).median()) if bool(ok.any()) else float("nan")


def phases(q0, s0, sy, c, want_s):
    """The steps of NWTensorFunction one by one (same calls), each bracketed by events."""
    t = {}
    clocks = {}
    a = ev()
    bank = SupportBank.build(s0, sy, c, "euclidean", "bf16")
    t["bank build"] = (a, ev())
    a = ev()
    q_bf16, q_sq = bank.prepare_queries(q0)
    lse = bank.class_lse_prepared(q_bf16, q_sq, 1.0)
    logp = logp_from_class_lse(lse)
    t["forward"] = (a, ev())
    g = torch.zeros_like(logp)
    g[torch.arange(len(q0)), torch.randint(0, c, (len(q0),), device=dev)] = -1.0 / len(q0)
    a = ev()
    row_lse, table = BW.backward_table(lse, g)
    t["table"] = (a, ev())
    for _ in range(3):  # back to back, SM clock measured in the kernel
        with ClockProbe() as cp:
            a = ev()
            w, rowsum = BW.coefficients(bank, q_bf16, q_sq, row_lse, table, 1.0, 0)
            t["coefficients W"] = (a, ev())
        clocks["coefficients W"] = cp.mhz
    with ClockProbe() as cp:
        bank.class_lse_prepared(q_bf16, q_sq, 1.0)
    clocks["forward"] = cp.mhz
    a = ev()
    st = BW.transpose_operand(bank.feats_bf16)
    t["transpose bank"] = (a, ev())
    a = ev()
    raw = BW.dense_products(w, st)
    t["grad_q GEMM"] = (a, ev())
    a = ev()
    BW.finish(raw, q_bf16, rowsum, None, bank.d)
    t["grad_q finish"] = (a, ev())
    del w, st, raw
    if want_s:
        a = ev()
        wt, colsum = BW.coefficients(bank, q_bf16, q_sq, row_lse, table, 1.0, 1)
        t["coefficients W^t"] = (a, ev())
        a = ev()
        BW.support_gradient(wt, q_bf16, bank, colsum)
        t["grad_s products (fused)"] = (a, ev())
        del wt
    torch.cuda.synchronize()
    out = {k: x.elapsed_time(y) for k, (x, y) in t.items()}
    out.update({f"MHz[{k}]": v for k, v in clocks.items()})
    return out


shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [
    (4096, 1280000, 2048, 1000, 1), (1024, 1280000, 2048, 1000, 0), (256, 1280000, 2048, 1000, 0),
    (4096, 160000, 2048, 1000, 1), (256, 160000, 2048, 1000, 1), (1024, 50000, 512, 200, 1)]
for shp in shapes:
    b, n, d, c = shp[:4]
    want_s = bool(shp[4]) if len(shp) > 4 else True
    g = torch.Generator(device=dev).manual_seed(0)
    per = n // c
    mu = torch.randn(c, d, generator=g, device=dev) * 0.6
    sy = torch.arange(n, device=dev) // per
    s0 = torch.empty(n, d, device=dev)
    for i in range(0, n, 65536):
        s0[i:i + 65536] = torch.relu(mu[sy[i:i + 65536]] + torch.randn(min(65536, n - i), d, generator=g, device=dev) + 0.5)
    qy = torch.randint(0, c, (b,), generator=g, device=dev)
    q0 = torch.relu(mu[qy] + torch.randn(b, d, generator=g, device=dev) + 0.5)

    def step(head):
        q = q0.clone().requires_grad_(True)
        s = s0.requires_grad_(want_s)
        s.grad = None
        torch.nn.functional.nll_loss(head(q, s, sy), qy).backward()
        return q.grad, s.grad

    head_t = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), c, backward_path="tensor")
    step(head_t)
    times = []
    for _ in range(4):  # (the first repetitions still grow the caching allocator: cudaMalloc under a running kernel stalls)
        torch.cuda.synchronize()
        a = ev()
        gq_t, gs_t = step(head_t)
        e = ev()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(e))
    ms = min(times)
    flop = 2.0 * b * n * d * (5 if want_s else 3)  # forward + (recompute + product) per gradient
    print(f"B={b} N={n} d={d} C={c} grad_s={int(want_s)}: tensor fwd+bwd {ms:.2f} ms (min of {[round(t, 1) for t in times]}) = {flop / ms / 1e9:.0f} TFLOP/s "
          f"over {5 if want_s else 3} contractions", flush=True)
    del gs_t
    s0.requires_grad_(False)
    phases(q0, s0, sy, c, want_s)
    ph = phases(q0, s0, sy, c, want_s)  # second pass: allocator warm
    print("   phases (ms): " + ", ".join(f"{k} {v:.2f}" for k, v in ph.items()), flush=True)
    if b * n <= 256 * 160000:
        head_d = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), c, backward_path="direct")
        step(head_d)
        torch.cuda.synchronize()
        a = ev()
        gq_d, gs_d = step(head_d)
        e = ev()
        torch.cuda.synchronize()
        err = (gq_t - gq_d).abs().max().item() / gq_d.abs().max().item()
        print(f"   direct fp32 path fwd+bwd {a.elapsed_time(e):.2f} ms; grad_q tensor vs direct: {err:.2e} of max-abs", flush=True)
        del gq_d, gs_d
    del s0, q0, gq_t
    torch.cuda.empty_cache()

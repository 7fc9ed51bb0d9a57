"""Differentiable NW head over a LARGE shared support (the reference trains through NWHead.forward at any N,
nwhead/nw.py:266-289): time forward+backward of the direct fp32 path and check it against torch autograd of the
reference's op sequence at a size the latter can hold.  Developer probe."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import nwhead_b200  # noqa: E402
from oracle import torch_port as TP  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for n, d, c, b in ((20000, 512, 200, 8), (160000, 2048, 1000, 8), (1280000, 2048, 1000, 8)):
    per = n // c
    mu = torch.randn(c, d, generator=g, device=dev) * 0.6
    sy = torch.arange(n, device=dev) // per
    s0 = torch.relu(mu[sy] + torch.randn(n, d, generator=g, device=dev) + 0.5)
    qy = torch.randint(0, c, (b,), generator=g, device=dev)
    q0 = torch.relu(mu[qy] + torch.randn(b, d, generator=g, device=dev) + 0.5)
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), c)

    def step(fwd):
        q = q0.clone().requires_grad_(True)
        s = s0.clone().requires_grad_(True) if n <= 160000 else s0
        torch.nn.functional.nll_loss(fwd(q, s, sy), qy).backward()
        return q.grad, (s.grad if s.requires_grad else None)

    for _ in range(2):
        gq, gs = step(head)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        gq, gs = step(head)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    msg = f"N={n} d={d} C={c} B={b}: fwd+bwd {ms:.2f} ms (grad_s: {gs is not None})"
    if n <= 20000:
        rq, rs = step(lambda a, b_, c_: TP.port_nw_forward(a, b_, c_, c, "euclidean"))
        msg += (f"  vs torch autograd: grad_q rel err {(gq - rq).abs().max().item() / rq.abs().max().item():.2e}, "
                f"grad_s rel err {(gs - rs).abs().max().item() / rs.abs().max().item():.2e}")
    print(msg, flush=True)
    del s0

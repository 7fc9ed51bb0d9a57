"""One tensor-core forward + backward step for an ncu capture (launch order of nw_forward_kernel inside it:
class-LSE forward, coefficients W, grad_q products, coefficients W^t, grad_s products)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import nwhead_b200  # noqa: E402

b, n, d, c = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096,320000,2048,1000").split(","))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
mu = torch.randn(c, d, generator=g, device=dev) * 0.6
sy = torch.arange(n, device=dev) // (n // c)
s0 = torch.relu(mu[sy] + torch.randn(n, d, generator=g, device=dev) + 0.5).requires_grad_(True)
qy = torch.randint(0, c, (b,), generator=g, device=dev)
q0 = torch.relu(mu[qy] + torch.randn(b, d, generator=g, device=dev) + 0.5).requires_grad_(True)
head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), c, backward_path="tensor")
torch.nn.functional.nll_loss(head(q0, s0, sy), qy).backward()
torch.cuda.synchronize()
print("ok", float(q0.grad.abs().max()), float(s0.grad.abs().max()))

"""One episodic head fwd+bwd (BASELINE config 2 shape) per arm, for the ncu launch list."""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import nwhead_b200
from oracle import torch_port as TP

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
sy = torch.randperm(200, generator=g, device=dev)[:10]
qy = sy[torch.randint(0, 10, (8,), generator=g, device=dev)]
s0 = torch.relu(torch.randn(10, 512, generator=g, device=dev) + 0.5)
q0 = torch.relu(torch.randn(8, 512, generator=g, device=dev) + 0.5)
head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 200)
arm = sys.argv[1]
for _ in range(3):
    q = q0.clone().requires_grad_(True)
    s = s0.clone().requires_grad_(True)
    out = head(q, s, sy) if arm == "ours" else TP.port_nw_forward(q, s, sy, 200, "euclidean")
    torch.nn.functional.nll_loss(out, qy).backward()
torch.cuda.synchronize()
print(arm, "ok")

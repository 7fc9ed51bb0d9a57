"""Config-4 cluster-mode predict (B=4096 queries vs 1000 centroids, d=2048, one centroid per class): what one
SupportBank.forward launches and how long each takes (run under ncu --metrics gpu__time_duration.sum)."""
import sys
import torch

sys.path.insert(0, ".")
from nwhead_b200 import SupportBank  # noqa: E402

DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)
N, d, C, B = 1000, 2048, 1000, 4096
y = torch.arange(N, device=DEV)
s = torch.relu(torch.randn(N, d, generator=g, device=DEV))
q = torch.relu(torch.randn(B, d, generator=g, device=DEV))
bank = SupportBank.build(s, y, C, "euclidean", "bf16")
for _ in range(3):
    out = bank.forward(q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    out = bank.forward(q)
e1.record()
torch.cuda.synchronize()
print(f"SupportBank.forward B={B} N={N} d={d} C={C}: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us per call")

"""Config-1 head call (B=8, N=5800, d=512, C=200): what one SupportBank.forward launches and how long each takes.
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel list; prints the event-timed call itself."""
import sys
import torch

sys.path.insert(0, ".")
from nwhead_b200 import SupportBank  # noqa: E402

DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)
N, d, C, B = 5800, 512, 200, 8
y = (torch.arange(N, device=DEV) % C).sort().values
s = torch.relu(torch.randn(N, d, generator=g, device=DEV))
q = torch.relu(torch.randn(B, d, generator=g, device=DEV))
bank = SupportBank.build(s, y, C, "euclidean", "bf16")
for _ in range(3):
    out = bank.forward(q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    out = bank.forward(q)
e1.record()
torch.cuda.synchronize()
print(f"SupportBank.forward B={B} N={N} d={d} C={C}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call")

for _ in range(4):
    out = bank.forward_auto(q)
torch.cuda.synchronize()
e0.record()
for _ in range(200):
    out = bank.forward_auto(q)
e1.record()
torch.cuda.synchronize()
print(f"SupportBank.forward_auto (CUDA-graph replay + clone): {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call")

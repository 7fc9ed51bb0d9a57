"""Class-probability error of the tensor-core path at config-3 scale vs a float64 restatement (developer probe)."""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from nwhead_b200 import SupportBank

DEV = torch.device("cuda:0")
N, D, C, B = 1280000, 2048, 1000, 128
g = torch.Generator(device=DEV).manual_seed(1234)
per = N // C
mu = torch.randn(C, D, generator=g, device=DEV) * 0.6
feats = torch.empty(N, D, device=DEV)
for i in range(0, N, 1 << 16):
    j = min(i + (1 << 16), N)
    lab = torch.arange(i, j, device=DEV) // per
    feats[i:j] = torch.relu(mu[lab] + torch.randn(j - i, D, generator=g, device=DEV) + 0.5)
labels = torch.arange(N, device=DEV) // per
qy = torch.randint(0, C, (B,), generator=g, device=DEV)
other = torch.randint(0, C, (B,), generator=g, device=DEV)
mix = torch.linspace(0.0, 0.5, B, device=DEV)  # queries between two classes: non-trivial posteriors
q = torch.relu((1 - mix)[:, None] * mu[qy] + mix[:, None] * mu[other] + torch.randn(B, D, generator=g, device=DEV) + 0.5)

scores = torch.empty((B, N), dtype=torch.float64, device=DEV)
for i in range(0, N, 1 << 15):
    blk = feats[i:i + (1 << 15)].double()
    for b in range(B):
        scores[b, i:i + blk.shape[0]] = -(blk - q[b].double()).norm(dim=1)
m = scores.max(dim=1, keepdim=True).values
w = torch.zeros((B, C), dtype=torch.float64, device=DEV).index_add_(1, labels, (scores - m).exp())
ref = w / w.sum(1, keepdim=True)
pm = ref.max(1).values; print("reference max-prob per query: min %.3f median %.3f; queries with 0.1 < pmax < 0.9: %d" % (pm.min(), pm.median(), int(((pm > 0.1) & (pm < 0.9)).sum())))
for prec, center in (("bf16", True), ("bf16", False), ("bf16x3", True)):
    bank = SupportBank.build(feats, labels, C, "euclidean", prec, use_center=center)
    got = bank.forward(q).double().exp()
    err = (got - ref).abs().max().item()
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"{prec:7s} center={center}: class-prob max-abs err {err:.3e}  top-1 agreement {agree:.3f}", flush=True)
    del bank

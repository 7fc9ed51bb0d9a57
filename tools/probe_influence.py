"""Component timing of SupportBank.support_influence (developer probe)."""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from nwhead_b200 import SupportBank, _abi

dev = torch.device("cuda:0")
B, N, d, C = 10000, 50000, 512, 200
feats = torch.relu(torch.randn(N, d, device=dev) + 0.5)
labels = (torch.arange(N, device=dev) * C // N)
bank = SupportBank.build(feats, labels, C, "euclidean", "bf16")
q = torch.relu(torch.randn(B, d, device=dev) + 0.5)
qy = torch.randint(0, C, (B,), device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


lse = bank.class_lse(q)
z = torch.logsumexp(lse, dim=1).contiguous()
p = torch.exp(lse.gather(1, qy[:, None])[:, 0] - z).contiguous()
qy32 = qy.to(torch.int32)
print("class_lse", timed(lambda: bank.class_lse(q)))
print("logsumexp+gather", timed(lambda: (torch.logsumexp(lse, dim=1), torch.exp(lse.gather(1, qy[:, None])[:, 0] - z))))
print("emit influence", timed(lambda: bank._emit(q, 1.0, _abi.EMIT_INFLUENCE, row_lse=z, p_query=p, qlabel=qy32)))
print("emit scores", timed(lambda: bank._emit(q, 1.0, _abi.EMIT_SCORES)))
print("total", timed(lambda: bank.support_influence(q, qy, source_order=False)))

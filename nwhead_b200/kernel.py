"""Similarity kernels — the API of the reference's nwhead/kernel.py, backed by libnw_sm100.

Each module carries a ``kind`` tag that NWHead uses to select the fused CUDA path; calling a module
directly (``kernel(x, y)``, as NWNet.get_neighbors does, reference nwhead/nw.py:248) returns the dense
score matrix computed by the direct fp32 kernel (nw_direct_scores).

Forward args (reference nwhead/kernel.py:6-11):
    x: (bs, num_x, embed_dim)   y: (bs, num_y, embed_dim)   ->  (bs, num_x, num_y)
2-D inputs (B, d), (N, d) -> (B, N) are accepted for every kernel.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _abi
from ._abi import KIND, check, load, ptr, stream_of


def dense_scores(kind: str, x: torch.Tensor, y: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """Score matrix through nw_direct_scores.  Not differentiable (NWHead owns the fused autograd)."""
    dev = _abi.require_cuda(x, y)
    lib = load()
    x = x.detach().float().contiguous()
    y = y.detach().float().contiguous()
    st = stream_of(dev)
    k = KIND[kind]
    if x.dim() == 2 and y.dim() == 2:
        b, d = x.shape
        n = y.shape[0]
        out = torch.empty((b, n), dtype=torch.float32, device=dev)
        check(lib.nw_direct_scores(k, scale, ptr(x), b, d, ptr(y), n, 0, ptr(out), st), "nw_direct_scores")
        return out
    if x.dim() != 3 or y.dim() != 3 or x.shape[0] != y.shape[0]:
        raise ValueError(f"expected (bs, num_x, d) and (bs, num_y, d), got {tuple(x.shape)} and {tuple(y.shape)}")
    bs, nx, d = x.shape
    ny = y.shape[1]
    out = torch.empty((bs, nx, ny), dtype=torch.float32, device=dev)
    if nx == 1:  # the shape NWHead uses: one query per batch element against its own support
        check(lib.nw_direct_scores(k, scale, ptr(x), bs, d, ptr(y), ny, 1, ptr(out), st), "nw_direct_scores")
    else:
        for i in range(bs):
            check(lib.nw_direct_scores(k, scale, ptr(x[i]), nx, d, ptr(y[i]), ny, 0, ptr(out[i]), st),
                  "nw_direct_scores")
    return out


class _Kernel(nn.Module):
    kind = None

    def scale_value(self) -> float:
        return 1.0

    def forward(self, x, y):
        return dense_scores(self.kind, x, y, self.scale_value())


class EuclideanDistance(_Kernel):       # reference nwhead/kernel.py:13-15
    kind = "euclidean"


class HypersphereEuclideanDistance(_Kernel):  # reference nwhead/kernel.py:17-21
    kind = "hypersphere_euclidean"


class CosineDistance(_Kernel):          # reference nwhead/kernel.py:23-28
    kind = "cosine"


class DotProduct(_Kernel):              # reference nwhead/kernel.py:30-33
    kind = "dotproduct"


class Clip(_Kernel):                    # reference nwhead/kernel.py:35-44
    kind = "clip"

    def __init__(self):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def scale_value(self) -> float:
        return float(self.logit_scale.detach().exp())


def get_kernel(kernel_type):
    """reference nwhead/kernel.py:80-97; unknown names raise NotImplementedError."""
    table = {
        "euclidean": EuclideanDistance,
        "hypersphere_euclidean": HypersphereEuclideanDistance,
        "cosine": CosineDistance,
        "dotproduct": DotProduct,
        "clip": Clip,
    }
    if kernel_type not in table:
        raise NotImplementedError
    return table[kernel_type]()

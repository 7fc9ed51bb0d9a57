"""support_influence on the GPU (API of the reference's util/metric.py:23-50)."""
import torch

from . import _abi
from ._abi import check, load, ptr, stream_of


def onehot_argmax(onehot: torch.Tensor) -> torch.Tensor:
    """argmax over the last axis of a one-hot (or any) fp32 matrix -> int32, first maximum wins."""
    lib = load()
    flat = onehot.detach().float().contiguous().view(-1, onehot.shape[-1])
    out = torch.empty((flat.shape[0],), dtype=torch.int32, device=flat.device)
    check(lib.nw_onehot_argmax(ptr(flat), flat.shape[0], flat.shape[1], ptr(out), stream_of(flat.device)),
          "nw_onehot_argmax")
    return out.view(onehot.shape[:-1])


def support_influence(softmaxes, qlabels, sweights, slabels):
    '''
    Influence is defined as L(rescaled_softmax, qlabel) - L(softmax, qlabel).
    Positive influence => removing support image increases loss => support image was helpful
    Negative influence => removing support image decreases loss => support image was harmful

    softmaxes: (bs, num_classes)
    qlabels: One-hot encoded query label (bs, num_classes)
    sweights: Weights between query and each support (bs, num_support)
    slabels: One-hot encoded support label (num_support, num_classes) -> result (bs, num_support);
             the documented (bs, num_support, num_classes) form reproduces the reference's broadcast
             (util/metric.py:43 takes the argmax over the whole tensor) -> result (bs, bs, num_support).

    One batched elementwise kernel (nw_support_influence) instead of the per-query Python loop.
    '''
    dev = _abi.require_cuda(softmaxes, qlabels, sweights, slabels)
    lib = load()
    P = softmaxes.detach().float().contiguous()
    w = sweights.detach().float().contiguous()
    b, c = P.shape
    n = w.shape[1]
    qy = onehot_argmax(qlabels)
    sy = onehot_argmax(slabels)
    sets = 1 if sy.dim() == 1 else sy.shape[0]
    out = torch.empty((b, sets, n), dtype=torch.float32, device=dev)
    check(lib.nw_support_influence(ptr(P), ptr(qy), ptr(w), ptr(sy.contiguous()), b, sets, n, c, ptr(out),
                                   stream_of(dev)), "nw_support_influence")
    return out[:, 0, :] if sy.dim() == 1 else out


def support_influence_from_labels(softmaxes, qlabel_idx, sweights, slabel_idx):
    """Same computation from categorical int labels (skips the one-hot argmax pass)."""
    dev = _abi.require_cuda(softmaxes, qlabel_idx, sweights, slabel_idx)
    lib = load()
    P = softmaxes.detach().float().contiguous()
    w = sweights.detach().float().contiguous()
    qy = qlabel_idx.to(torch.int32).contiguous()
    sy = slabel_idx.to(torch.int32).contiguous()
    b, c = P.shape
    n = w.shape[1]
    out = torch.empty((b, 1, n), dtype=torch.float32, device=dev)
    check(lib.nw_support_influence(ptr(P), ptr(qy), ptr(w), ptr(sy), b, 1, n, c, ptr(out), stream_of(dev)),
          "nw_support_influence")
    return out[:, 0, :]


def support_influence_from_features(qfeat, bank, qlabel_idx, scale: float = 1.0, source_order: bool = True):
    """Influence of every bank row on every query computed directly from the query features and a SupportBank
    (two tensor-core passes, SupportBank.support_influence): no (bs, num_support) weight matrix is needed."""
    return bank.support_influence(qfeat, qlabel_idx, scale, source_order)

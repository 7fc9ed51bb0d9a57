"""`util.metric` of the reference (util/metric.py) for unmodified callers (`from util import metric`,
`from util.metric import Metric, ECELoss`, reference train.py:15-17, 373, 398, 416).

`support_influence` (util/metric.py:23-50) is the hot-path function: it is the CUDA kernel of nwhead_b200.  The
other names are the reference's host-side bookkeeping on (batch,) / (batch, C) tensors — out of the hot path
 — restated compactly here so that the shim is a complete module: same names, arguments and
results; no per-bin host synchronisation in ECELoss.
"""
import numpy as np
import torch
from torch.nn.modules.loss import _WeightedLoss

from nwhead_b200.metric import (support_influence, support_influence_from_features,  # noqa: F401
                                support_influence_from_labels)


def check_type(x):
    return x.cpu().detach().numpy() if isinstance(x, torch.Tensor) else x


def acc(pred, targets):
    '''Accuracy of a batch of categorical predictions (util/metric.py:10-14: sklearn accuracy_score).'''
    pred, targets = np.asarray(check_type(pred)), np.asarray(check_type(targets))
    return float(np.mean(pred == targets))


def roc(pr, gt):
    '''100 x ROC-AUC of scores `pr` against the binary mask `gt` (util/metric.py:16-21).'''
    from sklearn.metrics import roc_auc_score

    return 100 * float(roc_auc_score(check_type(gt), check_type(pr)))


class Metric:
    '''Sample-weighted running mean (util/metric.py:52-72).'''

    def __init__(self) -> None:
        self.reset_state()

    def update_state(self, val, samples):
        if isinstance(val, torch.Tensor):
            val = val.cpu().detach().item()
        if isinstance(val, np.ndarray):
            val = val.item()
        self.num_samples += samples
        self.tot_val += val * samples

    def result(self):
        return self.tot_val / self.num_samples if self.num_samples else 0

    def reset_state(self):
        self.tot_val = 0
        self.num_samples = 0


class ECELoss(torch.nn.Module):
    '''Expected calibration error over `n_bins` equal-width confidence bins (lower, upper]
    (util/metric.py:75-112).  Input: probabilities (N, C) and labels (N,); returns a (1,) tensor.
    One bucketize + three scatter-adds instead of a Python loop with two host syncs per bin.'''

    def __init__(self, n_bins=15):
        super().__init__()
        self.n_bins = n_bins
        self.register_buffer("uppers", torch.linspace(0, 1, n_bins + 1)[1:], persistent=False)

    def forward(self, softmaxes, labels):
        conf, pred = torch.max(softmaxes, dim=1)
        correct = pred.eq(labels).to(conf.dtype)
        uppers = self.uppers.to(conf.device)
        bins = torch.bucketize(conf, uppers, right=False)  # first upper edge >= conf: conf in (lower, upper]
        keep = (conf > 0) & (bins < self.n_bins)
        bins = bins.clamp_max(self.n_bins - 1)
        w = keep.to(conf.dtype)
        zeros = torch.zeros(self.n_bins, dtype=conf.dtype, device=conf.device)
        count = zeros.index_add(0, bins, w)
        conf_sum = zeros.index_add(0, bins, conf * w)
        acc_sum = zeros.index_add(0, bins, correct * w)
        n = max(conf.numel(), 1)
        safe = count.clamp_min(1)
        gap = (conf_sum / safe - acc_sum / safe).abs() * (count / n)
        return gap.sum().reshape(1)


class SmoothNLLLoss(_WeightedLoss):
    '''NLL loss on log-probabilities with label smoothing (util/metric.py:114-142).'''

    def __init__(self, weight=None, reduction='mean', smoothing=0.0):
        super().__init__(weight=weight, reduction=reduction)
        self.smoothing = smoothing
        self.weight = weight
        self.reduction = reduction

    def forward(self, log_preds, targets):
        assert 0 <= self.smoothing < 1
        c = log_preds.size(-1)
        with torch.no_grad():
            soft = torch.full_like(log_preds, self.smoothing / (c - 1))
            soft.scatter_(1, targets.unsqueeze(1), 1.0 - self.smoothing)
        if self.weight is not None:
            log_preds = log_preds * self.weight.unsqueeze(0)
        loss = -(soft * log_preds).sum(dim=-1)
        if self.reduction == 'mean':
            return loss.mean()
        return loss.sum() if self.reduction == 'sum' else loss

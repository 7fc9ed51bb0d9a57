"""CPU: the drop-in import path.  The reference's callers write `from nwhead.nw import NWNet`,
`from util import metric`, `from util.metric import Metric, ECELoss` (reference train.py:15-18, README.md:39-94);
the shim packages at the repository root must serve exactly the product's objects under those names, and the
host-side metric helpers of `util.metric` must agree with the reference's."""
import numpy as np
import pytest
import torch

from oracle.ref_import import load_reference, reference_available


def test_reference_import_lines_resolve_to_the_product():
    import nwhead_b200
    from nwhead.kernel import get_kernel
    from nwhead.nw import NWHead, NWNet
    from nwhead.support import SupportSetEval, SupportSetTrain
    from nwhead.utils import FullDataset, InfiniteUniformClassLoader, compute_clusters
    from util import metric
    from util.metric import ECELoss, Metric  # noqa: F401

    assert NWNet is nwhead_b200.NWNet and NWHead is nwhead_b200.NWHead
    assert get_kernel is nwhead_b200.get_kernel
    assert metric.support_influence is nwhead_b200.support_influence
    assert compute_clusters is nwhead_b200.compute_clusters
    assert SupportSetEval.__module__ == "nwhead_b200.support" and SupportSetTrain.__module__ == "nwhead_b200.support"
    assert FullDataset.__module__ == "nwhead_b200.utils" and InfiniteUniformClassLoader.__module__ == "nwhead_b200.utils"
    with pytest.raises(NotImplementedError):
        get_kernel("bogus")  # reference nwhead/kernel.py:96


def test_host_metrics_basic():
    from util import metric

    assert metric.acc(torch.tensor([1, 2, 3, 3]), torch.tensor([1, 2, 0, 3])) == 0.75
    m = metric.Metric()
    assert m.result() == 0
    m.update_state(torch.tensor(2.0), 3)
    m.update_state(np.array(4.0), 1)
    assert m.result() == pytest.approx(2.5)
    m.reset_state()
    assert m.num_samples == 0


@pytest.mark.skipif(not reference_available(), reason="reference not mounted (GPU box)")
def test_host_metrics_match_reference():
    from util import metric

    ref = load_reference().metric
    g = torch.Generator().manual_seed(0)
    for n, c in [(1, 3), (257, 10), (1000, 200)]:
        probs = torch.softmax(torch.randn(n, c, generator=g) * 3, dim=1)
        probs[0, 0] = 1.0
        probs[0, 1:] = 0.0
        labels = torch.randint(0, c, (n,), generator=g)
        ours, theirs = metric.ECELoss()(probs, labels), ref.ECELoss()(probs, labels)
        assert ours.shape == theirs.shape
        assert torch.allclose(ours, theirs, atol=1e-6)
        assert metric.acc(probs.argmax(1), labels) == pytest.approx(ref.acc(probs.argmax(1), labels))
        logp = probs.clamp_min(1e-12).log()
        for sm in (0.0, 0.1):
            assert torch.allclose(metric.SmoothNLLLoss(smoothing=sm)(logp, labels),
                                  ref.SmoothNLLLoss(smoothing=sm)(logp, labels), atol=1e-5, rtol=1e-5)
    score, gt = torch.rand(50, generator=g), torch.arange(50) % 2
    assert metric.roc(score, gt) == pytest.approx(ref.roc(score, gt))

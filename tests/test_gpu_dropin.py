"""GPU: a reference-style caller, written against the REFERENCE's import path and call sequence, runs unmodified
on the product through the shim packages (`nwhead/`, `util/`).

The flow is the one of the reference's train.py (NWNet construction :227-235, nw_step :401-422, the epoch order
precompute -> eval random/full/cluster -> train :290-303, ECE :373) and of README.md:39-94; the data is a small
separable synthetic image set, so a few SGD steps must reduce the loss and every inference mode must classify it."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Blobs(torch.utils.data.Dataset):
    """(image, label) pairs with a `.targets` list and `.num_classes`, like the reference's datasets (data/bird.py)."""

    def __init__(self, n, num_classes, seed):
        g = torch.Generator().manual_seed(seed)
        self.num_classes = num_classes
        self.targets = [i % num_classes for i in range(n)]
        proto = torch.randn(num_classes, 3, 8, 8, generator=torch.Generator().manual_seed(99))
        self.x = proto[self.targets] + 0.3 * torch.randn(n, 3, 8, 8, generator=g)

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        return self.x[i], self.targets[i]


def test_train_py_style_caller_runs_unmodified(cuda_lib):
    # ---- the reference's import lines (train.py:15-18)
    from nwhead.nw import NWNet
    from util import metric
    from util.metric import ECELoss, Metric

    device = "cuda:0"
    torch.manual_seed(0)
    np.random.seed(0)
    train_dataset, val_dataset = Blobs(360, 12, 1), Blobs(48, 12, 2)
    train_loader = torch.utils.data.DataLoader(train_dataset, batch_size=8, shuffle=True)
    val_loader = torch.utils.data.DataLoader(val_dataset, batch_size=8, shuffle=False)
    num_classes = train_dataset.num_classes
    featurizer = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(192, 32), torch.nn.ReLU())
    feat_dim = 32
    # ---- train.py:227-235
    network = NWNet(featurizer,
                    num_classes,
                    support_dataset=train_dataset,
                    feat_dim=feat_dim,
                    proj_dim=0,
                    kernel_type='euclidean',
                    n_shot=1,
                    n_way=10,
                    debug_mode=False)
    network.to(device)
    criterion = torch.nn.NLLLoss()
    optimizer = torch.optim.SGD(network.parameters(), lr=1e-2, momentum=0.9, nesterov=True)

    def nw_step(batch, is_train=True, mode='random'):  # the call sequence of train.py:401-422
        img, label = batch
        img = img.float().to(device)
        label = label.to(device)
        optimizer.zero_grad()
        with torch.set_grad_enabled(is_train):
            output = network(img, label) if is_train else network.predict(img, mode)
            loss = criterion(output, label)
            if is_train:
                loss.backward()
                optimizer.step()
            acc = metric.acc(output.argmax(-1), label)
        return {'loss': loss.cpu().detach().numpy(), 'acc': acc * 100, 'batch_size': len(img),
                'prob': output.exp(), 'gt': label}

    def eval_epoch(mode):
        network.eval()
        m_loss, m_acc, probs, gts = Metric(), Metric(), [], []
        for batch in val_loader:
            res = nw_step(batch, is_train=False, mode=mode)
            m_loss.update_state(res['loss'], res['batch_size'])
            m_acc.update_state(res['acc'], res['batch_size'])
            probs.append(res['prob'])
            gts.append(res['gt'])
        ece = (ECELoss()(torch.cat(probs, dim=0), torch.cat(gts, dim=0)) * 100).item()
        return m_loss.result(), m_acc.result(), ece

    # epoch order of train.py:289-303: precompute + eval in every mode, then train
    network.eval()
    network.precompute()
    before = {mode: eval_epoch(mode) for mode in ('random', 'full', 'cluster')}
    network.train()
    losses = []
    for epoch in range(3):
        for batch in train_loader:
            losses.append(float(nw_step(batch, is_train=True)['loss']))
    network.eval()
    network.precompute()
    after = {mode: eval_epoch(mode) for mode in ('random', 'full', 'cluster')}

    assert np.isfinite(losses).all()
    assert np.mean(losses[-10:]) < np.mean(losses[:10])           # SGD through the CUDA backward learns
    for mode in ('random', 'full', 'cluster'):
        loss, acc, ece = after[mode]
        assert np.isfinite([loss, acc, ece]).all()
        assert acc >= 95.0, (mode, before[mode], after[mode])     # separable blobs
        assert loss <= before[mode][0] + 1e-3
    # README.md:80-94 evaluation snippet: predict under set_grad_enabled(False) returns (batch, classes) log-probs
    img, label = next(iter(val_loader))
    with torch.set_grad_enabled(False):
        output = network.predict(img.float().to(device), 'full')
    assert output.shape == (8, num_classes) and output.dtype == torch.float32 and output.device.type == 'cuda'
    assert not output.requires_grad
    assert torch.allclose(output.exp().sum(1), torch.ones(8, device=device), atol=1e-4)


def test_predict_is_differentiable_like_the_reference(cuda_lib):
    """reference NWNet.predict (nwhead/nw.py:127-160) is plain autograd code: gradients flow into the featurizer.
    With grad enabled the product routes predict through the differentiable direct path on the raw support rows."""
    from nwhead.nw import NWNet

    device = "cuda:0"
    torch.manual_seed(1)
    np.random.seed(1)
    ds = Blobs(120, 6, 3)
    featurizer = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(192, 16), torch.nn.ReLU())
    network = NWNet(featurizer, 6, support_dataset=ds, feat_dim=16, n_shot=1, n_way=6).to(device)
    network.eval()
    network.precompute()
    img = ds.x[:8].to(device)
    label = torch.tensor(ds.targets[:8], device=device)
    with torch.no_grad():
        ref = {m: network.predict(img, m) for m in ('full', 'cluster')}
    for mode in ('full', 'cluster', 'random', 'ensemble'):
        network.zero_grad()
        out = network.predict(img, mode)
        assert out.requires_grad
        torch.nn.functional.nll_loss(out, label).backward()
        g = featurizer[1].weight.grad
        assert g is not None and torch.isfinite(g).all() and g.abs().max() > 0
        if mode in ref:  # same numbers as the inference path (bf16x3 bank vs exact fp32 differences)
            assert torch.allclose(out.detach().exp(), ref[mode].exp(), atol=1e-3)


def test_half_features_and_inference_mode(cuda_lib):
    """autocast features (fp16/bf16) are upcast like torch.cdist does; inference-mode tensors (no
    version counter) must not crash the bank cache."""
    import nwhead_b200

    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    C, n, d = 5, 200, 64
    sx = torch.randn(n, d, generator=g, device=dev)
    sy = torch.arange(n, device=dev) % C
    q = torch.randn(16, d, generator=g, device=dev)
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), C)
    with torch.no_grad():
        base = head(q, sx, sy)
        half = head(q.half(), sx.half(), sy)
        assert half.dtype == torch.float32
        assert torch.allclose(half.exp(), head(q.half().float(), sx.half().float(), sy).exp(), atol=1e-5)
    with torch.inference_mode():
        out = head(q.clone(), sx.clone(), sy.clone())
        assert torch.allclose(out.exp(), base.exp(), atol=1e-5)
        bank = nwhead_b200.SupportBank.build(sx.clone(), sy.clone(), C, "euclidean")
        idx = bank.topk_exact(q.clone(), 3, sx.clone())
        assert idx.shape == (16, 3)


def test_large_class_count_backward_and_early_limit(cuda_lib):
    """d + C above the old 40 KB shared-memory limit works (opt-in dynamic shared memory), and a shape
    beyond the new limit raises BEFORE the forward."""
    import nwhead_b200

    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    C, d = 12000, 256
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), C)
    q = torch.randn(4, d, generator=g, device=dev, requires_grad=True)
    sx = torch.randn(12, d, generator=g, device=dev, requires_grad=True)
    sy = torch.randint(0, C, (12,), generator=g, device=dev)
    out = head(q, sx, sy)
    out.gather(1, sy[:4, None]).sum().backward()
    assert torch.isfinite(q.grad).all() and torch.isfinite(sx.grad).all() and q.grad.abs().max() > 0
    # reference autograd on the same inputs (torch ops on the GPU are the checker here, not the product)
    q2, s2 = q.detach().clone().requires_grad_(True), sx.detach().clone().requires_grad_(True)
    probs = torch.softmax(-torch.cdist(q2[:, None], s2[None].expand(4, -1, -1)), -1)
    ref = torch.log(torch.bmm(probs, torch.nn.functional.one_hot(sy, C).float()[None].expand(4, -1, -1)).squeeze(1) + 1e-12)
    ref.gather(1, sy[:4, None]).sum().backward()
    assert torch.allclose(q.grad, q2.grad, atol=2e-5, rtol=2e-4) and torch.allclose(sx.grad, s2.grad, atol=2e-5, rtol=2e-4)
    big = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 60000)
    with pytest.raises(NotImplementedError, match="feat_dim \\+ n_classes"):
        big(q, sx, sy)


def test_bad_label_forward_backward_is_memory_safe_and_raises(cuda_lib):
    """an out-of-range support label on the differentiable path must never index out of bounds in the
    backward, and must raise (F.one_hot's message, reference nwhead/nw.py:276): at the latest on the next forward or
    backward call once the flagged forward has finished, or right away through the blocking check."""
    import nwhead_b200
    from nwhead_b200 import nw as nwmod

    dev = torch.device("cuda:0")
    C, d = 10, 32
    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), C)
    g = torch.Generator(device=dev).manual_seed(0)
    guard = nwmod.label_guard(dev)
    msg = "Class values must be smaller than num_classes"
    for bad in (C, -100, 2 ** 40 + 3, 10 ** 6):
        q = torch.randn(8, d, generator=g, device=dev, requires_grad=True)
        sx = torch.randn(10, d, generator=g, device=dev, requires_grad=True)
        sy = torch.arange(10, device=dev) % C
        sy[3] = bad
        # (1) memory safety of the kernels themselves: run forward AND backward with the host-side poll disabled,
        #     as happens when the backward is issued before the forward has finished
        real_check, guard.check = guard.check, lambda block=False: None
        try:
            out = head(q, sx, sy)
            out.sum().backward()
            torch.cuda.synchronize()  # no illegal address / sticky context error
        finally:
            guard.check = real_check
        assert torch.isfinite(out).all() and torch.isfinite(q.grad).all() and torch.isfinite(sx.grad).all()
        # the bad row contributes to no class: the result equals the forward without its label mass
        with pytest.raises(RuntimeError, match=msg):
            head.check_labels(dev)
        # the flag was consumed: valid calls work again
        sy[3] = 0
        head(q, sx, sy).sum().backward()
        head.check_labels(dev)
    # (2) non-blocking poll: the error surfaces on a later call without any explicit check ...
    sy[5] = C + 7
    head(q, sx, sy)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match=msg):
        head(q, sx, sy.clamp_max(C - 1))
    # ... and in backward when the flagged forward has finished by then
    out = head(q, sx, sy)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match=msg):
        out.sum().backward()
    head.check_labels(dev)

"""GPU: the unmodified NWNet flow (precompute -> predict in three modes -> get_neighbors -> one training
step) with the CUDA head, against the reference run recorded in tests/golden/nwnet_flow.npz."""
import numpy as np
import pytest
import torch

from oracle import nw_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class TinyDataset(torch.utils.data.Dataset):
    def __init__(self, x, y):
        self.x, self.targets = torch.from_numpy(x), [int(v) for v in y]

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        return self.x[i], self.targets[i]


def make_net(g, kind, n_shot_cluster=1):
    import nwhead_b200

    feat = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(192, 16), torch.nn.ReLU())
    with torch.no_grad():
        feat[1].weight.copy_(torch.from_numpy(g["W"]))
        feat[1].bias.copy_(torch.from_numpy(g["b"]))
    ds = TinyDataset(g["ds_x"], g["ds_y"])
    net = nwhead_b200.NWNet(feat, 6, support_dataset=ds, feat_dim=16, kernel_type=kind, n_shot=2, n_way=4,
                            n_shot_random=2, n_shot_full=5, n_shot_cluster=n_shot_cluster, n_neighbors=3, device=DEV)
    return net.to(DEV), feat


@pytest.mark.parametrize("kind", ["euclidean", "cosine"])
def test_precompute_predict_neighbors(cuda_lib, golden_flow, kind):
    g = golden_flow
    net, _ = make_net(g, kind)
    net.eval()
    xq = torch.from_numpy(g["xq"]).to(DEV)
    with torch.no_grad():
        with pytest.raises(AttributeError, match="precompute"):
            net.predict(xq, mode="full")
        net.precompute()
        assert np.array_equal(net.full_y.cpu().numpy(), g[f"{kind}/full_y"])
        assert np.abs(net.full_feat.cpu().numpy() - g[f"{kind}/full_feat"]).max() < 1e-4
        assert np.array_equal(net.support_eval.cluster_y.cpu().numpy(), g[f"{kind}/cluster_y"])
        assert np.abs(net.support_eval.cluster_feat.cpu().numpy() - g[f"{kind}/cluster_feat"]).max() < 1e-4
        for mode in ("full", "cluster"):
            out = net.predict(xq, mode=mode).cpu().numpy()
            ref = g[f"{kind}/pred_{mode}"]
            assert np.abs(np.exp(out) - np.exp(ref)).max() < 1e-3
            assert (out.argmax(1) == ref.argmax(1)).all()
        np.random.seed(123)  # same numpy stream as the reference run -> same sampled support
        out = net.predict(xq, mode="random").cpu().numpy()
        assert np.abs(np.exp(out) - np.exp(g[f"{kind}/pred_random"])).max() < 1e-3
        # knn mode: k nearest bank rows of every query, concatenated into one shared support
        out = net.predict(xq, mode="knn").cpu().numpy()
        assert np.abs(np.exp(out) - np.exp(g[f"{kind}/pred_knn"])).max() < 1e-3
        with pytest.raises(NotImplementedError):
            net.predict(xq, mode="bogus")
        # 'hnsw' is served by the exact search: identical to knn mode here
        assert np.array_equal(net.predict(xq, mode="hnsw").cpu().numpy(), out)
        # a single environment: the ensemble of one equals full mode (reference nwhead/nw.py:143-154)
        ens = net.predict(xq, mode="ensemble").cpu().numpy()
        assert np.abs(np.exp(ens) - np.exp(g[f"{kind}/pred_full"])).max() < 1e-3
        if kind == "euclidean":
            nb = net.get_neighbors(xq).cpu().numpy()
            assert nb.shape == g[f"{kind}/neighbors"].shape and nb.dtype == np.int64
            assert np.array_equal(nb, g[f"{kind}/neighbors"])
            # the large-bank route (tensor-core block search + exact re-rank) gives the same neighbours and
            # the same knn-mode prediction; the thresholds are lowered so that this small bank takes it
            k = 5
            dense_nb = net.get_neighbors(xq, k).cpu().numpy()
            net.TOPK_EXACT_MIN_ROWS = 1
            assert np.array_equal(net.get_neighbors(xq, k).cpu().numpy(), dense_nb)
            assert sum(net.support_eval.full_bank.last_topk_path.values()) == len(xq)
            net.support_eval.knn.BANK_SEARCH_MIN_ROWS = 1
            assert np.array_equal(net.predict(xq, mode="knn").cpu().numpy(), out)


def test_cluster_mode_with_two_clusters_per_class(cuda_lib, golden_flow):
    """n_shot_cluster=2 (reference: one scikit-learn KMeans(2) per class, nwhead/support.py:118-123): two centroids
    per class from the GPU k-means, equal to the oracle's run on the same features; predict('cluster') is the head
    over them."""
    import nwhead_b200

    g = golden_flow
    net, _ = make_net(g, "euclidean", n_shot_cluster=2)
    net.eval()
    xq = torch.from_numpy(g["xq"]).to(DEV)
    with torch.no_grad():
        net.precompute()
        se = net.support_eval
        classes = np.unique(net.full_y.cpu().numpy())
        assert np.array_equal(se.cluster_y.cpu().numpy(), np.repeat(classes, 2))
        oc, _ = O.kmeans_centroids(net.full_feat.cpu().numpy(), net.full_y.cpu().numpy(), 2)
        assert np.abs(se.cluster_feat.cpu().numpy() - oc).max() < 1e-5
        out = net.predict(xq, mode="cluster")
        qf = net.featurizer(xq)
        ref = O.nw_forward(qf.cpu().numpy(), oc, np.repeat(classes, 2), 6, "euclidean")
        assert np.abs(np.exp(out.cpu().numpy()) - np.exp(ref)).max() < 1e-3


@pytest.mark.parametrize("kind", ["euclidean", "cosine"])
def test_training_step(cuda_lib, golden_flow, kind):
    g = golden_flow
    net, feat = make_net(g, kind)
    net.train()
    np.random.seed(321)
    x, y = torch.from_numpy(g["xq"][:4]).to(DEV), torch.from_numpy(g["yq"][:4]).to(DEV)
    logp = net(x, y)
    loss = torch.nn.functional.nll_loss(logp, y)
    loss.backward()
    assert np.abs(logp.detach().cpu().numpy() - g[f"{kind}/train_logp"]).max() < 1e-4
    gw = g[f"{kind}/train_gW"]
    assert np.abs(feat[1].weight.grad.cpu().numpy() - gw).max() < 1e-6 + 2e-4 * np.abs(gw).max()


def test_return_mask_and_functional_support(cuda_lib, golden_flow):
    import nwhead_b200

    g = golden_flow
    net, feat = make_net(g, "euclidean")
    net.return_mask = True
    net.train()
    x, y = torch.from_numpy(g["xq"][:4]).to(DEV), torch.from_numpy(g["yq"][:4]).to(DEV)
    sx = torch.from_numpy(g["ds_x"][:12]).to(DEV)
    sy = torch.from_numpy(g["ds_y"][:12]).to(DEV)
    logp, mask = net(x, y, support_data=(sx, sy, None))
    assert logp.shape == (4, 6) and mask.dtype == torch.bool
    assert torch.equal(mask, torch.isin(y, sy))
    logp.sum().backward()
    assert feat[1].weight.grad is not None and torch.isfinite(feat[1].weight.grad).all()
    # the same support through the plain head gives the same numbers
    with torch.no_grad():
        f = feat(torch.cat((x, sx)))
        ref = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 6)(f[:4], f[4:], sy)
    assert torch.allclose(ref, logp.detach(), atol=1e-6)
    net.eval()
    with torch.no_grad():
        net.precompute()
        out, m = net.predict(x, mode="full")
    assert m.all() and out.shape == (4, 6)


def test_environments_ensemble_and_irm(cuda_lib, golden_flow, golden_env_flow):
    """env_array supports: per-environment banks, mode='ensemble' (mean of per-environment probabilities) and
    train_type='irm' sampling, against the reference run in tests/golden/nwnet_env_flow.npz."""
    import nwhead_b200

    g, e = golden_flow, golden_env_flow
    feat = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(192, 16), torch.nn.ReLU())
    with torch.no_grad():
        feat[1].weight.copy_(torch.from_numpy(g["W"]))
        feat[1].bias.copy_(torch.from_numpy(g["b"]))
    ds = TinyDataset(g["ds_x"], g["ds_y"])
    net = nwhead_b200.NWNet(feat, 6, support_dataset=ds, feat_dim=16, kernel_type="euclidean", n_shot=2, n_way=4,
                            n_shot_random=2, n_shot_full=5, n_shot_cluster=1, n_neighbors=3, env_array=e["env"],
                            device=DEV).to(DEV)
    net.eval()
    xq = torch.from_numpy(g["xq"]).to(DEV)
    with torch.no_grad():
        net.precompute()
        assert np.array_equal(net.full_y.cpu().numpy(), e["full_y"])
        assert [len(b) for b in net.support_eval.env_banks] == e["env_sizes"].tolist()
        for mode in ("full", "cluster", "ensemble"):
            out = net.predict(xq, mode=mode).cpu().numpy()
            assert np.abs(np.exp(out) - np.exp(e[f"pred_{mode}"])).max() < 1e-3, mode
    irm = nwhead_b200.NWNet(feat, 6, support_dataset=ds, feat_dim=16, kernel_type="euclidean", train_type="irm",
                            n_shot=2, env_array=e["env"], device=DEV).to(DEV)
    irm.train()
    np.random.seed(2024)
    logp = irm(xq[:4].clone(), torch.tensor([0, 1, 2, 3], device=DEV))
    assert np.abs(logp.detach().cpu().numpy() - e["irm_train_logp"]).max() < 1e-4
    logp.sum().backward()
    assert torch.isfinite(feat[1].weight.grad).all()

"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only) and
pin the float64 oracle + torch port against it.

    python oracle/gen_golden.py            # writes tests/golden/, asserts oracle == reference

The reference ships no tests or golden vectors; these fixtures are outputs of
the reference itself (torch 2.11 CPU, fp32) on small seeded inputs, committed so that the
GPU box (which has no /root/reference) can check both the oracle and the CUDA path.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import nw_oracle as O  # noqa: E402
from oracle import torch_port as TP  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
KINDS = O.KERNEL_KINDS


def relu_feats(g, n, d, n_classes=None, labels=None, spread=1.0):
    """ReLU-like non-negative features with a class-dependent mean."""
    x = torch.randn(n, d, generator=g) * spread
    if labels is not None:
        mu = torch.randn(n_classes, d, generator=g)
        x = x + mu[labels]
    return torch.relu(x + 0.5)


def head_cases(ref):
    g = torch.Generator().manual_seed(20260101)
    out = {}
    shapes = {
        # name: (B, N, d, C, threeD)
        "exact_small": (8, 10, 32, 12, False),   # N <= 25 -> torch.cdist exact-difference path
        "mm_medium": (5, 60, 48, 7, False),      # N > 25  -> torch.cdist mm path
        "per_query_3d": (4, 12, 16, 5, True),    # (b, N, d) support + (b, N) labels
        "wide": (3, 300, 64, 40, False),
    }
    for name, (B, N, d, C, threeD) in shapes.items():
        if threeD:
            y = torch.randint(0, C, (B, N), generator=g)
            s = torch.relu(torch.randn(B, N, d, generator=g) + 0.5)
        else:
            y = torch.randint(0, C - 2, (N,), generator=g)  # classes C-2, C-1 never present
            y[1] = y[0]                                     # duplicate labels
            s = relu_feats(g, N, d, C, y)
        q = torch.relu(torch.randn(B, d, generator=g) + 0.5)
        if not threeD:
            q[0] = s[3]  # coincident point: zero distance (SURVEY.md 7.2 "sqrt at zero distance")
        G = torch.randn(B, C, generator=g)
        out[f"{name}/q"], out[f"{name}/s"], out[f"{name}/y"], out[f"{name}/G"] = q, s, y, G
        out[f"{name}/C"] = torch.tensor(C)
        for kind in KINDS:
            kern = ref.get_kernel(kind)
            head = ref.NWHead(kern, C)
            qq = q.clone().requires_grad_(True)
            ss = s.clone().requires_grad_(True)
            logp = head(qq, ss, y)
            (logp * G).sum().backward()
            out[f"{name}/{kind}/logp"] = logp.detach()
            out[f"{name}/{kind}/gq"] = qq.grad
            out[f"{name}/{kind}/gs"] = ss.grad
            if kind == "clip":
                out[f"{name}/{kind}/glogit"] = kern.logit_scale.grad.detach()
            # ---- pin the oracle + the torch port
            o = O.nw_forward(q.numpy(), s.numpy(), y.numpy(), C, kind)
            err = np.abs(o - logp.detach().numpy()).max()
            assert err < 2e-4, (name, kind, "fwd", err)
            ls = torch.tensor(O.CLIP_LOGIT_SCALE_INIT)
            pt = TP.port_nw_forward(q, s, y, C, kind, ls)
            assert torch.equal(pt, logp.detach()), (name, kind, "port")
            if not threeD:
                res = O.nw_backward(q.numpy(), s.numpy(), y.numpy(), C, G.numpy(), kind)
                scale = max(1.0, float(qq.grad.abs().max()), float(ss.grad.abs().max()))
                e1 = np.abs(res[0] - qq.grad.numpy()).max() / scale
                e2 = np.abs(res[1] - ss.grad.numpy()).max() / scale
                assert e1 < 5e-4 and e2 < 5e-4, (name, kind, "bwd", e1, e2)
                if kind == "clip":
                    e3 = abs(res[2] - float(kern.logit_scale.grad)) / max(1.0, abs(res[2]))
                    assert e3 < 5e-4, (name, "clip glogit", e3)
    return out


def influence_cases(ref):
    g = torch.Generator().manual_seed(77)
    out = {}
    B, N, C, d = 6, 40, 5, 8
    sy = torch.randint(0, C, (N,), generator=g)
    s = relu_feats(g, N, d, C, sy)
    q = torch.relu(torch.randn(B, d, generator=g) + 0.5)
    qy = torch.randint(0, C, (B,), generator=g)
    w = torch.softmax(-torch.cdist(q, s), dim=-1)
    P = w @ torch.nn.functional.one_hot(sy, C).float()
    qoh = torch.nn.functional.one_hot(qy, C).float()
    soh = torch.nn.functional.one_hot(sy, C).float()
    infl = ref.support_influence(P, qoh, w, soh)                      # (B,N)
    soh3 = soh[None].expand(B, N, C).contiguous()
    infl3 = ref.support_influence(P, qoh, w, soh3)                    # (B,B,N) quirk
    out.update(P=P, qy=qy, w=w, sy=sy, infl=infl, infl3=infl3)
    o = O.support_influence(P.numpy(), qy.numpy(), w.numpy(), sy.numpy())
    assert np.allclose(o, infl.numpy(), rtol=2e-4, atol=2e-5), np.abs(o - infl.numpy()).max()
    sy3 = sy[None].expand(B, N).numpy()
    o3 = O.support_influence(P.numpy(), qy.numpy(), w.numpy(), sy3)
    assert o3.shape == tuple(infl3.shape)
    assert np.allclose(o3, infl3.numpy(), rtol=2e-4, atol=2e-5)
    assert torch.equal(TP.port_support_influence(P, qoh, w, soh), infl)
    # a class whose only support carries all its mass -> +inf must be preserved
    P1 = torch.tensor([[0.25, 0.75]]); w1 = torch.tensor([[0.25, 0.75]])
    i1 = ref.support_influence(P1, torch.tensor([[1.0, 0.0]]), w1, torch.tensor([[1.0, 0.0], [0.0, 1.0]]))
    out.update(edge_P=P1, edge_w=w1, edge_infl=i1)
    return out


def cluster_cases(ref):
    g = torch.Generator().manual_seed(5)
    N, d, C = 90, 8, 6
    y = torch.sort(torch.randint(0, C, (N,), generator=g)).values
    y[y == 4] = 5                                   # class 4 absent: unique labels skip it
    f = relu_feats(g, N, d, C, y)
    cf, cy = ref.compute_clusters(f, y, 1)
    oc, oy = O.class_centroids(f.numpy(), y.numpy())
    assert np.array_equal(oy, cy.numpy())
    assert np.abs(oc - cf.numpy()).max() < 2e-6
    pc, py = TP.port_class_centroids(f, y)
    assert torch.equal(py, cy) and (pc - cf).abs().max() < 2e-6
    out = dict(f=f, y=y, cf=cf, cy=cy)
    # n_clusters = 3: per-class blobs whose clustering is unambiguous, so that sklearn's KMeans (reference) and a
    # k-means++/Lloyd run with any other random stream reach the same fixed point.  Labels are shuffled.
    k, C3, d3, per = 3, 5, 16, 20
    centres = torch.randn(C3, k, d3, generator=g) * 6
    yk = torch.arange(C3).repeat_interleave(k * per)
    yk[yk == 2] = 6                                  # class ids with gaps: 0,1,3,4,6
    fk = (centres.reshape(C3 * k, 1, d3) + 0.3 * torch.randn(C3 * k, per, d3, generator=g)).reshape(-1, d3)
    shuffle = torch.randperm(len(yk), generator=g)
    fk, yk = fk[shuffle].contiguous(), yk[shuffle].contiguous()
    cfk, cyk = ref.compute_clusters(fk.numpy(), yk.numpy(), k)
    ok, oyk = O.kmeans_centroids(fk.numpy(), yk.numpy(), k)
    assert np.array_equal(oyk, cyk.numpy())
    assert np.abs(ok - cfk.numpy()).max() < 1e-5       # row for row: scikit-learn's seeding stream is restated
    # the head over those centroids (order within a class does not matter)
    qk = fk[:7] + 0.1
    head = ref.NWHead(ref.get_kernel("euclidean"), 7)
    pk = head(qk, cfk, cyk)
    po = O.nw_forward(qk.numpy(), ok, oyk, 7, "euclidean")
    assert np.abs(np.exp(po) - np.exp(pk.numpy())).max() < 1e-5
    # closest=True: nearest real embedding to every centroid
    ck, _ = ref.compute_clusters(fk, yk.numpy(), k, closest=True)
    out.update(k3_f=fk, k3_y=yk, k3_cf=cfk, k3_cy=cyk, k3_q=qk, k3_logp=pk, k3_closest=ck)
    # AMBIGUOUS clusters (heavily overlapping blobs, ReLU features): the centroids depend on the seeding, so this
    # pins the restated scikit-learn random stream (fresh RandomState(0) per class, k-means++ with local trials) and
    # stopping rule — not just the fixed point.  Several k, uneven class sizes, shuffled labels.
    for tag, ka, Ca, da in (("amb2", 2, 7, 12), ("amb4", 4, 6, 24)):
        sizes = torch.randint(60, 200, (Ca,), generator=g)
        ya = torch.cat([torch.full((int(sz),), c) for c, sz in enumerate(sizes)])
        blobs = torch.randn(Ca, 5, da, generator=g) * 0.7
        which = torch.randint(0, 5, (len(ya),), generator=g)
        fa = torch.relu(blobs[ya, which] + torch.randn(len(ya), da, generator=g) + 0.5)
        sh = torch.randperm(len(ya), generator=g)
        fa, ya = fa[sh].contiguous(), ya[sh].contiguous()
        cfa, cya = ref.compute_clusters(fa.numpy(), ya.numpy(), ka)
        oa, oya = O.kmeans_centroids(fa.numpy(), ya.numpy(), ka)
        assert np.array_equal(oya, cya.numpy())
        assert np.abs(oa - cfa.numpy()).max() < 1e-5, tag
        gap = O.kmeans_inertia(fa.numpy(), ya.numpy(), oa, ka) / O.kmeans_inertia(fa.numpy(), ya.numpy(), cfa.numpy(), ka)
        assert np.abs(gap - 1).max() < 1e-6
        out.update({f"{tag}_f": fa, f"{tag}_y": ya, f"{tag}_cf": cfa, f"{tag}_cy": cya})
    return out


class TinyDataset(torch.utils.data.Dataset):
    """Synthetic image dataset with uneven class counts and a .targets attribute."""

    def __init__(self, n, n_classes, seed):
        g = torch.Generator().manual_seed(seed)
        self.targets = [int(v) for v in torch.randint(0, n_classes, (n,), generator=g)]
        # make sure every class is present at least 3 times
        for c in range(n_classes):
            for r in range(3):
                self.targets[c * 3 + r] = c
        self.x = torch.randn(n, 3, 8, 8, generator=g) + torch.tensor(self.targets).view(-1, 1, 1, 1) * 0.3

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        return self.x[i], self.targets[i]


def tiny_featurizer(seed=3):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(192, 16), torch.nn.ReLU())


def nwnet_flow(ref):
    out = {}
    C = 6
    ds = TinyDataset(80, C, seed=11)
    feat = tiny_featurizer()
    out["W"] = feat[1].weight.detach().clone()
    out["b"] = feat[1].bias.detach().clone()
    out["ds_x"], out["ds_y"] = ds.x, torch.tensor(ds.targets)
    for kind in ("euclidean", "cosine"):
        net = ref.NWNet(feat, C, support_dataset=ds, feat_dim=16, kernel_type=kind, n_shot=2, n_way=4,
                        n_shot_random=2, n_shot_full=5, n_shot_cluster=1, n_neighbors=3, device="cpu",
                        return_mask=False)
        net.eval()
        with torch.no_grad():
            net.precompute()
            g = torch.Generator().manual_seed(99)
            xq = torch.randn(7, 3, 8, 8, generator=g)
            yq = torch.tensor([0, 1, 2, 3, 3, 1, 0])
            out["xq"], out["yq"] = xq, yq
            out[f"{kind}/full_feat"] = net.full_feat
            out[f"{kind}/full_y"] = net.full_y
            out[f"{kind}/cluster_feat"] = net.support_eval.cluster_feat
            out[f"{kind}/cluster_y"] = net.support_eval.cluster_y
            out[f"{kind}/pred_full"] = net.predict(xq, mode="full")
            out[f"{kind}/pred_cluster"] = net.predict(xq, mode="cluster")
            np.random.seed(123)
            out[f"{kind}/pred_random"] = net.predict(xq, mode="random")
            out[f"{kind}/pred_knn"] = net.predict(xq, mode="knn")   # k nearest of every query, one shared support
            if kind == "euclidean":
                out[f"{kind}/neighbors"] = net.get_neighbors(xq)
        net.train()
        np.random.seed(321)
        qq, yy = xq[:4].clone(), yq[:4]      # len(qy) <= n_way (nwhead/utils.py:124)
        logp = net(qq, yy)
        loss = torch.nn.functional.nll_loss(logp, yy)
        net.zero_grad()
        loss.backward()
        out[f"{kind}/train_logp"] = logp.detach()
        out[f"{kind}/train_gW"] = feat[1].weight.grad.detach().clone()
        # bank order oracle
        keys = O.full_bank_keys(ds.targets, 5)
        assert np.array_equal(np.asarray(ds.targets)[keys], net.full_y.numpy())
    return out


def nwnet_env_flow(ref):
    """Environment-split supports: mode='ensemble' inference and train_type='irm' sampling
    (nwhead/support.py:17-56, 74-93; nwhead/nw.py:143-154)."""
    out = {}
    C = 6
    ds = TinyDataset(80, C, seed=11)
    env = np.array([(i // 3) % 2 for i in range(80)])
    feat = tiny_featurizer()
    out["env"] = env
    net = ref.NWNet(feat, C, support_dataset=ds, feat_dim=16, kernel_type="euclidean", n_shot=2, n_way=4,
                    n_shot_random=2, n_shot_full=5, n_shot_cluster=1, n_neighbors=3, env_array=env, device="cpu")
    net.eval()
    g = torch.Generator().manual_seed(99)
    xq = torch.randn(7, 3, 8, 8, generator=g)
    with torch.no_grad():
        net.precompute()
        out["full_y"] = net.full_y
        out["env_sizes"] = torch.tensor([len(f) for f in net.support_eval.full_feat_sep])
        out["pred_full"] = net.predict(xq, mode="full")
        out["pred_cluster"] = net.predict(xq, mode="cluster")
        out["pred_ensemble"] = net.predict(xq, mode="ensemble")
    irm = ref.NWNet(feat, C, support_dataset=ds, feat_dim=16, kernel_type="euclidean", train_type="irm", n_shot=2,
                    env_array=env, device="cpu")
    irm.train()
    np.random.seed(2024)
    yq = torch.tensor([0, 1, 2, 3])
    logp = irm(xq[:4].clone(), yq)
    out["irm_train_logp"] = logp.detach()
    return out


def save(name, d):
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}: {len(arrs)} arrays, {os.path.getsize(path)/1024:.1f} KiB")


def main():
    torch.set_num_threads(1)
    torch.use_deterministic_algorithms(True)
    ref = load_reference()
    os.makedirs(GOLD, exist_ok=True)
    save("head", head_cases(ref))
    save("influence", influence_cases(ref))
    save("clusters", cluster_cases(ref))
    save("nwnet_flow", nwnet_flow(ref))
    save("nwnet_env_flow", nwnet_env_flow(ref))
    print("oracle and torch port agree with the reference on every generated case")


if __name__ == "__main__":
    main()

"""Import the UNMODIFIED reference from /root/reference (build container only).

The reference hard-imports ``hnswlib`` (nwhead/utils.py:4), which is not installed and is
only used by ``mode='hnsw'`` (out of scope).  A stub module is injected before the import so
the reference source stays untouched.  /root/reference does not exist on the GPU box:
only ``oracle/gen_golden.py`` and CPU-side tests that skip when it is absent use this file.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NW_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "nwhead", "nw.py"))


def _install_hnswlib_stub() -> None:
    if "hnswlib" in sys.modules:
        return
    stub = types.ModuleType("hnswlib")

    class Index:  # only what SupportSetEval.build_infer_iters touches (nwhead/utils.py:203-207)
        def __init__(self, space, dim):
            self.space, self.dim = space, dim

        def init_index(self, **kw):
            pass

        def add_items(self, data):
            pass

        def knn_query(self, x, k):
            raise NotImplementedError("hnsw mode is out of scope")

    stub.Index = Index
    sys.modules["hnswlib"] = stub


def load_reference():
    """Returns a namespace with the reference's NWNet, NWHead, get_kernel, support_influence,
    compute_clusters, FullDataset, InfiniteUniformClassLoader."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_hnswlib_stub()
    # The reference's top-level packages are called `nwhead` and `util`; import them under
    # their own names from the reference root.  Our package is `nwhead_b200`, so no clash.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    nw = importlib.import_module("nwhead.nw")
    kern = importlib.import_module("nwhead.kernel")
    utils = importlib.import_module("nwhead.utils")
    support = importlib.import_module("nwhead.support")
    metric = importlib.import_module("util.metric")
    ns = types.SimpleNamespace(
        NWNet=nw.NWNet,
        NWHead=nw.NWHead,
        get_kernel=kern.get_kernel,
        support_influence=metric.support_influence,
        compute_clusters=utils.compute_clusters,
        FullDataset=utils.FullDataset,
        InfiniteUniformClassLoader=utils.InfiniteUniformClassLoader,
        get_separated_indices=utils.get_separated_indices,
        SupportSetEval=support.SupportSetEval,
        SupportSetTrain=support.SupportSetTrain,
    )
    return ns

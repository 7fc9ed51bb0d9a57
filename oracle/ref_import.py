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
    # The reference's top-level packages are called `nwhead` and `util` — the same names as the drop-in shim
    # packages at this repository's root.  They are loaded here under PRIVATE package names (the reference only
    # uses relative imports inside `nwhead/`), so neither sys.path nor sys.modules['nwhead'] is touched.
    import importlib
    import importlib.util

    def load_pkg(alias, directory):
        if alias in sys.modules:
            return sys.modules[alias]
        pkg = types.ModuleType(alias)
        pkg.__path__ = [directory]
        pkg.__package__ = alias
        sys.modules[alias] = pkg
        return pkg

    def load_mod(alias_pkg, name, directory):
        full = f"{alias_pkg}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        spec = importlib.util.spec_from_file_location(full, os.path.join(directory, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        return mod

    nw_dir, util_dir = os.path.join(REFERENCE_ROOT, "nwhead"), os.path.join(REFERENCE_ROOT, "util")
    load_pkg("_reference_nwhead", nw_dir)
    load_pkg("_reference_util", util_dir)
    utils = load_mod("_reference_nwhead", "utils", nw_dir)
    kern = load_mod("_reference_nwhead", "kernel", nw_dir)
    support = load_mod("_reference_nwhead", "support", nw_dir)
    nw = load_mod("_reference_nwhead", "nw", nw_dir)
    metric = load_mod("_reference_util", "metric", util_dir)
    ns = types.SimpleNamespace(
        NWNet=nw.NWNet,
        NWHead=nw.NWHead,
        get_kernel=kern.get_kernel,
        support_influence=metric.support_influence,
        metric=metric,
        compute_clusters=utils.compute_clusters,
        FullDataset=utils.FullDataset,
        InfiniteUniformClassLoader=utils.InfiniteUniformClassLoader,
        get_separated_indices=utils.get_separated_indices,
        SupportSetEval=support.SupportSetEval,
        SupportSetTrain=support.SupportSetTrain,
    )
    return ns

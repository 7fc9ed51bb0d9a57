"""CPU: libnw_sm100.so builds, loads, and exports every symbol include/nw_sm100.h declares; the ctypes
table mirrors the header.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nw_sm100.h")).read()
    return re.findall(r"^NW_API\s+[\w\s\*]+?\b(nw_[a-z0-9_]+)\(", text, flags=re.M)


@pytest.fixture(scope="module")
def lib():
    from nwhead_b200 import build

    return ctypes.CDLL(build.build())


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    assert len(syms) == len(set(syms)) >= 20
    for must in ("nw_forward_class_lse", "nw_logp_from_class_lse", "nw_rows_to_bf16", "nw_direct_backward",
                 "nw_class_centroids", "nw_support_influence", "nw_rank_rows", "nw_last_error"):
        assert must in syms


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in nw_sm100.h but not exported"


def test_ctypes_table_matches_header():
    from nwhead_b200 import _abi

    assert sorted(_abi.SIGNATURES) == sorted(declared_symbols())


def test_pure_host_entry_points(lib):
    lib.nw_row_elems.restype = ctypes.c_int
    assert lib.nw_abi_version() == 3
    assert lib.nw_row_elems(2048, 1) == 2048
    assert lib.nw_row_elems(512, 3) == 1536
    assert lib.nw_row_elems(16, 1) == 64
    assert lib.nw_row_elems(100, 3) == 320
    assert lib.nw_row_elems(0, 1) < 0 and lib.nw_row_elems(8, 2) < 0
    lib.nw_direct_backward_workspace_elems.restype = ctypes.c_int64
    lib.nw_direct_backward_workspace_elems.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int]
    assert lib.nw_direct_backward_workspace_elems(8, 512, 10, 0) == 80 + 10 + 8
    assert lib.nw_direct_backward_workspace_elems(8, 512, 10, 1) == 80 + 80 + 8
    # large shared support: + the split grad_q partials, one (8, 512) slab per chunk of 1024 supports
    assert lib.nw_direct_backward_workspace_elems(8, 512, 5000, 0) == 8 * 5000 + 5000 + 8 + 5 * 8 * 512


def test_errors_are_reported_not_swallowed(lib):
    lib.nw_last_error.restype = ctypes.c_char_p
    # NULL pointers are rejected before any CUDA call is made
    rc = lib.nw_logp_from_class_lse(None, 4, 4, None, None)
    assert rc == -1 and b"NULL" in lib.nw_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nwhead_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{fn} references the oracle"


def test_no_cpu_fallback():
    import torch

    import nwhead_b200
    from nwhead_b200._abi import NWLibraryError

    head = nwhead_b200.NWHead(nwhead_b200.get_kernel("euclidean"), 3)
    with pytest.raises(NWLibraryError):
        head(torch.randn(2, 4), torch.randn(30, 4), torch.randint(0, 3, (30,)))
    with pytest.raises(NotImplementedError):
        nwhead_b200.get_kernel("relationnet")

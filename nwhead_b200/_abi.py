"""ctypes binding of libnw_sm100.so (include/nw_sm100.h).

PyTorch is used only for device memory and the current stream; every compute call goes through the
C ABI.  There is NO fallback: if the shared library is missing or the device is not sm_100 the
calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

NW_OK = 0
KIND = {"euclidean": 0, "hypersphere_euclidean": 1, "cosine": 2, "dotproduct": 3, "clip": 4}
EPI_EUCLID, EPI_LINEAR = 0, 1
PREC_BF16, PREC_BF16X3 = 1, 3
ROWS_BANK, ROWS_QUERY = 0, 1
EMIT_SCORES, EMIT_INFLUENCE, EMIT_BLOCK_BEST = 0, 1, 2
DIRECT_BACKWARD_MAX_D_PLUS_C = 49152  # NW_DIRECT_BACKWARD_MAX_D_PLUS_C

# NW_B200_LIB: developer override to A/B two builds of the library on the same GPU box
_LIB_PATH = os.environ.get("NW_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnw_sm100.so")


class NWLibraryError(RuntimeError):
    """A libnw_sm100 entry point returned an error status."""


class ForwardPlan(ctypes.Structure):
    _fields_ = [
        ("q_tiles", c_int),
        ("s_tiles", c_int),
        ("chunks", c_int),
        ("tiles_per_chunk", c_int),
        ("grid", c_int),
        ("cta_pair", c_int),
        ("side_elems", c_int64),
    ]


# name -> (restype, argtypes); mirrors include/nw_sm100.h one to one
SIGNATURES = {
    "nw_last_error": (c_char_p, []),
    "nw_abi_version": (c_int, []),
    "nw_device_check": (c_int, []),
    "nw_row_elems": (c_int, [c_int, c_int]),
    "nw_labels_to_i32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "nw_class_offsets": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "nw_column_mean_workspace_bytes": (c_size_t, [c_int]),
    "nw_column_mean": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nw_rows_to_bf16": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_void_p, c_int, c_void_p, c_void_p]),
    "nw_rows_to_bf16_peers": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int, c_int, POINTER(c_void_p),
                                      POINTER(c_void_p), c_int, c_int64, c_int64, c_int, c_void_p]),
    "nw_rounding_residual": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "nw_forward_plan": (c_int, [c_int, c_int64, POINTER(ForwardPlan)]),
    "nw_forward_set_clock_probe": (c_int, [c_void_p, c_int64]),
    "nw_forward_epilogue_sets": (c_int, [c_int]),
    "nw_forward_class_lse": (c_int, [c_int, c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "nw_forward_class_lse_peers": (c_int, [c_int, c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                           c_int64, c_int, c_int, POINTER(c_void_p), c_int, c_int, c_void_p,
                                           c_int64, c_void_p]),
    "nw_forward_emit": (c_int, [c_int, c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "nw_backward_coefficients_workspace_elems": (c_int64, [c_int64, c_int64]),
    "nw_backward_coefficients": (c_int, [c_int, c_float, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                         c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                         c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "nw_dense_products_transposed": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                             c_int, c_void_p, c_int64, c_void_p]),
    "nw_transpose_kblocks": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "nw_backward_finish": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p,
                                   c_int64, c_void_p]),
    "nw_dense_products": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_int64,
                                  c_void_p]),
    "nw_logp_from_class_lse": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "nw_row_stats": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nw_class_lse_merge": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "nw_direct_scores": (c_int, [c_int, c_float, c_void_p, c_int, c_int, c_void_p, c_int64, c_int, c_void_p,
                                 c_void_p]),
    "nw_direct_aggregate": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "nw_direct_forward": (c_int, [c_int, c_float, c_void_p, c_int, c_int, c_void_p, c_int64, c_int, c_void_p, c_int,
                                  c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nw_direct_backward_workspace_elems": (c_int64, [c_int, c_int, c_int64, c_int]),
    "nw_direct_backward": (c_int, [c_int, c_float, c_void_p, c_int, c_int, c_void_p, c_int64, c_int, c_void_p,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "nw_class_centroids_workspace_bytes": (c_size_t, [c_int, c_int]),
    "nw_class_centroids": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "nw_kmeans_assign": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p]),
    "nw_kmeans_seed_dist": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int,
                                    c_void_p, c_void_p]),
    "nw_kmeans_seed_pick": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "nw_onehot_argmax": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "nw_support_influence": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int,
                                     c_void_p, c_void_p]),
    "nw_rank_rows_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "nw_rank_rows": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nw_topk_refine": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
}

_lib = None


class _Stream(c_void_p):
    """A cudaStream_t for the ABI's void* stream argument that remembers which device it belongs to."""
    device = None


class _DeviceBoundLib:
    """The loaded library with every entry point wrapped so that it runs with the RIGHT CUDA device current.

    The C ABI takes plain pointers and a stream; kernels launch in the current-device context.  A caller that
    works on `cuda:1` while `cuda:0` is current (the reference does `.to(x.device)` and never calls set_device)
    would otherwise launch on the wrong GPU, or fail with cudaErrorInvalidResourceHandle for a non-default
    stream.  Every launching entry point takes the stream as its LAST argument, and `stream_of(device)` returns a
    stream handle tagged with its device: the wrapper switches to that device for the duration of the call when it
    is not already current (one integer compare on the usual path)."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name in SIGNATURES:
            setattr(self, name, self._bind(getattr(cdll, name)))

    @staticmethod
    def _bind(fn):
        def call(*args):
            dev = getattr(args[-1], "device", None) if args else None
            if dev is not None and dev != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    return fn(*args)
            return fn(*args)

        call.__name__ = fn.__name__
        call.argtypes, call.restype = fn.argtypes, fn.restype
        return call


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load libnw_sm100.so (built in-tree by nwhead_b200.build).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NWLibraryError(
            f"{_LIB_PATH} is missing: run `python -m nwhead_b200.build` (nvcc, sm_100a). "
            "nwhead_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.nw_abi_version() != 3:
        raise NWLibraryError("libnw_sm100.so ABI version mismatch")
    _lib = _DeviceBoundLib(lib)
    return _lib


def check(status: int, what: str) -> None:
    if status != NW_OK:
        msg = load().nw_last_error().decode("utf-8", "replace")
        raise NWLibraryError(f"{what} failed with status {status}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_of(device) -> c_void_p:
    """Current stream of `device` as the ABI's void*, tagged with the device index (see _DeviceBoundLib)."""
    device = torch.device(device)
    st = _Stream(torch.cuda.current_stream(device).cuda_stream)
    st.device = device.index if device.index is not None else torch.cuda.current_device()
    return st


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise NWLibraryError(
                "nwhead_b200 runs only on a CUDA (sm_100) device and has no CPU fallback; got a tensor on "
                f"{t.device}"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise NWLibraryError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def forward_plan(n_query: int, n_support: int) -> ForwardPlan:
    plan = ForwardPlan()
    check(load().nw_forward_plan(n_query, n_support, ctypes.byref(plan)), "nw_forward_plan")
    return plan

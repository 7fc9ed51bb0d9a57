"""GPU: the ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would add) is executed as
written — only the library path is pointed at the in-tree build — and checked against the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import nw_oracle as O
from gpu_util import assert_head_parity, clustered_features

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_stub_runs_against_the_library(cuda_lib):
    from nwhead_b200 import _abi

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# nwhead/_nw_sm100\.py.*?)```", text, re.S).group(1)
    assert 'ctypes.CDLL("libnw_sm100.so")' in code
    code = code.replace('ctypes.CDLL("libnw_sm100.so")', f"ctypes.CDLL({_abi.lib_path()!r})")
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    for C, per, d, B in ((7, 30, 96, 5), (12, 50, 256, 140)):
        q, s, y, _ = clustered_features(C, per, d, B, seed=C)
        bank = ns["build_bank"](torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), C)
        out = ns["nw_forward"](torch.from_numpy(q).to(DEV), bank)
        assert out.shape == (B, C)
        assert_head_parity(out, O.nw_forward(q, s, y, C, "euclidean"))
    with pytest.raises(RuntimeError):   # errors surface through nw_last_error, as the stub's _ok() expects
        ns["_ok"](ns["_lib"].nw_forward_plan(0, ctypes.c_int64(0), None))

"""Drop-in `nwhead` package: the import path of the reference (alanqrwang/nwhead) served by the B200 product.

The reference's callers write `from nwhead.nw import NWNet` (train.py:18, README.md:40) and `from util import
metric` (train.py:17).  These shim packages re-export `nwhead_b200` under those names so such code runs
UNMODIFIED: copy `nwhead/`, `util/metric.py` and `nwhead_b200/` (with its built libnw_sm100.so) over the
reference tree, or put this repository's root first on `sys.path`.  Nothing is implemented here.
"""
from nwhead_b200 import NWHead, NWNet, SupportBank, get_kernel  # noqa: F401

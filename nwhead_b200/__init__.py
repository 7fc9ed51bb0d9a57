"""nwhead_b200 — B200 (sm_100a) implementation of the Nadaraya-Watson head hot path, drop-in for the
Python API of alanqrwang/nwhead (NWHead / NWNet / get_kernel / support_influence).

Importing the package does not need a GPU; every compute call does, and raises without one.
"""
from . import _abi
from .bank import SupportBank, logp_from_class_lse
from .kernel import get_kernel
from .metric import support_influence
from .nw import NWHead, NWNet
from .serving import FullModePredictor
from .utils import compute_clusters

__all__ = ["NWHead", "NWNet", "SupportBank", "get_kernel", "support_influence", "compute_clusters",
           "logp_from_class_lse", "FullModePredictor", "_abi"]

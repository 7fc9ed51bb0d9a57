"""`nwhead.kernel` of the reference (nwhead/kernel.py:13-97) -> nwhead_b200.kernel."""
from nwhead_b200.kernel import (Clip, CosineDistance, DotProduct, EuclideanDistance,  # noqa: F401
                                HypersphereEuclideanDistance, get_kernel)

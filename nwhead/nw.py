"""`nwhead.nw` of the reference (nwhead/nw.py: NWNet :12-249, NWHead :251-289) -> nwhead_b200.nw."""
from nwhead_b200.nw import NWHead, NWNet  # noqa: F401

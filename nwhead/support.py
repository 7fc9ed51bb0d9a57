"""`nwhead.support` of the reference (nwhead/support.py:7-165) -> nwhead_b200.support."""
from nwhead_b200.support import SupportSet, SupportSetEval, SupportSetTrain  # noqa: F401

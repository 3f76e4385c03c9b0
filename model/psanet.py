"""model.psanet of the reference (model/psanet.py) re-exported from the B200-native implementation."""
from semseg_b200.psanet import PSA, PSANet  # noqa: F401

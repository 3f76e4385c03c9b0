"""model.pspnet of the reference (model/pspnet.py) re-exported from the B200-native implementation."""
from semseg_b200.pspnet import PPM, PSPNet  # noqa: F401

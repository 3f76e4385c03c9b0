"""Drop-in package: `from model.pspnet import PSPNet` / `from model.psanet import PSANet` (tool/train.py:123,127)
resolve to the B200-native implementation when this repository precedes the reference on PYTHONPATH."""

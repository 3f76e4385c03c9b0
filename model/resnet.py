"""model.resnet of the reference (model/resnet.py) re-exported from the B200-native implementation."""
from semseg_b200.resnet import ResNet, Bottleneck, resnet50, resnet101, resnet152  # noqa: F401

"""train.py:31 / utils/KD_loss.py:6 import `HardBinaryConv` from here."""
from bdbnn_b200.modules import HardBinaryConv  # noqa: F401
from bdbnn_b200.resnet import resnet18, resnet34  # noqa: F401

"""train.py:32 imports `HardBinaryConv_cifar` from here; `BinarizeConv2d` is the north-star name."""
from bdbnn_b200.modules import BinarizeConv2d, HardBinaryConv_cifar  # noqa: F401

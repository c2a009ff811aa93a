"""`models` package at the import paths the reference's train.py / utils/KD_loss.py expect
(train.py:27-32, KD_loss.py:6-7). Everything is implemented in bdbnn_b200."""
from . import cifar10, imagenet, bin_module  # noqa: F401

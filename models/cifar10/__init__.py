"""Zero-argument CIFAR builders looked up at train.py:50-52,257,285: `cifar_models.__dict__[arch]()`."""
from bdbnn_b200.resnet import resnet20  # noqa: F401

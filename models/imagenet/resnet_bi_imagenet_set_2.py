"""train.py:30 / utils/KD_loss.py:7 import `HardBinaryConv_react` from here."""
from bdbnn_b200.modules import HardBinaryConv_react  # noqa: F401
from bdbnn_b200.resnet import ResNetImageNet


def resnet18_react(pretrained=False, **kw):
    return ResNetImageNet([2, 2, 2, 2], conv_cls=HardBinaryConv_react, **kw)

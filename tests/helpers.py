"""Shared test helpers."""


def tiny_net(wrapped):
    """Two-conv block with torchvision-style names (conv1 / layer1.0.* / downsample), optionally
    wrapped like nn.DataParallel ('module.' prefix)."""
    import torch.nn as nn
    net = nn.Module()
    net.conv1 = nn.Conv2d(3, 8, 3, bias=False)
    blk = nn.Module()
    blk.conv1 = nn.Conv2d(8, 8, 3, bias=False)
    blk.conv2 = nn.Conv2d(8, 16, 3, bias=False)
    blk.downsample = nn.Sequential(nn.Conv2d(8, 16, 1, bias=False))
    net.layer1 = nn.Sequential(blk)
    net.fc = nn.Linear(16, 4)
    if wrapped:
        outer = nn.Module()
        outer.module = net
        return outer
    return net



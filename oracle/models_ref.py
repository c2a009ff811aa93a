"""ORACLE (test infrastructure, NOT product code) — network shells around the pure-PyTorch binary conv,
written independently of bdbnn_b200.resnet (no product import: topology, names and the KD pairing rule
are restated here so that a product bug cannot hide in a shared helper).

What the reference pins about the topology (the `models/` package itself is absent upstream):
  * torchvision naming: stem `conv1`, stages `layer1..4`, shortcuts `downsample`    (KD_loss.py:60-64)
  * ImageNet ResNet-18 has 19 hookable convs after the stem (train.py:467-470) = 16 binary 3x3 + 3 1x1
  * state_dict keys equal the product's, so weights move between the two with load_state_dict.
Blocks follow the Bi-Real-Net double-shortcut form (one shortcut per binary conv).
This is the CPU implementation `bench.py --impl reference` / `cpu_baseline` times."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import losses_ref
from .binconv_ref import RefBinarizeConv2d, RefBinarizeConv2dCifar


class _RefBlock(nn.Module):
    """out1 = bn1(conv1(x)) + shortcut(x);  out2 = bn2(conv2(out1)) + out1."""

    def __init__(self, cin, cout, stride, conv_cls, shortcut):
        super().__init__()
        self.conv1 = conv_cls(cin, cout, 3, stride, 1)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = conv_cls(cout, cout, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(cout)
        if shortcut == "conv":                       # ImageNet: fp32 1x1 conv + BN, named `downsample`
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))
        else:
            self.downsample = None
        self._pad = cout // 4 if shortcut == "pad" else 0      # CIFAR option-A shortcut (parameter-free)
        self._sub = shortcut == "pad"

    def forward(self, x):
        if self.downsample is not None:
            sc = self.downsample(x)
        elif self._sub:
            sc = F.pad(x[:, :, ::2, ::2], (0, 0, 0, 0, self._pad, self._pad))
        else:
            sc = x
        mid = self.bn1(self.conv1(x)) + sc
        return self.bn2(self.conv2(mid)) + mid


class _RefCifarBlock(_RefBlock):
    """Same arithmetic; the product names the parameter-free CIFAR shortcut `shortcut` (no parameters, so
    the state_dict is unaffected)."""


def _stage(cin, cout, n, stride, conv_cls, kind):
    first = kind if (stride != 1 or cin != cout) else None
    blocks = [_RefBlock(cin, cout, stride, conv_cls, first)]
    blocks += [_RefBlock(cout, cout, 1, conv_cls, None) for _ in range(n - 1)]
    return nn.Sequential(*blocks)


class RefResNetImageNet(nn.Module):
    def __init__(self, layers, num_classes=1000, conv_cls=RefBinarizeConv2d):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = _stage(64, 64, layers[0], 1, conv_cls, "conv")
        self.layer2 = _stage(64, 128, layers[1], 2, conv_cls, "conv")
        self.layer3 = _stage(128, 256, layers[2], 2, conv_cls, "conv")
        self.layer4 = _stage(256, 512, layers[3], 2, conv_cls, "conv")
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(512, num_classes)

    def forward(self, x):
        x = self.maxpool(self.bn1(self.conv1(x)))
        for stage in (self.layer1, self.layer2, self.layer3, self.layer4):
            x = stage(x)
        return self.fc(self.avgpool(x).flatten(1))


class RefResNetCifar(nn.Module):
    def __init__(self, n_blocks=3, num_classes=10, conv_cls=RefBinarizeConv2dCifar):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        self.layer1 = _stage(16, 16, n_blocks, 1, conv_cls, "pad")
        self.layer2 = _stage(16, 32, n_blocks, 2, conv_cls, "pad")
        self.layer3 = _stage(32, 64, n_blocks, 2, conv_cls, "pad")
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(64, num_classes)

    def forward(self, x):
        x = self.bn1(self.conv1(x))
        for stage in (self.layer1, self.layer2, self.layer3):
            x = stage(x)
        return self.fc(self.avgpool(x).flatten(1))


def resnet18_ref(**kw):
    return RefResNetImageNet([2, 2, 2, 2], **kw)


def resnet34_ref(**kw):
    return RefResNetImageNet([3, 4, 6, 3], **kw)


def resnet20_ref(**kw):
    return RefResNetCifar(3, **kw)


def resnet20_fp32(**kw):
    """fp32 teacher of the CIFAR shape (same names/shapes as the student, KD_loss.py:63)."""
    return RefResNetCifar(3, conv_cls=lambda i, o, k, s, p: nn.Conv2d(i, o, k, s, p, bias=False), **kw)


def kd_pairs_ref(model_stud, model_teacher, binary_types=()):
    """Pairing rule of DistributionLoss_layer.forward (utils/KD_loss.py:59-66), restated with the same
    double loop: teacher modules that are Conv2d (or one of the binary classes — all of which subclass
    Conv2d here) and are not named 'module.conv1'; student module of the SAME name, unless the name
    contains 'downsample'."""
    pairs = []
    for name, module in model_teacher.named_modules():
        if isinstance(module, (nn.Conv2d,) + tuple(binary_types)) and name != 'module.conv1':
            for name_s, module_s in model_stud.named_modules():
                if name_s == name and 'downsample' not in name:
                    pairs.append((name, module_s, module))
    return pairs


class RefOps:
    """Loss terms exactly as the reference computes them (kurtosis.py / utils/KD_loss.py restated), in the
    shape bdbnn_b200.step.TrainStep's `ops` argument takes — used by the host-logic tests that drive the
    product's step driver on CPU (gloo all-reduce)."""

    @staticmethod
    def kurtosis(weights, targets, mode, n_hooks, lam):
        losses = [losses_ref.kurtosis_ref(w, t)[1] for w, t in zip(weights, targets)]
        return losses_ref.aggregate_kurtosis(losses, mode, n_hooks, lam)

    kd_logits = staticmethod(losses_ref.kd_logits_ref)

    @staticmethod
    def kd_layer(out_s, out_t, model_s, model_t, T):
        pairs = kd_pairs_ref(model_s, model_t)
        return losses_ref.kd_layer_ref([p[1].weight for p in pairs], [p[2].weight for p in pairs])

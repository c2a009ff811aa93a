"""Operator surface of the hot path: the binarised conv2d modules the reference imports.

The reference imports three class names from a `models` package that upstream never committed
(train.py:30-32, utils/KD_loss.py:6-7):

    models.imagenet.resnet_bi_imagenet_set_2_2.HardBinaryConv
    models.imagenet.resnet_bi_imagenet_set_2.HardBinaryConv_react
    models.bin_module.binarized_modules.HardBinaryConv_cifar

All three are provided here as thin subclasses of `BinarizeConv2d` (the name BASELINE.json's
north_star uses) and re-exported at those module paths by the top-level `models/` package.

Contract kept: nn.Conv2d subclass (so `isinstance(m, nn.Conv2d)` at train.py:392,413
holds and the EDE loop can assign `.k` / `.t`), one 4-D fp32 parameter named `weight`, no bias,
forward(x[N,Cin,H,W]) -> [N,Cout,Ho,Wo], differentiable through torch.autograd.
"""
import torch
import torch.nn as nn

from .functional import binconv2d, max_pool2d_nhwc


class BinarizeConv2d(nn.Conv2d):
    """1W/1A conv2d: y = mean|W|_o * conv2d(sign(x), sign(W)), STE backward.

    `impl`: None/'auto' (tcgen05 implicit GEMM when the shape qualifies, else XNOR-popcount),
    'xnor' (bit-serial CUDA-core kernels) or 'tc'.  Runs only on CUDA tensors."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False,
                 impl=None):
        if bias:
            raise ValueError("BinarizeConv2d has no bias (the reference's binary convs are bias-free)")
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias=False)
        if self.stride[0] != self.stride[1] or self.padding[0] != self.padding[1]:
            raise ValueError("BinarizeConv2d needs equal stride/padding in h and w")
        if self.dilation != (1, 1) or self.groups != 1:
            raise ValueError("BinarizeConv2d supports dilation=1, groups=1 only")
        self.impl = impl
        # EDE attributes the reference writes every epoch onto every nn.Conv2d (train.py:409-415).
        # Plain attributes here: only HardBinaryConv_cifar gives them a meaning.
        self._ede_k = None
        self._ede_t = None

    # `.k` / `.t`: readable defaults, and assignment (what `--ede` does each epoch) is recorded.
    @property
    def k(self):
        return self._ede_k if self._ede_k is not None else torch.tensor([1.0])

    @k.setter
    def k(self, value):
        self._ede_k = value

    @property
    def t(self):
        return self._ede_t if self._ede_t is not None else torch.tensor([1.0])

    @t.setter
    def t(self, value):
        self._ede_t = value

    #: classes that honour an assigned (k, t) pair in their backward set this to True
    supports_ede = False

    @property
    def ede_active(self):
        return self.supports_ede and self._ede_k is not None and self._ede_t is not None

    def _ede(self, x):
        if not self.ede_active:
            return None
        k, t = torch.as_tensor(self._ede_k), torch.as_tensor(self._ede_t)
        return (k.to(x.device, torch.float32), t.to(x.device, torch.float32))

    def forward(self, x):
        return binconv2d(x, self.weight, self.stride[0], self.padding[0], self.impl, self._ede(x))

    def extra_repr(self):
        return super().extra_repr() + f", binarized=1W/1A, impl={self.impl or 'auto'}"


class HardBinaryConv(BinarizeConv2d):
    """ImageNet 'set_2_2' binary conv (train.py:31, KD_loss.py:6)."""

    def __init__(self, in_chn, out_chn, kernel_size=3, stride=1, padding=1, **kw):
        super().__init__(in_chn, out_chn, kernel_size, stride, padding, **kw)


class HardBinaryConv_react(BinarizeConv2d):
    """ImageNet 'set_2' (ReAct-style training recipe) binary conv (train.py:30, KD_loss.py:7).
    The --react flag only changes loss weights (train.py:605-609); the conv contract is identical."""

    def __init__(self, in_chn, out_chn, kernel_size=3, stride=1, padding=1, **kw):
        super().__init__(in_chn, out_chn, kernel_size, stride, padding, **kw)


class HardBinaryConv_cifar(BinarizeConv2d):
    """CIFAR binary conv (train.py:32,392); carries the EDE `.k/.t` tensors (train.py:412-415).

    Until the training loop assigns `.k` and `.t` the backward is the hard-tanh STE of DESIGN.md §2;
    once both are assigned (the reference's `--ede`, schedule utils/utils.py:8-14) both STE
    derivatives become k*t*(1 - tanh(t*v)^2) (v = x for the activation, v = W for the weight)."""

    supports_ede = True

    def __init__(self, in_chn, out_chn, kernel_size=3, stride=1, padding=1, **kw):
        super().__init__(in_chn, out_chn, kernel_size, stride, padding, **kw)


class MaxPool2dNHWC(nn.Module):
    """nn.MaxPool2d(kernel_size, stride, padding) semantics on channels_last CUDA tensors."""

    def __init__(self, kernel_size, stride=None, padding=0):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride or kernel_size, padding

    def forward(self, x):
        return max_pool2d_nhwc(x, self.kernel_size, self.stride, self.padding)

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}, stride={self.stride}, padding={self.padding}"

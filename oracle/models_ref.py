"""ORACLE (test infrastructure) — the same network shells as bdbnn_b200.resnet, built around the
pure-PyTorch RefBinarizeConv2d, plus pure-PyTorch loss ops for the shared TrainStep driver.
This is the CPU implementation `bench.py --impl reference` times."""
import torch
import torch.nn as nn

from bdbnn_b200 import resnet as _resnet
from bdbnn_b200.losses import matched_weight_pairs

from .binconv_ref import RefBinarizeConv2d, RefBinarizeConv2dCifar
from . import losses_ref


def resnet18_ref(**kw):
    return _resnet.ResNetImageNet([2, 2, 2, 2], conv_cls=RefBinarizeConv2d, pool_cls=nn.MaxPool2d, **kw)


def resnet34_ref(**kw):
    return _resnet.ResNetImageNet([3, 4, 6, 3], conv_cls=RefBinarizeConv2d, pool_cls=nn.MaxPool2d, **kw)


def resnet20_ref(**kw):
    return _resnet.ResNetCifar(3, conv_cls=RefBinarizeConv2dCifar, **kw)


class RefOps:
    """Loss terms exactly as the reference computes them (kurtosis.py / utils/KD_loss.py restated)."""

    @staticmethod
    def kurtosis(weights, targets, mode, n_hooks, lam):
        losses = [losses_ref.kurtosis_ref(w, t)[1] for w, t in zip(weights, targets)]
        return losses_ref.aggregate_kurtosis(losses, mode, n_hooks, lam)

    kd_logits = staticmethod(losses_ref.kd_logits_ref)

    @staticmethod
    def kd_layer(out_s, out_t, model_s, model_t, T):
        pairs = matched_weight_pairs(model_s, model_t)
        return losses_ref.kd_layer_ref([p[1].weight for p in pairs], [p[2].weight for p in pairs])

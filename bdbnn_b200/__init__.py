"""bdbnn_b200 — B200-native (sm_100a) implementation of BD-BNN's training hot path:
binarised conv2d fwd/bwd, kurtosis regulariser, KD losses."""
from .modules import BinarizeConv2d, HardBinaryConv, HardBinaryConv_react, HardBinaryConv_cifar
from .losses import (KurtosisWeight, DistributionLoss, DistributionLoss_layer,
                     kurtosis_regularization, matched_weight_pairs)
from .functional import binconv2d, kurtosis_multi, kd_logits_loss, kd_layer_loss

__all__ = ["BinarizeConv2d", "HardBinaryConv", "HardBinaryConv_react", "HardBinaryConv_cifar",
           "KurtosisWeight", "DistributionLoss", "DistributionLoss_layer", "kurtosis_regularization",
           "matched_weight_pairs", "binconv2d", "kurtosis_multi", "kd_logits_loss", "kd_layer_loss"]

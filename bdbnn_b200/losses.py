"""Loss surface of the hot path, same class names / call signatures as the reference:

    KurtosisWeight            kurtosis.py:5-39      (used at train.py:479,501-504 / 598,622-625)
    DistributionLoss          utils/KD_loss.py:10-43   (criterion_kl_c, train.py:341,612)
    DistributionLoss_layer    utils/KD_loss.py:46-67   (criterion_kl,   train.py:340,611)

plus `kurtosis_regularization()`, the fused multi-layer form the step driver uses (one launch for all
19 hooked layers instead of 19 x ~21 ATen launches)."""
import torch
import torch.nn as nn
from torch.nn.modules import loss as _loss

from . import functional as F_
from .modules import HardBinaryConv, HardBinaryConv_react


class KurtosisWeight:
    """Drop-in for kurtosis.py:5-20.  fn_regularization() returns None and leaves the result on
    `.kurtosis_loss` (0-d, attached to the autograd graph) / `.kurtosis`; `.KLDiv_loss` stays 0."""

    def __init__(self, weight_tensor, name, kurtosis_target=2.0, k_mode='avg', KLD=False):
        self.kurtosis_loss = 0
        self.kurtosis = 0
        self.weight_tensor = weight_tensor
        self.name = name
        self.k_mode = k_mode
        self.kurtosis_target = kurtosis_target
        self.KLDiv_loss = 0
        self.KLD = KLD

    def fn_regularization(self):
        return self.kurtosis_calc()

    def kurtosis_calc(self):
        if isinstance(self.kurtosis_target, (list, tuple)):
            # the reference would raise `Tensor - list` here (latent bug B3)
            raise TypeError("kurtosis_target must be a scalar")
        loss, kurt = F_.kurtosis_multi([self.weight_tensor], [float(self.kurtosis_target)])
        # k_mode 'avg' | 'max' | 'sum' are identities on a 0-d value (kurtosis.py:31-39)
        self.kurtosis_loss = loss[0]
        self.kurtosis = kurt[0]


class RidgeRegularization:
    """kurtosis.py:42-53 (imported at train.py:29; only instantiated under `args.w_l2_reg`, which the
    reference's parser never defines — dead upstream, kept so the import block of train.py resolves).
    `l2_regularization()` returns None and leaves sum(W^2) on `.l2_loss`.  Plain torch ops: this object
    is not on the hot path."""

    def __init__(self, weight_tensor, name):
        self.weight_tensor, self.name, self.l2_loss = weight_tensor, name, 0

    def l2_calc(self):
        self.l2_loss = self.weight_tensor.pow(2).sum()

    def l2_regularization(self):
        return self.l2_calc()


class WeightRegularization:
    """kurtosis.py:56-70 (same status as RidgeRegularization, flag `args.w_wr_reg`):
    `.wr_loss` = || |W| - 1 ||_2 over the whole tensor; `.size` = element count."""

    def __init__(self, weight_tensor, name):
        self.weight_tensor, self.name, self.wr_loss = weight_tensor, name, 0
        self.size = int(weight_tensor.numel())

    def wr_calc(self):
        self.wr_loss = torch.linalg.vector_norm(self.weight_tensor.abs() - 1, ord=2)

    def w_regularization(self):
        return self.wr_calc()


def kurtosis_regularization(weights, targets, mode='avg', n_hooks=None, lam=1.0):
    """train.py:495-513 for all hooked layers at once. Returns (regulariser, per-layer losses, kurtosis).
    mode: 'sum' | 'avg' (sum / len(weight_to_hook)) | 'max'."""
    losses, kurt = F_.kurtosis_multi(list(weights), list(targets))
    n_hooks = len(weights) if n_hooks is None else n_hooks
    if mode == 'sum':
        tot = losses.sum()
    elif mode == 'avg':
        tot = losses.sum() / n_hooks
    elif mode == 'max':
        tot = losses.max()
    else:
        raise ValueError(f"kurtosis mode {mode!r}")
    return (10 ** 0) * lam * tot, losses, kurt


class DistributionLoss(_loss._Loss):
    """-(1/N) sum_n sum_c softmax(teacher) * log_softmax(student); see utils/KD_loss.py:16-43."""

    def forward(self, stud_output, teacher_output):
        self.size_average = True
        if teacher_output.requires_grad:
            raise ValueError("real network output should not require gradients.")
        return F_.kd_logits_loss(stud_output, teacher_output)


def matched_weight_pairs(model_stud, model_teacher):
    """Pairing rule of DistributionLoss_layer.forward (utils/KD_loss.py:59-66), literally: every teacher
    module that is a Conv2d / HardBinaryConv / HardBinaryConv_react and whose name is not
    'module.conv1', paired with the student module of the same name unless 'downsample' is in it."""
    stud = dict(model_stud.named_modules())
    pairs = []
    for name, module in model_teacher.named_modules():
        if isinstance(module, (nn.Conv2d, HardBinaryConv, HardBinaryConv_react)) and name != 'module.conv1':
            m_s = stud.get(name)
            if m_s is not None and 'downsample' not in name:
                pairs.append((name, m_s, module))
    return pairs


class DistributionLoss_layer(_loss._Loss):
    """sum over matched layers of KLDivLoss(log_target=True)(W_student, W_teacher)
    (utils/KD_loss.py:52-67).  `T` is accepted and unused, as in the reference."""

    def forward(self, stud_output, teacher_output, model_stud, model_teacher, T=1):
        pairs = matched_weight_pairs(model_stud, model_teacher)
        if not pairs:
            return 0
        return F_.kd_layer_loss([p[1].weight for p in pairs], [p[2].weight for p in pairs])

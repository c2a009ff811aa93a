"""Drop-in for the reference's utils/utils.py (`from utils import utils`, train.py:26): the helper
names train.py calls — cpt_tk (:411), find_weight_tensor_by_name (:402), save_checkpoint (:433),
AverageMeter / ProgressMeter (:442-454), accuracy (:518).

`utils/` is deliberately a namespace package (no __init__.py), as it is upstream: a maintainer who keeps
their own utils/utils.py next to this repo's utils/KD_loss.py gets both.  cpt_tk and accuracy are the
step pieces of the hot path and live in bdbnn_b200.step; the rest is small host-side bookkeeping with the
attribute names train.py reads (`.avg`, `.get_avg()`, `.display(i)`)."""
import os
import shutil

import torch

from bdbnn_b200.step import accuracy, cpt_tk as _cpt_tk

T_min_g, T_max_g = 1e-2, 1e1


def cpt_tk(epoch, tot_epochs):
    """(t, k) of the EDE schedule, 1-element fp32 tensors (reference utils/utils.py:8-14)."""
    return _cpt_tk(epoch, tot_epochs, T_min_g, T_max_g)


def find_weight_tensor_by_name(model, name_in):
    """The parameter registered under `name_in`, or None (reference utils/utils.py:16-19)."""
    return dict(model.named_parameters()).get(name_in)


def save_checkpoint(state, is_best, save_path):
    """checkpoint.pth.tar (+ model_best.pth.tar copy when is_best) under save_path (utils/utils.py:21-25)."""
    target = os.path.join(save_path, 'checkpoint.pth.tar')
    torch.save(state, target)
    if is_best:
        shutil.copyfile(target, os.path.join(save_path, 'model_best.pth.tar'))


class AverageMeter:
    """Running value / sum / count / avg with the reference's `name val (avg)` rendering."""

    def __init__(self, name, fmt=':f'):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum = self.sum + val * n
        self.count = self.count + n
        self.avg = self.sum / self.count

    def get_avg(self):
        return self.avg

    def __str__(self):
        return ('{name} {val' + self.fmt + '} ({avg' + self.fmt + '})').format(
            name=self.name, val=self.val, avg=self.avg)


class ProgressMeter:
    """`prefix[ i/N]<TAB>meter<TAB>meter...` lines through logger.info (train.py:450-454, 532)."""

    def __init__(self, num_batches, meters, logger, prefix=""):
        width = len(str(int(num_batches)))
        self._batch = '[{:' + str(width) + 'd}/' + ('{:' + str(width) + 'd}').format(int(num_batches)) + ']'
        self.meters, self.prefix, self.logger = meters, prefix, logger

    def display(self, batch):
        self.logger.info('\t'.join([self.prefix + self._batch.format(batch)] + [str(m) for m in self.meters]))


__all__ = ["cpt_tk", "find_weight_tensor_by_name", "save_checkpoint", "AverageMeter", "ProgressMeter", "accuracy",
           "T_min_g", "T_max_g"]

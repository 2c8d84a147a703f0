"""Optimizer state and learning-rate schedules of the adapter path.

`AdapterSGD` stands in for the `torch.optim.SGD` objects built by the reference's set_optimizer /
set_optimizer_reg (demo/util.py:118-136): it exposes `param_groups[0]['lr']` (the only field the
schedules touch) and owns the flat gradient / momentum buffers the fused kernels update.  The update rule
itself (weight decay on every tensor, momentum buffer initialised with the first gradient, no dampening /
Nesterov) is executed by libdbmm's k_sgd.

The schedule helpers keep the reference's names and call signatures (demo/util.py:70-115).
"""
from __future__ import annotations

import math

import numpy as np

from . import ops


class AdapterSGD:
    def __init__(self, adapter_module, lr, momentum, weight_decay):
        self.adapter_module = adapter_module
        self.param_groups = [{"lr": lr, "momentum": momentum, "weight_decay": weight_decay}]
        t = adapter_module.tensors()
        self.buffers = ops.TrainBuffers(t.D, t.H, device=t.W1.device)

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @property
    def momentum(self):
        return self.param_groups[0]["momentum"]

    @property
    def weight_decay(self):
        return self.param_groups[0]["weight_decay"]

    def zero_grad(self):
        self.buffers.grads.zero_()


class LinearSGD:
    """Optimizer state of a LinearClassifier (linear probing): flat gradient / momentum buffers W | b."""

    def __init__(self, module, lr, momentum, weight_decay):
        self.module = module
        self.param_groups = [{"lr": lr, "momentum": momentum, "weight_decay": weight_decay}]
        w = module.fc.weight
        n = w.numel() + module.fc.bias.numel()
        import torch
        self.grads = torch.zeros(n, dtype=torch.float32, device=w.device)
        self.momentum_buf = torch.zeros(n, dtype=torch.float32, device=w.device)
        self.first_step = True

    lr = AdapterSGD.lr
    momentum = AdapterSGD.momentum
    weight_decay = AdapterSGD.weight_decay

    def zero_grad(self):
        self.grads.zero_()


def trainable_adapter(model):
    """The adapter whose tensors the optimizer updates: `new_adapter` of a MultipleAdapter (stage 2 freezes every
    parameter whose name contains "old_cls", demo/util.py:128), else the classifier's only adapter."""
    return model.new_adapter if hasattr(model, "new_adapter") else model.adapter


def set_optimizer(opt, model):
    if hasattr(model, "fc"):                                      # LinearClassifier (--tl_method linear_probing)
        return LinearSGD(model, opt.learning_rate, opt.momentum, opt.weight_decay)
    return AdapterSGD(trainable_adapter(model), opt.learning_rate, opt.momentum, opt.weight_decay)


def set_optimizer_reg(opt, model, freeze_old=True):
    if not freeze_old and hasattr(model, "new_adapter"):
        raise NotImplementedError("freeze_old=False is never used by the reference's drivers")
    return AdapterSGD(trainable_adapter(model), opt.learning_rate_reg, opt.momentum, opt.weight_decay)


def get_lr(optimizer):
    return optimizer.param_groups[0]["lr"]


def _set_lr(optimizer, lr):
    for group in optimizer.param_groups:
        group["lr"] = lr


def _decayed(base_lr, epoch, args, cosine_span):
    if args.cosine:
        eta_min = base_lr * (args.lr_decay_rate ** 3)
        return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * epoch / cosine_span)) / 2
    passed = np.sum(epoch > np.asarray(args.lr_decay_epochs))
    return base_lr * (args.lr_decay_rate ** passed) if passed > 0 else base_lr


def adjust_learning_rate(args, optimizer, epoch):
    _set_lr(optimizer, _decayed(args.learning_rate, epoch, args, args.epochs))


def adjust_learning_rate_reg(args, optimizer, epoch):
    # The reference's cosine branch reads a misspelt attribute (demo/util.py:89) and crashes; the intended
    # span (epochs - epochs_feature_learning) is used here instead of reproducing the crash.
    span = args.epochs - (args.epochs_feature_learning or 0)
    _set_lr(optimizer, _decayed(args.learning_rate_reg, epoch, args, span))


def _warm(optimizer, epoch, batch_id, total_batches, warm_epochs, lo, hi):
    if epoch <= warm_epochs:
        p = (batch_id + (epoch - 1) * total_batches) / (warm_epochs * total_batches)
        _set_lr(optimizer, lo + p * (hi - lo))


def warmup_learning_rate(args, epoch, batch_id, total_batches, optimizer):
    if args.warm:
        _warm(optimizer, epoch, batch_id, total_batches, args.warm_epochs, args.warmup_from, args.warmup_to)


def warmup_learning_rate_reg(args, epoch, batch_id, total_batches, optimizer):
    if args.warm_reg:
        _warm(optimizer, epoch, batch_id, total_batches, args.warm_epochs_reg, args.warmup_from_reg, args.warmup_to_reg)

"""Train-step mirror of /root/reference/src/dl/train.py: ``ModelEMA`` (52-73) and the body of
``Trainer.train`` for one batch (550-586) + ``optimizer_step`` (512-535), fp32 path
(``amp_enabled=False``, train.py:577-581).  Data loading, evaluation, logging and checkpoint
writing are out of scope; ``src/dl/train.py`` wires this into the
reference's ``Trainer`` / ``main`` names.
"""
from __future__ import annotations

import math
from copy import deepcopy

import torch
from torch.nn.parallel import DistributedDataParallel as DDP


class ModelEMA:
    """EMA of every floating-point state entry, momentum m*(1-exp(-it/2000)) (train.py:52-73).
    The reference walks ~920 tensors with two tiny kernels each; here the entries are updated with
    two multi-tensor launches."""

    def __init__(self, student, ema_momentum):
        if isinstance(student, DDP):
            student = student.module
        self.model = deepcopy(student).eval()
        for p in self.model.parameters():
            p.requires_grad_(False)
        self.ema_momentum = ema_momentum
        self._pairs = None

    def ema_scheduler(self, x):
        return self.ema_momentum * (1 - math.exp(-x / 2000))

    def _build_pairs(self, student):
        src = student.state_dict()
        ema, stu = [], []
        for name, p in self.model.state_dict().items():
            if p.dtype.is_floating_point:
                ema.append(p)
                stu.append(src[name].detach())
        self._pairs = (ema, stu)

    @torch.no_grad()
    def update(self, iters, student):
        if isinstance(student, DDP):
            student = student.module
        if self._pairs is None:
            self._build_pairs(student)
        m = self.ema_scheduler(iters)
        ema, stu = self._pairs
        torch._foreach_mul_(ema, m)
        torch._foreach_add_(ema, stu, alpha=1.0 - m)


class TrainStep:
    """One optimisation step on one batch: forward, criterion, backward, clip, AdamW, scheduler,
    zero_grad, EMA — the reference's hot loop body."""

    def __init__(self, model, loss_fn, optimizer, scheduler=None, ema=None, clip_max_norm=0.1, accum_steps=1):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.scheduler, self.ema, self.clip_max_norm = scheduler, ema, clip_max_norm
        self.accum_steps, self.ema_iter, self.batch_idx = accum_steps, 0, 0

    def optimizer_step(self, step_scheduler=True):
        if self.clip_max_norm:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip_max_norm)
        self.optimizer.step()
        if step_scheduler and self.scheduler is not None:
            self.scheduler.step()
        self.optimizer.zero_grad()
        if self.ema is not None:
            self.ema_iter += 1
            self.ema.update(self.ema_iter, self.model)

    def __call__(self, inputs, targets):
        """inputs float32 [B,3,H,W] on the device; targets list of dicts (labels int64 [T], boxes [T,4])."""
        output = self.model(inputs, targets=targets)
        loss_dict = self.loss_fn(output, targets)
        loss = sum(loss_dict.values()) / self.accum_steps
        loss.backward()
        self.batch_idx += 1
        if self.batch_idx % self.accum_steps == 0:
            self.optimizer_step()
        return loss.detach(), loss_dict

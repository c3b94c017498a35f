"""Fused losses of the hot path (host-side mirrors of the reference's inline loss code).

* ``CrossEntropyLoss2d(weight)(logits, targets)`` — same name, constructor and call as the class the drivers
  define inline (train_new_task_step2.py:84-92, train_new_task_step3.py:85-93, train_RAPFT_step1.py:89-97).
* ``OutputKD()(student_logits, teacher_logits)`` — what the drivers compute as
  ``KLDivLoss()(F.softmax(student, 1), F.softmax(teacher, 1))`` (train_new_task_step2.py:241,296-297).

Both read the logits once and produce the loss and the logit gradient in the same kernel.
"""
from __future__ import annotations

import torch

from . import functional as F_


class CrossEntropyLoss2d(torch.nn.Module):
    """``global_norm=True`` (one process per GPU only): normalise by the class-weight sum of the GLOBAL batch, as
    nn.DataParallel's loss on the gathered logits does (train_new_task_step2.py:293,474), instead of each rank's own
    sum — one all-reduce of two fp64 scalars.  ``group`` = the torch.distributed group (default: world)."""

    def __init__(self, weight=None, global_norm: bool = False, group=None):
        super().__init__()
        self.register_buffer("weight", None if weight is None else torch.as_tensor(weight, dtype=torch.float32))
        self.global_norm, self.group = bool(global_norm), group

    def forward(self, outputs, targets):
        gn = None if not self.global_norm else (self.group if self.group is not None else False)
        return F_.CrossEntropy2dFn.apply(outputs, targets, self.weight, gn)


FusedCrossEntropyLoss2d = CrossEntropyLoss2d


class OutputKD(torch.nn.Module):
    def forward(self, student_logits, teacher_logits):
        return F_.OutputKDFn.apply(student_logits, teacher_logits)


FusedOutputKD = OutputKD

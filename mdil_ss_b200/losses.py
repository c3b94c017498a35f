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
    def __init__(self, weight=None):
        super().__init__()
        self.register_buffer("weight", None if weight is None else torch.as_tensor(weight, dtype=torch.float32))

    def forward(self, outputs, targets):
        return F_.CrossEntropy2dFn.apply(outputs, targets, self.weight)


FusedCrossEntropyLoss2d = CrossEntropyLoss2d


class OutputKD(torch.nn.Module):
    def forward(self, student_logits, teacher_logits):
        return F_.OutputKDFn.apply(student_logits, teacher_logits)


FusedOutputKD = OutputKD

"""Validation metric with the reference's iouEval API (iouEval.py:8-77), computed from ONE fused pass over the
logits (argmax over classes + C x C confusion histogram) instead of two one-hot [N,C,H,W] tensors and nine
reductions.  tp/fp/fn are accumulated on the host in float64 exactly like the reference so getIoU() is identical
whenever the argmax is."""
from __future__ import annotations

import torch

from . import functional as F_


class iouEval:
    def __init__(self, nClasses, ignoreIndex=19):
        self.nClasses = nClasses
        self.ignoreIndex = ignoreIndex if nClasses > ignoreIndex else -1
        self.reset()

    def reset(self):
        classes = self.nClasses if self.ignoreIndex == -1 else self.nClasses - 1
        self.tp = torch.zeros(classes).double()
        self.fp = torch.zeros(classes).double()
        self.fn = torch.zeros(classes).double()

    def add_confusion(self, conf: torch.Tensor):
        """conf[gt, pred] counts (int64 [C,C])."""
        conf = conf.double().cpu()
        k = self.nClasses if self.ignoreIndex == -1 else self.ignoreIndex
        if self.ignoreIndex != -1:
            assert self.ignoreIndex == self.nClasses - 1, "only 'ignore = last class' (the drivers' usage) is supported"
        valid = conf[:k]                       # rows with gt == ignore are dropped (iouEval.py:49-52,61-62)
        tp = valid.diagonal()[:k]
        self.tp += tp
        self.fp += valid[:, :k].sum(0) - tp    # predicted c, gt another non-ignored class
        self.fn += valid.sum(1) - tp           # gt c, predicted anything else (incl. the ignore class)

    def addLogits(self, logits: torch.Tensor, y: torch.Tensor):
        """Fused path: logits [N,C,H,W] (CUDA), y labels [N,1,H,W] or [N,H,W] int64."""
        _, conf = F_.argmax_confusion(logits, y)
        self.add_confusion(conf)

    def addBatch(self, x, y):
        """Reference signature: x = predictions [N,1,H,W] int64 (outputs.max(1)[1].unsqueeze(1)), y = labels [N,1,H,W]."""
        c = self.nClasses
        idx = (y.reshape(-1).long() * c + x.reshape(-1).long())
        conf = torch.bincount(idx, minlength=c * c).reshape(c, c)
        self.add_confusion(conf)

    def getIoU(self):
        num = self.tp
        den = self.tp + self.fp + self.fn + 1e-15
        iou = num / den
        return torch.mean(iou), iou

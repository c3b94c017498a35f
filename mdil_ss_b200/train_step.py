"""The reference drivers' training iterations, restated over the B200 modules (model + fused losses + one
gradient all-reduce + fused Adam).  The drivers themselves (train_RAPFT_step1.py, train_new_task_step2.py,
train_new_task_step3.py) stay usable unchanged with ``models.erfnet_RA_parallel.Net``; these classes are what
bench.py and the tests drive, and what a torchrun launcher would call per iteration.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .losses import CrossEntropyLoss2d, OutputKD
from .parallel import FlatAdam

# class-weight literals: train_new_task_step2.py:121-131 (IDD / BDD / Cityscapes), last class zeroed (:133-135)
WEIGHT_IDD = [3.235635601598852, 6.76221624390441, 9.458242359884549, 9.446818215454014, 9.947040673126763,
              9.789672819856547, 9.476665808564432, 10.465565126694731, 9.59189547383129, 7.637805282159825,
              8.990899026692638, 9.26222234098628, 10.265657138809514, 9.386517631614392, 8.357391489170013,
              9.910382864314824, 10.389977663948363, 8.997422571963602, 10.418070541191673, 10.483262606962834,
              9.511436923349441, 7.597725385711079, 6.1734896019878205, 9.787631041755187, 3.9178330193378708,
              4.417448652936843, 0.0]
WEIGHT_BDD = [3.6525147483016243, 8.799815287822142, 4.781908267406055, 10.034828238618045, 9.5567865464289,
              9.645099012085169, 10.315292989325766, 10.163473632969513, 4.791692009441432, 9.556915153488912,
              4.142994047786311, 10.246903827488143, 10.47145010979545, 6.006704177894196, 9.60620532303246,
              9.964959813857726, 10.478333987902301, 10.468010534454706, 10.440929141422366, 0.0]
WEIGHT_CITY = [2.8159904084894922, 6.9874672455551075, 3.7901719017455604, 9.94305485286704, 9.77037625072462,
               9.511470001589007, 10.310780572569994, 10.025305236316246, 4.6341256102158805, 9.561389195953845,
               7.869695292372276, 9.518873463871952, 10.374050047877898, 6.662394711556909, 10.26054487392723,
               10.28786101490449, 10.289883605859952, 10.405463349170795, 10.138502340710136, 0.0]


def class_weights(dataset: str, device=None) -> torch.Tensor:
    table = {"cityscapes": WEIGHT_CITY, "BDD": WEIGHT_BDD, "IDD": WEIGHT_IDD}
    return torch.tensor(table[dataset], dtype=torch.float32, device=device)


def is_shared(n: str) -> bool:
    """train_new_task_step2.py:95-96."""
    return "encoder" in n and "parallel_conv" not in n and "bn" not in n


def is_DS_curr(n: str, current_task: int) -> bool:
    """train_new_task_step2.py:99-105."""
    if "decoder.{}".format(current_task) in n:
        return True
    if "encoder" in n and ("bn" in n or "parallel_conv" in n):
        return ".{}.weight".format(current_task) in n or ".{}.bias".format(current_task) in n
    return False


def apply_incremental_freeze(model: torch.nn.Module, current_task: int) -> None:
    """Freezing policy of steps 2/3 (train_new_task_step2.py:205-215)."""
    for name, m in model.named_parameters():
        if "decoder" in name:
            if "decoder.{}".format(current_task) not in name:
                m.requires_grad = False
        elif "encoder" in name:
            if "bn" in name or "parallel_conv" in name:
                if ".{}.weight".format(current_task) in name or ".{}.bias".format(current_task) in name:
                    continue
                m.requires_grad = False


def incremental_param_groups(model: torch.nn.Module, current_task: int) -> List[dict]:
    """train_new_task_step2.py:229-235: shared encoder convs at lr 5e-6, current-domain parameters at the base lr."""
    params = list(model.named_parameters())
    return [{"params": [p for n, p in params if is_shared(n)], "lr": 5e-6},
            {"params": [p for n, p in params if is_DS_curr(n, current_task)]}]


def poly_lr_factor(epoch: int, num_epochs: int) -> float:
    """lambda1 (train_new_task_step2.py:244)."""
    return pow((1 - ((epoch - 1) / num_epochs)), 0.9)


class Step1Trainer:
    """train_RAPFT_step1.py:287-305: fwd(train) -> CE -> backward -> Adam over every parameter."""

    def __init__(self, model, weight: torch.Tensor, task: int = 0, lr: float = 5e-4):
        self.model, self.task = model, task
        self.criterion = CrossEntropyLoss2d(weight, global_norm=True).to(weight.device)   # DataParallel's global sum(w)
        self.optimizer = FlatAdam([{"params": list(model.parameters())}], lr)

    def step(self, images: torch.Tensor, labels: torch.Tensor):
        self.model.train()
        outputs = self.model(images, self.task)
        self.optimizer.zero_grad()
        loss = self.criterion(outputs, labels[:, 0])
        loss.backward()
        self.optimizer.step()
        return loss.detach()


class Step2Trainer:
    """train_new_task_step2.py:271-306: student fwd on domain t and t-1 (train mode), frozen teacher fwd on t-1
    (eval mode), CE + lambda * KD, one backward, 2-group Adam."""

    def __init__(self, model, model_old, weight: torch.Tensor, task: int, lambdac: float = 0.1):
        self.model, self.model_old, self.task, self.lambdac = model, model_old, task, lambdac
        for p in model_old.parameters():
            p.requires_grad = False
        apply_incremental_freeze(model, task)
        self.criterion = CrossEntropyLoss2d(weight, global_norm=True).to(weight.device)   # DataParallel's global sum(w)
        self.kd = OutputKD()
        self.optimizer = FlatAdam(incremental_param_groups(model, task), 5e-4)

    def step(self, images: torch.Tensor, labels: torch.Tensor):
        self.model.train()
        self.model_old.eval()
        outputs = self.model(images, self.task)
        outputs_prev_task = self.model(images, self.task - 1)
        with torch.no_grad():
            outputs_prev_model = self.model_old(images, self.task - 1)
        ce_loss = self.criterion(outputs, labels[:, 0])
        kld_loss = self.kd(outputs_prev_task, outputs_prev_model)
        total = ce_loss + self.lambdac * kld_loss
        self.optimizer.zero_grad()
        total.backward()
        self.optimizer.step()
        return total.detach(), ce_loss.detach(), kld_loss.detach()


class Step3Trainer:
    """train_new_task_step3.py:303-356: CE step, then a KD step against the previous model on both old domains
    (two optimiser steps per iteration).  As in the reference the teacher is never switched to eval() (:301)."""

    def __init__(self, model, model_old, weight: torch.Tensor, task: int = 2, lambdac: float = 0.1):
        self.model, self.model_old, self.task, self.lambdac = model, model_old, task, lambdac
        for p in model_old.parameters():
            p.requires_grad = False
        apply_incremental_freeze(model, task)
        self.criterion = CrossEntropyLoss2d(weight, global_norm=True).to(weight.device)   # DataParallel's global sum(w)
        self.kd = OutputKD()
        self.optimizer = FlatAdam(incremental_param_groups(model, task), 5e-4)

    def step(self, images: torch.Tensor, labels: torch.Tensor):
        self.model.train()
        outputs = self.model(images, self.task)
        ce_loss = self.criterion(outputs, labels[:, 0])
        self.optimizer.zero_grad()
        ce_loss.backward()
        self.optimizer.step()
        out_prev_1 = self.model(images, self.task - 1)
        out_prev_0 = self.model(images, self.task - 2)
        with torch.no_grad():
            old_1 = self.model_old(images, self.task - 1)
            old_0 = self.model_old(images, self.task - 2)
        kd_loss = self.lambdac * (self.kd(out_prev_1, old_1) + self.kd(out_prev_0, old_0))
        self.optimizer.zero_grad()
        kd_loss.backward()
        self.optimizer.step()
        return ce_loss.detach(), kd_loss.detach()


class MultiTaskTrainer:
    """train_multi_task.py:209-265 run over the RAP network (BASELINE config 5): every iteration visits the
    datasets in turn -- forward on domain i, its class-weighted CE, backward, one optimiser step -- with the encoder
    at lr 5e-4 / nb_tasks and the decoders at 5e-4 (:209-218).  Tensors a visit does not reach (the other domains'
    adapters, BatchNorms and decoders) are left untouched by that visit's optimiser step, as torch.optim.Adam does."""

    def __init__(self, model, weights: Sequence[torch.Tensor], lr: float = 5e-4):
        self.model = model
        self.nb_tasks = len(weights)
        self.criteria = [CrossEntropyLoss2d(w, global_norm=True).to(w.device) for w in weights]
        params = list(model.named_parameters())
        self.optimizer = FlatAdam([{"params": [p for n, p in params if "encoder" in n], "lr": lr / self.nb_tasks},
                                   {"params": [p for n, p in params if "decoder" in n]}], lr)

    def visit(self, ind: int, images, labels):
        """One dataset visit (train_multi_task.py:247-262): forward on domain ``ind``, CE, backward, optimiser step."""
        self.model.train()
        outputs = self.model(images, ind)
        self.optimizer.zero_grad()
        loss = self.criteria[ind](outputs, labels[:, 0])
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def step(self, batches):
        """``batches[i] = (images_i, labels_i)`` for dataset i; returns the list of CE losses."""
        return [self.visit(ind, images, labels) for ind, (images, labels) in enumerate(batches)]


class GraphedStep:
    """One training iteration captured in a CUDA graph and replayed: ~300 kernel launches, the weight
    re-packing, the gradient all-reduce and the fused Adam become ONE host call, which takes the Python / launch overhead
    (~10 ms per step from idle) off the critical path -- what limits end-to-end scaling when 8 ranks share one host.

    ``GraphedStep(trainer, images, labels)`` warms the trainer up on the given (static-shape) batch, switches its FlatAdam
    to device-side step counters and captures ``trainer.step``; ``step(images, labels)`` copies the new batch into the
    captured input buffers (device-to-device) and replays.  The returned loss tensors are the captured ones (read them
    after the replay, e.g. every k steps).  Usable where every trainable parameter receives a gradient in every step
    (Step1Trainer, Step2Trainer); Step3Trainer's KD step is not (FlatAdam raises).  NCCL collectives inside the step are
    captured with it."""

    def __init__(self, trainer, images: torch.Tensor, labels: torch.Tensor, warmup: int = 3):
        self.trainer = trainer
        self.static_images = images.clone()
        self.static_labels = labels.clone()
        trainer.optimizer.enable_graph_safe()
        side = torch.cuda.Stream(images.device)
        side.wait_stream(torch.cuda.current_stream(images.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):        # also finishes every lazy initialisation (function attributes, caches)
                trainer.step(self.static_images, self.static_labels)
        torch.cuda.current_stream(images.device).wait_stream(side)
        torch.cuda.synchronize(images.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = trainer.step(self.static_images, self.static_labels)

    def step(self, images: torch.Tensor, labels: torch.Tensor):
        if images.data_ptr() != self.static_images.data_ptr():
            self.static_images.copy_(images, non_blocking=True)
        if labels.data_ptr() != self.static_labels.data_ptr():
            self.static_labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        return self.outputs

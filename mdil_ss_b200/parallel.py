"""Data-parallel plumbing: one process per GPU, replicas hold identical weights, each rank runs the whole
hot path on its own batch shard, and the ONLY data-path collective is one all-reduce over a single flat fp32
gradient buffer per optimiser step (replaces the reference's nn.DataParallel replicate / scatter / gather /
reduce, train_new_task_step2.py:473-475).

The flat buffers ARE the storage of the parameters' ``.data`` / ``.grad`` (views), so there is no pack / unpack
around the collective and the fused multi-tensor Adam (mdil_adam_step) walks the same memory.
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib as L


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class FlatBuffer:
    """Flatten ``params`` into one contiguous fp32 buffer (params become views) with a matching gradient buffer."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("FlatBuffer: no parameters")
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        total = self.offsets[-1]
        self.data = torch.empty(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o, n in zip(self.params, self.offsets, self.sizes):
                self.data[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.data[o:o + n].view(p.shape)
                p.grad = self.grad[o:o + n].view(p.shape)

    def numel(self) -> int:
        return self.offsets[-1]

    def zero_grad(self) -> None:
        self.grad.zero_()
        for p, o, n in zip(self.params, self.offsets, self.sizes):  # re-attach if something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + n].view(p.shape)


class GradAllReduce:
    """One all-reduce(sum) over the flat gradient buffer(s), then 1/world scaling folded into the optimiser."""

    def __init__(self, buffers: Sequence[FlatBuffer], group=None):
        self.buffers = list(buffers)
        self.group = group
        # the buffers of all groups live in ONE allocation so that a single collective covers them
        total = sum(b.numel() for b in self.buffers)
        dev = self.buffers[0].grad.device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for b in self.buffers:
            n = b.numel()
            b.grad = self.flat[off:off + n]
            for p, o, sz in zip(b.params, b.offsets, b.sizes):
                p.grad = b.grad[o:o + sz].view(p.shape)
            off += n
        self.calls = 0

    def zero_grad(self) -> None:
        self.flat.zero_()
        for b in self.buffers:
            for p, o, n in zip(b.params, b.offsets, b.sizes):
                if p.grad is None or p.grad.data_ptr() != b.grad.data_ptr() + 4 * o:
                    p.grad = b.grad[o:o + n].view(p.shape)
                # the backward kernels may write this step's first gradient straight into the buffer
                # (functional.grad_target): saves one temporary and one accumulation kernel per parameter
                p._mdil_grad_fresh = True

    def allreduce(self) -> float:
        """Returns the factor the optimiser must apply to the summed gradient (1/world)."""
        w = _world(self.group)
        if w > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
        return 1.0 / w


class FlatAdam:
    """torch.optim.Adam as the drivers configure it (train_new_task_step2.py:229-239: lr per group, betas
    (0.9, 0.999), eps 1e-8, L2 weight_decay 1e-4 added to the gradient) as ONE fused launch per parameter group
    over the flat buffers.  On CUDA it calls mdil_adam_step; on CPU tensors (gloo tests) the same update is
    written with torch ops."""

    def __init__(self, groups: Sequence[dict], lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-4):
        self.graph_safe = False      # see enable_graph_safe()
        self.groups = []
        for g in groups:
            params = [p for p in g["params"] if p.requires_grad]
            if not params:
                continue
            buf = FlatBuffer(params)
            self.groups.append({"buf": buf, "lr": g.get("lr", lr), "initial_lr": g.get("lr", lr),
                                "exp_avg": torch.zeros_like(buf.data), "exp_avg_sq": torch.zeros_like(buf.data)})
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.step_count = 0
        self.reducer = GradAllReduce([g["buf"] for g in self.groups])

    def zero_grad(self) -> None:
        self.reducer.zero_grad()

    def enable_graph_safe(self) -> None:
        """Keep the step counter and the learning rate of every group in device memory (mdil_adam_step_dev) so that
        ``step()`` can be captured in a CUDA graph and replayed (train_step.GraphedStep).  Requires that every parameter
        of a group receives a gradient in every step (true for steps 1 and 2, not for step 3's KD step)."""
        if self.graph_safe:
            return
        for g in self.groups:
            buf = g["buf"]
            if not buf.data.is_cuda:
                raise RuntimeError("FlatAdam.enable_graph_safe: CUDA parameters only")
            steps = g.get("steps", [0] * len(buf.params))
            if len(set(steps)) > 1:
                raise RuntimeError("FlatAdam.enable_graph_safe: the parameters of a group have different step counts")
            g["dev_state"] = torch.tensor([float(steps[0]), float(g["lr"]), 0.0, 0.0], device=buf.data.device,
                                          dtype=torch.float32)
        self.graph_safe = True

    def _sync_steps_from_device(self) -> None:
        if self.graph_safe:
            for g in self.groups:
                n = int(round(float(g["dev_state"][0])))
                g["steps"] = [n] * len(g["buf"].params)
            self.step_count = max([g["steps"][0] for g in self.groups] or [0])

    def set_lr_factor(self, factor: float) -> None:
        """LambdaLR semantics (train_new_task_step2.py:244-245): lr = initial_lr * factor."""
        for g in self.groups:
            g["lr"] = g["initial_lr"] * factor
            if self.graph_safe:
                g["dev_state"][1] = g["lr"]        # read by the captured step on its next replay

    def step(self, allreduce: bool = True) -> None:
        scale = self.reducer.allreduce() if allreduce else 1.0
        self.step_count += 1
        b1, b2 = self.betas
        for g in self.groups:
            buf = g["buf"]
            steps = g.setdefault("steps", [0] * len(buf.params))
            if self.graph_safe:
                if any(getattr(p, "_mdil_grad_fresh", False) for p in buf.params):
                    raise RuntimeError("FlatAdam (graph-safe mode): a parameter received no gradient in this step")
                with torch.cuda.device_of(buf.data):
                    L.check(L.lib().mdil_adam_step_dev(buf.data.data_ptr(), buf.grad.data_ptr(), g["exp_avg"].data_ptr(),
                                                       g["exp_avg_sq"].data_ptr(), buf.numel(), g["dev_state"].data_ptr(),
                                                       b1, b2, self.eps, self.weight_decay, scale,
                                                       torch.cuda.current_stream().cuda_stream), "mdil_adam_step_dev")
                continue
            if buf.data.is_cuda:
                # torch.optim.Adam skips parameters whose .grad is None, which is what the drivers' zero_grad() leaves
                # behind for tensors no loss reached (step 3's KD step never touches the new domain's tensors,
                # train_new_task_step3.py:349-353): a parameter whose gradient was not written since zero_grad()
                # (functional.grad_target clears `_mdil_grad_fresh` on the first write) keeps its value, moments and
                # step count.  Contiguous runs of updated parameters with the same step count share one fused launch.
                touched = [not getattr(p, "_mdil_grad_fresh", False) for p in buf.params]
                i, n = 0, len(buf.params)
                with torch.cuda.device_of(buf.data):
                    while i < n:
                        if not touched[i]:
                            i += 1
                            continue
                        j = i
                        while j + 1 < n and touched[j + 1] and steps[j + 1] == steps[i]:
                            j += 1
                        lo, hi = buf.offsets[i], buf.offsets[j + 1]
                        L.check(L.lib().mdil_adam_step(buf.data.data_ptr() + 4 * lo, buf.grad.data_ptr() + 4 * lo,
                                                       g["exp_avg"].data_ptr() + 4 * lo, g["exp_avg_sq"].data_ptr() + 4 * lo,
                                                       hi - lo, g["lr"], b1, b2, self.eps, self.weight_decay, steps[i] + 1,
                                                       scale, torch.cuda.current_stream().cuda_stream), "mdil_adam_step")
                        for k in range(i, j + 1):
                            steps[k] += 1
                        i = j + 1
            else:
                for k in range(len(steps)):
                    steps[k] += 1
                t = steps[0]
                with torch.no_grad():
                    grad = buf.grad * scale + self.weight_decay * buf.data
                    g["exp_avg"].mul_(b1).add_(grad, alpha=1 - b1)
                    g["exp_avg_sq"].mul_(b2).addcmul_(grad, grad, value=1 - b2)
                    bc1 = 1 - b1 ** t
                    bc2 = 1 - b2 ** t
                    denom = (g["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(self.eps)
                    buf.data.addcdiv_(g["exp_avg"], denom, value=-g["lr"] / bc1)
        self._bump_versions()

    # ---- torch.optim.Adam's checkpoint format (the drivers save optimizer.state_dict(), train_new_task_step2.py:380)
    def state_dict(self) -> dict:
        self._sync_steps_from_device()
        state, groups, idx = {}, [], 0
        for g in self.groups:
            buf = g["buf"]
            steps = g.get("steps", [0] * len(buf.params))
            ids = []
            for k, (o, n, p) in enumerate(zip(buf.offsets, buf.sizes, buf.params)):
                if steps[k] > 0:
                    state[idx] = {"step": torch.tensor(float(steps[k])),
                                  "exp_avg": g["exp_avg"][o:o + n].view(p.shape).clone(),
                                  "exp_avg_sq": g["exp_avg_sq"][o:o + n].view(p.shape).clone()}
                ids.append(idx)
                idx += 1
            groups.append({"lr": g["lr"], "initial_lr": g["initial_lr"], "betas": tuple(self.betas), "eps": self.eps,
                           "weight_decay": self.weight_decay, "amsgrad": False, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict) -> None:
        if len(sd["param_groups"]) != len(self.groups):
            raise ValueError("FlatAdam.load_state_dict: parameter group count differs")
        for g, sg in zip(self.groups, sd["param_groups"]):
            buf = g["buf"]
            if len(sg["params"]) != len(buf.params):
                raise ValueError("FlatAdam.load_state_dict: parameter count of a group differs")
            g["lr"] = sg["lr"]
            g["initial_lr"] = sg.get("initial_lr", g["initial_lr"])
            steps = g.setdefault("steps", [0] * len(buf.params))
            with torch.no_grad():
                for k, (o, n, pid) in enumerate(zip(buf.offsets, buf.sizes, sg["params"])):
                    st = sd["state"].get(pid)
                    if st is None:
                        steps[k] = 0
                        g["exp_avg"][o:o + n].zero_()
                        g["exp_avg_sq"][o:o + n].zero_()
                    else:
                        steps[k] = int(float(st["step"]))
                        g["exp_avg"][o:o + n].copy_(st["exp_avg"].reshape(-1))
                        g["exp_avg_sq"][o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
        self.step_count = max([max(g.get("steps", [0]) or [0]) for g in self.groups] or [0])
        if self.graph_safe:
            for g in self.groups:
                steps = g["steps"]
                if len(set(steps)) > 1:
                    raise RuntimeError("FlatAdam (graph-safe mode): loaded state has per-parameter step counts")
                g["dev_state"][0] = float(steps[0])
                g["dev_state"][1] = float(g["lr"])

    def _bump_versions(self) -> None:
        # the parameters changed in place through the flat buffer: bump the version counters that the
        # weight-pack caches (functional.PackedCache) key on
        for g in self.groups:
            for p in g["buf"].params:
                torch.autograd.graph.increment_version(p)

    def trainable_numel(self) -> int:
        return sum(g["buf"].numel() for g in self.groups)


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers."""
    if _world(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)
            # dist.broadcast does not bump the version counter the packed-weight caches key on
            torch.autograd.graph.increment_version(t)
    from .functional import invalidate_packed
    invalidate_packed(module)


def shard_batch(n_global: int, rank: int, world: int) -> slice:
    """Rank r gets crops [r*N/k, (r+1)*N/k)."""
    if n_global % world:
        raise ValueError("global batch must be divisible by the world size")
    per = n_global // world
    return slice(rank * per, (rank + 1) * per)

"""Checkpoint surgery of the incremental drivers, restated as functions (pure host code).

* ``transfer_previous_step``  — train_new_task_step2.py:499-529 / train_new_task_step3.py (same block): start step t from
  the step t-1 checkpoint: every tensor the two models share is taken as is, the domain-(t-1) adapters / BatchNorm affine
  parameters initialise the domain-t ones, decoder t-1 initialises decoder t except its ``output_conv``.
* ``imagenet_encoder_rename`` — train_new_task_step2.py:490-497: the ImageNet-pretrained encoder checkpoint stores its
  tensors under ``module.features.*``.
* ``strip_module_prefix`` / ``add_module_prefix`` — the drivers save ``nn.DataParallel`` state_dicts (``module.`` keys,
  Evaluation_Notebook.ipynb cell 11).

The regular expressions are the reference's own (their ``.`` is a wildcard there too)."""
from __future__ import annotations

import re
from typing import Dict, Mapping

import torch


def strip_module_prefix(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}


def add_module_prefix(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k if k.startswith("module.") else "module." + k): v for k, v in sd.items()}


def imagenet_encoder_rename(saved: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """train_new_task_step2.py:493-497."""
    return {re.sub("module.features", "module", k): v for k, v in saved.items()}


def previous_step_init_dict(saved: Mapping[str, torch.Tensor], model_keys, current_task: int) -> Dict[str, torch.Tensor]:
    """The dictionary the drivers pass to ``model.load_state_dict(..., strict=False)`` (train_new_task_step2.py:501-527)."""
    model_keys = set(model_keys)
    t = current_task
    new = {}
    for k, v in saved.items():
        if k in model_keys:                       # take all the common params as it is
            new[k] = v
    for k, v in saved.items():
        if "encoder" in k:
            if "parallel_conv" in k or "bn" in k:
                if ".{}.weight".format(t - 1) in k:
                    new[re.sub(".{}.weight".format(t - 1), ".{}.weight".format(t), k)] = v
                elif ".{}.bias".format(t - 1) in k:
                    new[re.sub(".{}.bias".format(t - 1), ".{}.bias".format(t), k)] = v
        elif "decoder" in k and "output_conv" not in k:
            new[re.sub("decoder.{}".format(t - 1), "decoder.{}".format(t), k)] = v
    return new


def transfer_previous_step(saved: Mapping[str, torch.Tensor], model: torch.nn.Module, current_task: int):
    """Initialise ``model`` (step t) from the step t-1 ``state_dict`` exactly as the drivers do.  Works for a bare
    ``Net`` and for a ``DataParallel``-style wrapper alike: the checkpoint's ``module.`` prefix is adapted to the
    model's own key style.  Returns ``load_state_dict``'s (missing, unexpected) report."""
    keys = list(model.state_dict().keys())
    wrapped = bool(keys) and keys[0].startswith("module.")
    saved = add_module_prefix(saved) if wrapped else strip_module_prefix(saved)
    new = previous_step_init_dict(saved, keys, current_task)
    # tensors whose shape differs (a decoder head with another class count can never be hit: output_conv is skipped)
    return model.load_state_dict(new, strict=False)

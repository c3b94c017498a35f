"""GPU-side training co-transform: drop-in for the reference's ``MyCoTransform`` (train_new_task_step2.py:48-81) plus
``ToTensor`` / ``ToLabel`` / ``Relabel(255, NUM_CLASSES - 1)`` (transform.py:63-79), for whole batches of uint8 images
already on the device.  At ~370 crops/s per GPU a 4-worker PIL loader cannot feed the step; decoding stays on the host,
everything after it runs in one kernel (csrc/cotransform.cu) and is bit-exact with the reference (Pillow's fixed-point
BILINEAR resampler and its NEAREST index rule are reproduced; tests/golden/cotransform.npz comes from the reference's
own class).

Random draws follow the reference's order per sample -- ``random.random()`` (hflip when < 0.5), then
``random.randint(-2, 2)`` for transX and transY -- on Python's ``random`` module, so seeding it reproduces the
reference's augmentation stream.
"""
from __future__ import annotations

import math
import random
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

_PRECISION_BITS = 32 - 8 - 2      # Pillow Resample.c


def _bilinear_table(in_size: int, out_size: int) -> Tuple[np.ndarray, int]:
    """[out][2 + K] int32: first source index, tap count, Pillow's 8-bit-path coefficients of the BILINEAR filter
    (precompute_coeffs + normalize_coeffs_8bpc: triangle filter, support scaled by max(in/out, 1), 22-bit fixed point)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    tab = np.zeros((out_size, 2 + ksize), np.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        x0 = max(int(center - support + 0.5), 0)
        x1 = min(int(center + support + 0.5), in_size)
        n = x1 - x0
        w = [max(0.0, 1.0 - abs((x + x0 - center + 0.5) * inv)) for x in range(n)]
        tot = sum(w)
        if tot != 0.0:
            w = [v / tot for v in w]
        tab[xx, 0], tab[xx, 1] = x0, n
        for x in range(n):
            tab[xx, 2 + x] = int(w[x] * (1 << _PRECISION_BITS) + 0.5)
    return tab, ksize


def _nearest_table(in_size: int, out_size: int) -> np.ndarray:
    """Pillow's NEAREST resize: source index floor((x + 0.5) * in / out)."""
    idx = np.floor((np.arange(out_size) + 0.5) * (in_size / out_size)).astype(np.int64)
    return np.clip(idx, 0, in_size - 1).astype(np.int32)


class GpuCoTransform:
    """``GpuCoTransform(augment, height, width, num_classes)(images_u8, labels_u8)`` with images [N,Hs,Ws,3] and labels
    [N,Hs,Ws] uint8 CUDA tensors returns (float32 [N,3,H,W] in [0,1], int64 [N,1,H,W]), what the reference's DataLoader
    yields after MyCoTransform and default collation."""

    def __init__(self, augment: bool = True, height: int = 512, width: int = 1024, num_classes: int = 20):
        self.augment, self.height, self.width, self.num_classes = augment, height, width, num_classes
        self._tables = {}

    def _get_tables(self, hs: int, ws: int, device):
        key = (hs, ws, str(device))
        t = self._tables.get(key)
        if t is None:
            xt, kx = _bilinear_table(ws, self.width)
            yt, ky = _bilinear_table(hs, self.height)
            t = (torch.from_numpy(xt).to(device), kx, torch.from_numpy(yt).to(device), ky,
                 torch.from_numpy(_nearest_table(ws, self.width)).to(device),
                 torch.from_numpy(_nearest_table(hs, self.height)).to(device))
            self._tables[key] = t
        return t

    @staticmethod
    def draw_params(n: int) -> np.ndarray:
        """The reference's draws for n samples, in its order (train_new_task_step2.py:58-66)."""
        out = np.zeros((n, 3), np.int32)
        for i in range(n):
            out[i, 0] = 1 if random.random() < 0.5 else 0
            out[i, 1] = random.randint(-2, 2)
            out[i, 2] = random.randint(-2, 2)
        return out

    def __call__(self, images: torch.Tensor, labels: torch.Tensor, params: Optional[Sequence] = None):
        if not images.is_cuda or not labels.is_cuda:
            raise RuntimeError("GpuCoTransform runs on CUDA tensors only (decode on the host, transform on the device)")
        if images.dtype != torch.uint8 or labels.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
            raise RuntimeError("GpuCoTransform: images uint8 [N,Hs,Ws,3], labels uint8 [N,Hs,Ws]")
        n, hs, ws, _ = images.shape
        if tuple(labels.shape) != (n, hs, ws):
            raise RuntimeError("GpuCoTransform: label shape does not match the images")
        images, labels = images.contiguous(), labels.contiguous()
        xt, kx, yt, ky, xn, yn = self._get_tables(hs, ws, images.device)
        par = None
        if self.augment:
            p = self.draw_params(n) if params is None else np.asarray(params, np.int32).reshape(n, 3)
            par = torch.from_numpy(np.ascontiguousarray(p)).to(images.device)
        with torch.cuda.device_of(images):
            out_img = torch.empty((n, 3, self.height, self.width), device=images.device, dtype=torch.float32)
            out_lab = torch.empty((n, 1, self.height, self.width), device=images.device, dtype=torch.int64)
            L.check(L.lib().mdil_cotransform(images.data_ptr(), labels.data_ptr(), n, hs, ws, self.height, self.width,
                                             xt.data_ptr(), kx, yt.data_ptr(), ky, xn.data_ptr(), yn.data_ptr(),
                                             None if par is None else par.data_ptr(), self.num_classes,
                                             out_img.data_ptr(), out_lab.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "mdil_cotransform")
        return out_img, out_lab

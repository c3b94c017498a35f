"""CPU oracle of the reference's training co-transform (TEST INFRASTRUCTURE -- not product code).

Restates ``MyCoTransform.__call__`` (train_new_task_step2.py:48-81) on numpy uint8 arrays: Resize (PIL BILINEAR for the
image, NEAREST for the label: Pillow's fixed-point resampler, restated below from its published algorithm -- Pillow
12.2 is the reference's third-party dependency for this step), random horizontal flip, random translation by -2..2 pixels
(ImageOps.expand with fill 0 / 255, then crop: a NEGATIVE shift exposes a strip that Image.crop pads with 0 for image AND
label -- a reference quirk kept as is), ToTensor (uint8 / 255), ToLabel, Relabel(255 -> C-1) (transform.py:63-79).

Parity pin: tests/golden/make_golden_cotransform.py runs the reference's own MyCoTransform (Pillow + torchvision) on
synthetic images and commits inputs, the random draws and outputs; tests/test_oracle_golden.py checks this file against
them bit for bit.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2      # Pillow Resample.c


def bilinear_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter.
    Returns (xmin [out], count [out], k [out][ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        x0 = int(center - support + 0.5)
        x0 = max(x0, 0)
        x1 = int(center + support + 0.5)
        x1 = min(x1, in_size)
        n = x1 - x0
        w = np.zeros(ksize, np.float64)
        for x in range(n):
            a = abs((x + x0 - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        tot = w.sum()
        if tot != 0.0:
            w[:n] /= tot
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(v - 0.5) if w[x] < 0 else int(v + 0.5)
        xmin[xx], cnt[xx] = x0, n
    return xmin, cnt, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img [H,W,C] uint8 -> [out_h,out_w,C] uint8, horizontal pass then vertical pass with an 8-bit intermediate."""
    h, w, c = img.shape
    cur = img
    if out_w != w:
        xmin, cnt, kk = bilinear_coeffs(w, out_w)
        out = np.zeros((h, out_w, c), np.uint8)
        for xx in range(out_w):
            acc = np.full((h, c), 1 << (PRECISION_BITS - 1), np.int64)
            for x in range(cnt[xx]):
                acc += cur[:, xmin[xx] + x, :].astype(np.int64) * int(kk[xx, x])
            out[:, xx, :] = _clip8(acc)
        cur = out
    if out_h != h:
        ymin, cnt, kk = bilinear_coeffs(h, out_h)
        out = np.zeros((out_h, cur.shape[1], c), np.uint8)
        for yy in range(out_h):
            acc = np.full((cur.shape[1], c), 1 << (PRECISION_BITS - 1), np.int64)
            for y in range(cnt[yy]):
                acc += cur[ymin[yy] + y, :, :].astype(np.int64) * int(kk[yy, y])
            out[yy] = _clip8(acc)
        cur = out
    return cur


def nearest_index(in_size: int, out_size: int) -> np.ndarray:
    """Pillow's NEAREST resize (affine scale transform): source index floor((x + 0.5) * in/out)."""
    idx = np.floor((np.arange(out_size) + 0.5) * (in_size / out_size)).astype(np.int64)
    return np.clip(idx, 0, in_size - 1)


def resize_nearest_u8(lab: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    return lab[nearest_index(lab.shape[0], out_h)][:, nearest_index(lab.shape[1], out_w)]


def translate(arr: np.ndarray, tx: int, ty: int, fill: int) -> np.ndarray:
    """ImageOps.expand(border=(tx, ty, 0, 0), fill) followed by crop((0, 0, W, H)) (train_new_task_step2.py:68-73)."""
    h, w = arr.shape[:2]
    out = np.zeros_like(arr)                       # what Image.crop pads with beyond the expanded image
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    inside = (xx < w + tx) & (yy < h + ty)         # inside the expanded image
    sx, sy = xx - tx, yy - ty
    src_ok = inside & (sx >= 0) & (sy >= 0)
    out[inside & ~src_ok] = fill
    out[src_ok] = arr[sy[src_ok], sx[src_ok]]
    return out


def cotransform(img: np.ndarray, lab: np.ndarray, height: int, width: int, num_classes: int, augment: bool,
                hflip: bool = False, tx: int = 0, ty: int = 0):
    """img [Hs,Ws,3] uint8, lab [Hs,Ws] uint8 -> (float32 [3,H,W] in [0,1], int64 [1,H,W])."""
    im = resize_bilinear_u8(img, height, width)
    lb = resize_nearest_u8(lab, height, width)
    if augment:
        if hflip:
            im, lb = im[:, ::-1], lb[:, ::-1]
        im = translate(np.ascontiguousarray(im), tx, ty, 0)
        lb = translate(np.ascontiguousarray(lb), tx, ty, 255)
    out_img = (im.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)
    out_lab = lb.astype(np.int64)[None]
    out_lab[out_lab == 255] = num_classes - 1
    return np.ascontiguousarray(out_img), out_lab

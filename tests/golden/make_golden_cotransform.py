"""Golden vectors of the training co-transform from the UNMODIFIED reference (train_new_task_step2.py:48-81 MyCoTransform,
Pillow + torchvision), on synthetic images:  python tests/golden/make_golden_cotransform.py

For every case the fixture stores the source image / label, the reference's random draws (hflip, transX, transY: the
`random` module is seeded and the draws are replayed to record them) and the reference outputs."""
import os
import random
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import reference_modules  # noqa: E402

OUT = os.path.join(HERE, "cotransform.npz")
# (source H, W) -> (H, W), classes, seed of the `random` stream
CASES = [((96, 192), (48, 96), 20, 1), ((90, 160), (64, 128), 20, 2), ((54, 96), (64, 128), 27, 3), ((96, 192), (48, 96), 20, 4),
         ((72, 128), (64, 128), 27, 5), ((64, 128), (64, 128), 20, 6), ((96, 192), (48, 96), 20, 7), ((90, 160), (64, 128), 20, 9)]


def main():
    _, step2 = reference_modules()
    rng = np.random.default_rng(0)
    out = {}
    for i, ((hs, ws), (h, w), ncls, seed) in enumerate(CASES):
        # smooth-ish image with sharp edges, labels in blocks with some 255 (ignore) pixels
        img = (rng.integers(0, 256, (hs // 6 + 1, ws // 6 + 1, 3)).repeat(6, 0).repeat(6, 1)[:hs, :ws] * 0.7
               + rng.integers(0, 77, (hs, ws, 3))).astype(np.uint8)
        lab = rng.integers(0, ncls - 1, (hs // 8 + 1, ws // 8 + 1)).repeat(8, 0).repeat(8, 1)[:hs, :ws].astype(np.uint8)
        lab[rng.random((hs, ws)) < 0.03] = 255
        step2.NUM_CLASSES = ncls
        co = step2.MyCoTransform(augment=True, height=h, width=w)
        random.seed(seed)
        hflip = random.random() < 0.5
        tx, ty = random.randint(-2, 2), random.randint(-2, 2)
        random.seed(seed)
        x, y = co(Image.fromarray(img), Image.fromarray(lab))
        out[f"img{i}"], out[f"lab{i}"] = img, lab
        out[f"par{i}"] = np.array([int(hflip), tx, ty, h, w, ncls], np.int32)
        out[f"x{i}"], out[f"y{i}"] = x.numpy(), y.numpy()
    # validation-style call (augment off)
    co = step2.MyCoTransform(augment=False, height=48, width=96)
    step2.NUM_CLASSES = 20
    x, y = co(Image.fromarray(out["img0"]), Image.fromarray(out["lab0"]))
    out["x_noaug"], out["y_noaug"] = x.numpy(), y.numpy()
    out["n"] = np.array(len(CASES))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes", [tuple(out[f"par{i}"][:3]) for i in range(len(CASES))])


if __name__ == "__main__":
    main()

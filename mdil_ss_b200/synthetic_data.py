"""Synthetic dataset trees in the directory and file-name conventions the reference's loaders walk, so the
UNMODIFIED drivers can be pointed at data on a box that has none:

    cityscapes  <root>/leftImg8bit/<subset>/<city>/<id>_leftImg8bit.png
                <root>/gtFine/<subset>/<city>/<id>_gtFine_labelTrainIds.png          dataset.py:75-92   (19 classes + 255)
    IDD         <root>/leftImg8bit/<subset>/<seq>/<id>_leftImg8bit.png
                <root>/gtFine/<subset>/<seq>/<id>_gtFine_labellevel3Ids.png          dataset.py:118-140 (26 classes + 255)
    BDD         <root>/images/<subset>/<id>.jpg
                <root>/labels/<subset>/<id>_train_id.png                             dataset.py:222-240 (19 classes + 255)

The loaders sort the image list and the label list independently and pair them by index (dataset.py:84-92), so ids are
zero-padded and shared between the two files.  Labels are piecewise-constant blocks with about 10 % of the pixels set to
255, which the drivers' co-transform relabels to NUM_CLASSES - 1, the zero-weight class (train_new_task_step2.py:79,
133-135).  Every image encodes its own label map (``label_colour``), so a test can check that image i and label i
belong together after the loaders' sort.

Host-side test/bring-up infrastructure only: nothing on the hot path imports this module.
"""
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

KINDS = {
    # kind: (image dir, label dir, label suffix, image suffix, nested?, number of real classes)
    "cityscapes": ("leftImg8bit", "gtFine", "_gtFine_labelTrainIds.png", "_leftImg8bit.png", True, 19),
    "IDD": ("leftImg8bit", "gtFine", "_gtFine_labellevel3Ids.png", "_leftImg8bit.png", True, 26),
    "BDD": ("images", "labels", "_train_id.png", ".jpg", False, 19),
}
# the roots the drivers hard-code (train_new_task_step2.py:140-142); create the trees there, or bind-mount them
REFERENCE_ROOTS = {
    "cityscapes": "/ssd_scratch/cvit/prachigarg/cityscapes/",
    "BDD": "/ssd_scratch/cvit/prachigarg/bdd100k/seg/",
    "IDD": "/ssd_scratch/cvit/prachigarg/IDD_Segmentation/",
}
IGNORE = 255


def label_colour(label: np.ndarray) -> np.ndarray:
    """RGB value an image carries where its label map says ``label`` (before noise): class k -> (8k+4, 255-8k, 64+4k),
    ignore -> black."""
    k = label.astype(np.int32)
    rgb = np.stack([8 * k + 4, 255 - 8 * k, 64 + 4 * k], axis=-1)
    rgb[label == IGNORE] = 0
    return np.clip(rgb, 0, 255).astype(np.uint8)


def make_label_map(rng: np.random.Generator, height: int, width: int, num_classes: int, block: int = 32,
                   ignore_fraction: float = 0.10) -> np.ndarray:
    bh, bw = -(-height // block), -(-width // block)
    coarse = rng.integers(0, num_classes, size=(bh, bw), dtype=np.int64).astype(np.uint8)
    coarse[rng.random((bh, bw)) < ignore_fraction] = IGNORE
    if ignore_fraction > 0 and not (coarse == IGNORE).any():       # every map exercises the ignore path
        coarse[rng.integers(0, bh), rng.integers(0, bw)] = IGNORE
    return np.repeat(np.repeat(coarse, block, axis=0), block, axis=1)[:height, :width].copy()


def write_dataset_tree(root: str, kind: str, subsets: Sequence[str] = ("train", "val"), per_subset: int = 4,
                       size: Tuple[int, int] = (256, 512), seed: int = 0, noise: int = 3) -> Dict[str, List[Tuple[str, str]]]:
    """Writes ``per_subset`` image/label pairs per subset under ``root`` and returns ``{subset: [(image path, label
    path), ...]}`` in the order the reference's loaders will see them."""
    from PIL import Image
    if kind not in KINDS:
        raise ValueError(f"unknown dataset kind {kind!r} (expected one of {sorted(KINDS)})")
    img_dir, lab_dir, lab_suffix, img_suffix, nested, ncls = KINDS[kind]
    height, width = size
    rng = np.random.default_rng(seed)
    out: Dict[str, List[Tuple[str, str]]] = {}
    for subset in subsets:
        pairs = []
        for i in range(per_subset):
            group = f"{'city' if kind == 'cityscapes' else 'seq'}{i % 2:02d}" if nested else ""
            ident = f"{group + '_' if group else ''}{i:06d}"
            idir = os.path.join(root, img_dir, subset, group)
            ldir = os.path.join(root, lab_dir, subset, group)
            os.makedirs(idir, exist_ok=True)
            os.makedirs(ldir, exist_ok=True)
            label = make_label_map(rng, height, width, ncls)
            image = label_colour(label).astype(np.int16)
            if noise and img_suffix.endswith(".png"):
                image += rng.integers(-noise, noise + 1, size=image.shape, dtype=np.int16)
            ipath = os.path.join(idir, ident + img_suffix)
            lpath = os.path.join(ldir, ident + lab_suffix)
            im = Image.fromarray(np.clip(image, 0, 255).astype(np.uint8), "RGB")
            if img_suffix == ".jpg":
                im.save(ipath, quality=95)
            else:
                im.save(ipath)
            Image.fromarray(label, "L").save(lpath)
            pairs.append((ipath, lpath))
        pairs.sort()
        out[subset] = pairs
    return out


def write_driver_stubs(directory: str) -> List[str]:
    """The two modules the drivers import and never use (train_new_task_step2.py:28,37), absent from the reference."""
    os.makedirs(directory, exist_ok=True)
    files = {"config_task.py": "# imported by the drivers, never used\n",
             "torchsummary.py": "def summary(*args, **kwargs):\n    return None\n"}
    written = []
    for name, text in files.items():
        path = os.path.join(directory, name)
        with open(path, "w") as f:
            f.write(text)
        written.append(path)
    return written

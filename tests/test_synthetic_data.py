"""CPU: the synthetic dataset trees (mdil_ss_b200/synthetic_data.py) follow the reference loaders' file
conventions; when the reference is present (build container) its own dataset classes and co-transform read them."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import REFERENCE, REPO

from mdil_ss_b200 import synthetic_data as sd


@pytest.mark.parametrize("kind", ["cityscapes", "IDD", "BDD"])
def test_tree_layout_and_contents(tmp_path, kind):
    from PIL import Image
    pairs = sd.write_dataset_tree(str(tmp_path), kind, per_subset=3, size=(128, 256), seed=5)
    assert sorted(pairs) == ["train", "val"]
    img_dir, lab_dir, lab_suffix, img_suffix, nested, ncls = sd.KINDS[kind]
    for subset, items in pairs.items():
        assert len(items) == 3
        # the loaders sort images and labels independently and pair by index (dataset.py:84-92)
        assert [p[0] for p in items] == sorted(p[0] for p in items)
        assert [p[1] for p in items] == sorted(p[1] for p in items)
        for ipath, lpath in items:
            assert ipath.startswith(os.path.join(str(tmp_path), img_dir, subset)) and ipath.endswith(img_suffix)
            assert lpath.startswith(os.path.join(str(tmp_path), lab_dir, subset)) and lpath.endswith(lab_suffix)
            assert (os.path.dirname(ipath) != os.path.join(str(tmp_path), img_dir, subset)) == nested
            lab = np.array(Image.open(lpath))
            img = np.array(Image.open(ipath).convert("RGB")).astype(np.int32)
            assert lab.dtype == np.uint8 and lab.shape == (128, 256) and img.shape == (128, 256, 3)
            vals = set(np.unique(lab).tolist())
            assert vals <= set(range(ncls)) | {sd.IGNORE} and len(vals) > 3
            assert 0.0 < float((lab == sd.IGNORE).mean()) < 0.4
            err = np.abs(img - sd.label_colour(lab).astype(np.int32))
            assert float(err.mean()) < (6.0 if img_suffix == ".jpg" else 3.5)
    # deterministic under the seed
    again = sd.write_dataset_tree(str(tmp_path / "again"), kind, per_subset=3, size=(128, 256), seed=5)
    a = np.array(Image.open(pairs["train"][0][1]))
    b = np.array(Image.open(again["train"][0][1]))
    assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        sd.write_dataset_tree(str(tmp_path), "VOC")


def test_driver_stubs_import(tmp_path):
    files = sd.write_driver_stubs(str(tmp_path))
    assert sorted(os.path.basename(f) for f in files) == ["config_task.py", "torchsummary.py"]
    ns = {}
    exec(open(os.path.join(str(tmp_path), "torchsummary.py")).read(), ns)
    assert ns["summary"](1, 2, x=3) is None


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("kind,cls_name,nclass", [("cityscapes", "cityscapes", 20), ("IDD", "IDD", 27), ("BDD", "BDD100k", 20)])
def test_reference_loaders_read_the_tree(tmp_path, kind, cls_name, nclass):
    """The reference's own dataset class + MyCoTransform (train_new_task_step2.py:48-81) over a synthetic tree: tensor
    contract of the hot path's inputs, 255 -> NUM_CLASSES-1 relabelling, and image/label pairing."""
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    from make_golden import load_by_path, reference_modules
    _, step2 = reference_modules()
    ref_dataset = load_by_path("ref_dataset", os.path.join(REFERENCE, "dataset.py"))
    sd.write_dataset_tree(str(tmp_path), kind, per_subset=3, size=(128, 256), seed=9)
    step2.NUM_CLASSES = nclass
    ds = getattr(ref_dataset, cls_name)(str(tmp_path), step2.MyCoTransform(False, 64, 128), "val")
    assert len(ds) == 3
    for i in range(3):
        image, label = ds[i]
        assert image.dtype == torch.float32 and tuple(image.shape) == (3, 64, 128)
        assert 0.0 <= float(image.min()) and float(image.max()) <= 1.0
        assert label.dtype == torch.int64 and tuple(label.shape) == (1, 64, 128)
        assert int(label.max()) <= nclass - 1 and int((label == 255).sum()) == 0
        assert int((label == nclass - 1).sum()) > 0                     # the relabelled ignore pixels
        # pairing: at block centres (32-pixel blocks halved by the resize) the image carries its label's colour
        lab = label[0, 8::16, 8::16].numpy()
        pix = (image[:, 8::16, 8::16].permute(1, 2, 0).numpy() * 255.0)
        src = np.where(lab == nclass - 1, sd.IGNORE, lab).astype(np.uint8)
        assert float(np.abs(pix - sd.label_colour(src)).mean()) < 8.0
    # the augmenting transform (flip, +-2 pixel translation with 255 padding) keeps the contract
    aug = getattr(ref_dataset, cls_name)(str(tmp_path), step2.MyCoTransform(True, 64, 128), "train")
    image, label = aug[0]
    assert tuple(image.shape) == (3, 64, 128) and tuple(label.shape) == (1, 64, 128) and int(label.max()) <= nclass - 1

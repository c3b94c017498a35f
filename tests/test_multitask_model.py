"""The multi-task joint model drop-in (mdil_ss_b200/erfnet_multi_task.py for the reference's
models/erfnet_multi_task.py): constructor contract and oracle restatement on CPU against fixtures produced by the
unmodified reference (tests/golden/make_golden_multitask_model.py); forward/backward parity on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, assert_close, golden, noise_list, oracle

TOL = 1e-3
CLASSES = [20, 20, 27]


def _seeded_sd(seed, bn_seed):
    """The fixture's weights: this repo's constructor under the reference's seed (bit-identical initial state_dict,
    asserted by the generating script) + the deterministic BatchNorm perturbation."""
    from mdil_ss_b200.erfnet_multi_task import Net
    torch.manual_seed(seed)
    net = Net(CLASSES, 3)
    sd = oracle.perturb_bn_(oracle.clone_sd(net.state_dict()), seed=bn_seed)
    net.load_state_dict(sd, strict=True)
    return net, sd


def test_constructor_contract(capsys):
    from mdil_ss_b200.erfnet_multi_task import Net
    c = json.load(open(os.path.join(GOLDEN, "mt_contract.json")))
    torch.manual_seed(0)
    net = Net(CLASSES, 3)
    assert "hi, inside erfnet_multi_task.py" in capsys.readouterr().out
    sd = net.state_dict()
    assert list(sd.keys()) == c["keys"] and len(c["keys"]) == 519
    assert [list(v.shape) for v in sd.values()] == c["shapes"]
    assert [n for n, _ in net.named_parameters()] == c["params"]
    cs = float(sum(v.double().sum() for v in sd.values() if v.dtype.is_floating_point))
    assert abs(cs - c["checksum"]) <= 1e-9 * max(1.0, abs(c["checksum"]))
    import models.erfnet_multi_task as shim
    assert shim.Net is Net


def _oracle_train(g, sd):
    gen = torch.Generator().manual_seed(int(g["train_x_seed"]))
    x = torch.rand(2, 3, 32, 64, generator=gen)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=gen)
    noise = noise_list(g, "noise_")
    task = int(g["train_task"])
    rap = oracle.multitask_sd_as_rap(sd, 3)
    names = [n for n in oracle.param_names(rap)]
    work = oracle._with_grad(rap, names)
    logits = oracle.net_forward(work, x, task, True, noise)
    loss = oracle.cross_entropy2d(logits, labels[:, 0], torch.tensor(oracle.WEIGHT_BDD))
    grads = dict(zip(names, torch.autograd.grad(loss, [work[n] for n in names], allow_unused=True)))
    back = {}
    for n, gr in grads.items():
        if gr is None or "parallel_conv" in n:
            continue
        m = n
        for src, dst in ((f".bns_1.{task}.", ".bn1."), (f".bns_2.{task}.", ".bn2."), (f".bn_ini.{task}.", ".bn.")):
            m = m.replace(src, dst)
        if ".bns_" in m or ".bn_ini." in m:
            continue
        back[m] = gr
    return x, labels, noise, task, logits.detach(), loss.detach(), back


def test_oracle_restatement_matches_reference_fixture():
    """oracle.multitask_sd_as_rap: the multi-task model evaluated through the RAP restatement reproduces the
    reference module's eval logits, train logits, loss and gradients."""
    g = golden("mt_model.npz")
    _, sd = _seeded_sd(int(g["eval_seed"]), int(g["eval_bn_seed"]))
    x = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(int(g["eval_x_seed"])))
    with torch.no_grad():
        y = oracle.net_forward(oracle.multitask_sd_as_rap(sd, 3), x, int(g["eval_task"]), False)
    assert_close(y, torch.from_numpy(g["eval_logits"]), 2e-5, "eval logits")
    _, sd = _seeded_sd(int(g["train_seed"]), int(g["train_bn_seed"]))
    _, _, _, _, logits, loss, grads = _oracle_train(g, sd)
    assert_close(logits, torch.from_numpy(g["train_logits"]), 2e-5, "train logits")
    assert abs(float(loss) - float(g["train_loss"])) <= 1e-5 * abs(float(g["train_loss"]))
    names = [str(s) for s in g["grad_names"]]
    assert sorted(names) == sorted(grads.keys())
    gabs = np.array([float(grads[n].double().abs().sum()) for n in names])
    np.testing.assert_allclose(gabs, g["grad_abs"], rtol=5e-4, atol=2e-5)
    bn_sum = np.array([float(sd[str(k)].double().sum()) for k in g["bn_names"]])
    np.testing.assert_allclose(bn_sum, g["bn_sum"], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_gpu_eval_logits_match_reference():
    g = golden("mt_model.npz")
    net, _ = _seeded_sd(int(g["eval_seed"]), int(g["eval_bn_seed"]))
    net = net.to("cuda").eval()
    x = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(int(g["eval_x_seed"])))
    with torch.no_grad():
        y = net(x.to("cuda"), int(g["eval_task"]))
    ref = torch.from_numpy(g["eval_logits"])
    assert tuple(y.shape) == tuple(ref.shape)
    assert_close(y, ref, TOL, "eval logits")


@pytest.mark.gpu
def test_gpu_train_forward_backward_matches_reference():
    from mdil_ss_b200.losses import CrossEntropyLoss2d
    g = golden("mt_model.npz")
    net, _ = _seeded_sd(int(g["train_seed"]), int(g["train_bn_seed"]))
    net = net.to("cuda").train()
    counted = {k: int(v) for k, v in net.state_dict().items() if k.endswith("num_batches_tracked")}
    gen = torch.Generator().manual_seed(int(g["train_x_seed"]))
    x = torch.rand(2, 3, 32, 64, generator=gen)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=gen)
    noise = [None if t is None else t.to("cuda") for t in noise_list(g, "noise_")]
    task = int(g["train_task"])
    logits = net(x.to("cuda"), task, drop_noise=noise)
    loss = CrossEntropyLoss2d(torch.tensor(oracle.WEIGHT_BDD)).to("cuda")(logits, labels[:, 0].to("cuda"))
    loss.backward()
    assert_close(logits, torch.from_numpy(g["train_logits"]), TOL, "train logits")
    assert abs(float(loss) - float(g["train_loss"])) <= TOL * abs(float(g["train_loss"]))
    grads = {n: p.grad for n, p in net.named_parameters() if p.grad is not None}
    names = [str(s) for s in g["grad_names"]]
    assert list(grads.keys()) == names
    # statistical gradient check (ReLU near-ties, see test_gpu_net.test_train_forward_backward_matches_reference)
    gabs = np.array([float(grads[n].double().abs().sum()) for n in names])
    np.testing.assert_allclose(gabs, g["grad_abs"], rtol=3e-2, atol=1e-4)
    # per-tensor: relative L2 <= 6e-2 on this random-init 32 x 64 fixture.  The first convolution's gradient collects
    # every ReLU near-tie flip of the network above it (measured 2.9e-2 here; the CPU oracle moves by 6e-3 against itself
    # with another thread count).  The sharp network-level gradient check (5e-3 at 512 x 1024) runs on the trained weights:
    # tests/test_gpu_net.py::test_pretrained_train_step_matches_oracle; the strict 1e-3 per-element checks are the block tests
    for i, n in enumerate(str(s) for s in g["pick"]):
        ref = torch.from_numpy(g[f"grad_{i}"]).double()
        l2 = float((grads[n].double().cpu() - ref).norm() / ref.norm())
        assert l2 <= 6e-2, f"{n}: gradient relative L2 {l2:.2e}"
    after = net.state_dict()
    bn_sum = np.array([float(after[str(k)].double().sum()) for k in g["bn_names"]])
    np.testing.assert_allclose(bn_sum, g["bn_sum"], rtol=1e-4, atol=1e-5)
    # every BatchNorm the step ran through counted one batch (the forward kernels update the buffer), the others none
    counts = {k: int(v) - counted[k] for k, v in after.items() if k.endswith("num_batches_tracked")}
    assert counts["encoder.initial_block.bn.num_batches_tracked"] == 1 and set(counts.values()) <= {0, 1}
    assert sum(counts.values()) >= 30
    assert net.decoder[0].output_conv.weight.grad is None and net.decoder[2].output_conv.weight.grad is None

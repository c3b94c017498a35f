"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE (imported from
/root/reference, CPU, fp32) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The fixtures pin oracle/erfnet_rap_oracle.py (tests/test_oracle_golden.py, CPU) and are the targets of the GPU
parity tests (tests/test_gpu_*.py).  Nothing at test time needs /root/reference.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("MDIL_REFERENCE", "/root/reference")


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_modules():
    """The reference's model file and its driver (for CrossEntropyLoss2d), imported without touching this repo's
    `models` package.  config_task / torchsummary are imported-but-unused by the driver and absent: stub them."""
    for stub in ("config_task", "torchsummary"):
        m = types.ModuleType(stub)
        m.summary = lambda *a, **k: None
        sys.modules.setdefault(stub, m)
    saved = list(sys.path)
    sys.path[:] = [REF] + [p for p in saved if os.path.abspath(p or ".") != REPO]
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    try:
        ref_model = load_by_path("ref_erfnet_RA_parallel", os.path.join(REF, "models", "erfnet_RA_parallel.py"))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_step2 = load_by_path("ref_train_new_task_step2", os.path.join(REF, "train_new_task_step2.py"))
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.path[:] = saved
    return ref_model, ref_step2


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    oracle = load_by_path("erfnet_rap_oracle", os.path.join(REPO, "oracle", "erfnet_rap_oracle.py"))
    ref_model, ref_step2 = reference_modules()
    import warnings
    warnings.simplefilter("ignore")
    out = {}

    # ------------------------------------------------------------------ 1. state_dict / parameter contract
    contract = {}
    for classes in ([20], [20, 20], [20, 20, 27]):
        torch.manual_seed(0)
        net = ref_model.Net(classes, len(classes), len(classes) - 1)
        sd = net.state_dict()
        contract[str(len(classes))] = {
            "keys": list(sd.keys()),
            "shapes": [list(v.shape) for v in sd.values()],
            "params": [n for n, _ in net.named_parameters()],
            "checksum": float(sum(v.double().sum() for v in sd.values() if v.dtype.is_floating_point)),
            "repr_len": len(str(net)),
        }
    with open(os.path.join(HERE, "contract.json"), "w") as f:
        json.dump(contract, f)

    # ------------------------------------------------------------------ 2. eval forward, 1 task and 3 tasks
    for tag, classes, task, seed in (("eval_1task", [20], 0, 0), ("eval_3task_t2", [20, 20, 27], 2, 3),
                                     ("eval_3task_t0", [20, 20, 27], 0, 3)):
        torch.manual_seed(seed)
        net = ref_model.Net(classes, len(classes), len(classes) - 1)
        sd = oracle.perturb_bn_(oracle.clone_sd(net.state_dict()), seed=7 + seed)
        net.load_state_dict(sd)
        net.eval()
        x = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(100 + seed))
        with torch.no_grad():
            y = net(x, task)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), logits=y.numpy(), classes=np.array(classes),
                            task=task, seed=seed, bn_seed=7 + seed, x_seed=100 + seed)

    # ------------------------------------------------------------------ 3. train forward/backward with CE, 2 tasks
    classes, task, seed = [20, 20], 1, 5
    torch.manual_seed(seed)
    net = ref_model.Net(classes, 2, 1)
    sd0 = oracle.perturb_bn_(oracle.clone_sd(net.state_dict()), seed=11)
    net.load_state_dict(sd0)
    net.train()
    g = torch.Generator().manual_seed(200)
    x = torch.rand(2, 3, 32, 64, generator=g)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=g)
    weight = torch.tensor(oracle.WEIGHT_BDD)
    torch.manual_seed(77)
    noise = oracle.make_dropout_noise(2, True)
    torch.manual_seed(77)  # the reference now draws the same Dropout2d noise in the same order
    crit = ref_step2.CrossEntropyLoss2d(weight)
    logits = net(x, task)
    loss = crit(logits, labels[:, 0])
    loss.backward()
    grads = {n: p.grad for n, p in net.named_parameters() if p.grad is not None}
    sd1 = net.state_dict()
    pick = ["encoder.initial_block.conv.weight", "encoder.layers.1.conv3x1_1.weight", "encoder.layers.1.parallel_conv_1.1.weight",
            "encoder.layers.10.conv1x3_2.weight", "encoder.layers.10.bns_2.1.weight", "encoder.layers.14.parallel_conv_2.1.bias",
            "encoder.layers.6.conv.weight", "decoder.1.layers.0.conv.weight", "decoder.1.layers.4.conv3x1_1.weight",
            "decoder.1.output_conv.weight", "decoder.1.output_conv.bias", "decoder.1.layers.3.bn.weight"]
    save = {"logits": logits.detach().numpy(), "loss": float(loss), "x_seed": 200, "init_seed": seed, "bn_seed": 11,
            "noise_seed": 77, "grad_names": np.array(list(grads.keys())),
            "grad_sum": np.array([float(v.double().sum()) for v in grads.values()]),
            "grad_abs": np.array([float(v.double().abs().sum()) for v in grads.values()]),
            "bn_names": np.array([k for k in sd1 if "running" in k]),
            "bn_sum": np.array([float(sd1[k].double().sum()) for k in sd1 if "running" in k])}
    for i, n in enumerate(pick):
        save[f"grad_{i}"] = grads[n].numpy()
    save["pick"] = np.array(pick)
    for i, t in enumerate(noise):
        if t is not None:
            save[f"noise_{i}"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "train_2task_t1.npz"), **save)

    # ------------------------------------------------------------------ 4. losses
    g = torch.Generator().manual_seed(300)
    lg20 = (torch.randn(2, 20, 16, 32, generator=g) * 3).requires_grad_(True)
    lb20 = torch.randint(0, 20, (2, 16, 32), generator=g)
    lb20[torch.rand(2, 16, 32, generator=g) < 0.1] = 19
    lg27 = (torch.randn(2, 27, 16, 32, generator=g) * 3).requires_grad_(True)
    lb27 = torch.randint(0, 27, (2, 16, 32), generator=g)
    lb27[torch.rand(2, 16, 32, generator=g) < 0.1] = 26
    ce20 = ref_step2.CrossEntropyLoss2d(torch.tensor(oracle.WEIGHT_CITY))(lg20, lb20)
    ce20.backward()
    ce27 = ref_step2.CrossEntropyLoss2d(torch.tensor(oracle.WEIGHT_IDD))(lg27, lb27)
    ce27.backward()
    st = (torch.randn(2, 20, 16, 32, generator=g) * 3).requires_grad_(True)
    te = torch.randn(2, 20, 16, 32, generator=g) * 3
    kd = torch.nn.KLDivLoss()(torch.nn.functional.softmax(st, dim=1), torch.nn.functional.softmax(te, dim=1))
    kd.backward()
    np.savez_compressed(os.path.join(HERE, "losses.npz"), lg20=lg20.detach().numpy(), lb20=lb20.numpy(), ce20=float(ce20),
                        dlg20=lg20.grad.numpy(), lg27=lg27.detach().numpy(), lb27=lb27.numpy(), ce27=float(ce27),
                        dlg27=lg27.grad.numpy(), st=st.detach().numpy(), te=te.numpy(), kd=float(kd), dst=st.grad.numpy())

    # ------------------------------------------------------------------ 5. one restated step-2 iteration (CS -> BDD)
    torch.manual_seed(9)
    teacher = ref_model.Net([20], 1, 0)
    sd_old = oracle.perturb_bn_(oracle.clone_sd(teacher.state_dict()), seed=13)
    teacher.load_state_dict(sd_old)
    torch.manual_seed(10)
    student = ref_model.Net([20, 20], 2, 1)
    sd_new = oracle.perturb_bn_(oracle.clone_sd(student.state_dict()), seed=14)
    student.load_state_dict(sd_new)
    for p in teacher.parameters():
        p.requires_grad = False
    for name, m in student.named_parameters():  # train_new_task_step2.py:205-215 with current_task = 1
        if "decoder" in name:
            if "decoder.1" not in name:
                m.requires_grad = False
        elif "encoder" in name and ("bn" in name or "parallel_conv" in name):
            if not (".1.weight" in name or ".1.bias" in name):
                m.requires_grad = False
    ref_step2.current_task = 1
    params = list(student.named_parameters())
    opt = torch.optim.Adam([{"params": [p for n, p in params if ref_step2.is_shared(n)], "lr": 5e-6},
                            {"params": [p for n, p in params if ref_step2.is_DS_curr(n)]}],
                           5e-4, (0.9, 0.999), eps=1e-08, weight_decay=1e-4)
    g = torch.Generator().manual_seed(400)
    x = torch.rand(2, 3, 32, 64, generator=g)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=g)
    weight = torch.tensor(oracle.WEIGHT_BDD)
    crit = ref_step2.CrossEntropyLoss2d(weight)
    kl = torch.nn.KLDivLoss()
    student.train()
    teacher.eval()
    torch.manual_seed(88)
    noise_t = oracle.make_dropout_noise(2, True)
    noise_prev = oracle.make_dropout_noise(2, True)
    torch.manual_seed(88)
    out_t = student(x, 1)
    out_prev = student(x, 0)
    out_old = teacher(x, 0)
    ce = crit(out_t, labels[:, 0])
    kld = kl(torch.nn.functional.softmax(out_prev, dim=1), torch.nn.functional.softmax(out_old, dim=1))
    total = ce + 0.1 * kld
    opt.zero_grad()
    total.backward()
    gsum = {n: float(p.grad.double().sum()) for n, p in student.named_parameters() if p.grad is not None}
    gabs = {n: float(p.grad.double().abs().sum()) for n, p in student.named_parameters() if p.grad is not None}
    opt.step()
    sd_after = student.state_dict()
    save = {"ce": float(ce), "kd": float(kld), "total": float(total), "out_t": out_t.detach().numpy(),
            "grad_names": np.array(list(gsum.keys())), "grad_sum": np.array(list(gsum.values())),
            "grad_abs": np.array(list(gabs.values())),
            "after_names": np.array(list(sd_after.keys())),
            "after_sum": np.array([float(v.double().sum()) for v in sd_after.values()]),
            "delta_abs": np.array([float((sd_after[k].double() - sd_new[k].double()).abs().sum()) for k in sd_after])}
    for i, t in enumerate(noise_t):
        if t is not None:
            save[f"noise_t_{i}"] = t.numpy()
    for i, t in enumerate(noise_prev):
        if t is not None:
            save[f"noise_prev_{i}"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "step2_iter.npz"), **save)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f"  {f:28s} {os.path.getsize(os.path.join(HERE, f)) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()

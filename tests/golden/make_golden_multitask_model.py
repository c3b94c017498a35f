"""Golden fixtures for the multi-task joint model (models/erfnet_multi_task.py, SURVEY 8a note / 8f-4), produced by
RUNNING THE UNMODIFIED REFERENCE module on CPU:

    python tests/golden/make_golden_multitask_model.py

* mt_contract.json — state_dict keys / shapes / parameter order / checksum of Net([20, 20, 27], 3) under seed 0;
* mt_model.npz     — eval logits (task 2) and one train forward/backward with the reference's CrossEntropyLoss2d (task 1,
                     Dropout2d noise recorded) on seeded weights.
The script also asserts that this repo's constructor (mdil_ss_b200/erfnet_multi_task.py, parameter containers only, CPU)
reproduces the reference's initial state_dict bit-exactly under the same seeds, which is how the GPU tests rebuild the
fixture's weights without the reference.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, REPO, load_by_path, reference_modules  # noqa: E402

CLASSES = [20, 20, 27]


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    oracle = load_by_path("erfnet_rap_oracle", os.path.join(REPO, "oracle", "erfnet_rap_oracle.py"))
    _, ref_step2 = reference_modules()
    ref_mt = load_by_path("ref_erfnet_multi_task", os.path.join(REF, "models", "erfnet_multi_task.py"))
    sys.path.insert(0, REPO)
    from mdil_ss_b200 import erfnet_multi_task as ours

    def same_init(seed):
        torch.manual_seed(seed)
        a = ref_mt.Net(CLASSES, 3).state_dict()
        torch.manual_seed(seed)
        b = ours.Net(CLASSES, 3).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a), "constructor does not reproduce the reference's init"

    # ---- contract
    torch.manual_seed(0)
    net = ref_mt.Net(CLASSES, 3)
    sd = net.state_dict()
    contract = {"keys": list(sd.keys()), "shapes": [list(v.shape) for v in sd.values()],
                "params": [n for n, _ in net.named_parameters()],
                "checksum": float(sum(v.double().sum() for v in sd.values() if v.dtype.is_floating_point))}
    with open(os.path.join(HERE, "mt_contract.json"), "w") as f:
        json.dump(contract, f)
    for seed in (0, 41, 43):
        same_init(seed)

    save = {"classes": np.array(CLASSES)}
    # ---- eval forward, task 2
    torch.manual_seed(41)
    net = ref_mt.Net(CLASSES, 3)
    net.load_state_dict(oracle.perturb_bn_(oracle.clone_sd(net.state_dict()), seed=42))
    net.eval()
    x = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(140))
    with torch.no_grad():
        save["eval_logits"] = net(x, 2).numpy()
    save.update(eval_seed=41, eval_bn_seed=42, eval_x_seed=140, eval_task=2)

    # ---- train forward/backward, task 1
    torch.manual_seed(43)
    net = ref_mt.Net(CLASSES, 3)
    net.load_state_dict(oracle.perturb_bn_(oracle.clone_sd(net.state_dict()), seed=44))
    net.train()
    g = torch.Generator().manual_seed(240)
    x = torch.rand(2, 3, 32, 64, generator=g)
    labels = torch.randint(0, 20, (2, 1, 32, 64), generator=g)
    torch.manual_seed(79)
    noise = oracle.make_dropout_noise(2, True)
    torch.manual_seed(79)          # the reference now draws the same Dropout2d noise in the same order
    logits = net(x, 1)
    loss = ref_step2.CrossEntropyLoss2d(torch.tensor(oracle.WEIGHT_BDD))(logits, labels[:, 0])
    loss.backward()
    grads = {n: p.grad for n, p in net.named_parameters() if p.grad is not None}
    sd1 = net.state_dict()
    save.update(train_seed=43, train_bn_seed=44, train_x_seed=240, train_noise_seed=79, train_task=1,
                train_logits=logits.detach().numpy(), train_loss=float(loss),
                grad_names=np.array(list(grads.keys())),
                grad_abs=np.array([float(v.double().abs().sum()) for v in grads.values()]),
                grad_sq=np.array([float(v.double().pow(2).sum()) for v in grads.values()]),
                bn_names=np.array([k for k in sd1 if "running" in k]),
                bn_sum=np.array([float(sd1[k].double().sum()) for k in sd1 if "running" in k]))
    pick = ["encoder.initial_block.conv.weight", "encoder.layers.1.conv3x1_1.weight", "encoder.layers.10.conv1x3_2.weight",
            "encoder.layers.10.bn2.weight", "decoder.1.layers.0.conv.weight", "decoder.1.output_conv.weight"]
    save["pick"] = np.array(pick)
    for i, n in enumerate(pick):
        save[f"grad_{i}"] = grads[n].numpy()
    for i, t in enumerate(noise):
        if t is not None:
            save[f"noise_{i}"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "mt_model.npz"), **save)
    print("wrote mt_contract.json, mt_model.npz;", len(contract["keys"]), "state_dict entries,", len(grads), "gradients, loss", float(loss))


if __name__ == "__main__":
    main()

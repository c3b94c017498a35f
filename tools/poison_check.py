"""Debug aid (GPU box): every buffer the host layer hands to the C-ABI uninitialised (torch.empty / empty_like:
activations, saved tensors, workspaces) is filled with NaN bit patterns first.  A kernel that reads a location no
kernel wrote then shows up as NaN (or as a changed value) in the logits, the loss or a gradient.

    python tools/poison_check.py [H W [batch]]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from _util import make_sd, oracle, poisoned_empty as poisoned  # noqa: E402

DEV = "cuda"


def run(net, crit, x, y, task, noise):
    for p in net.parameters():
        p.grad = None
    out = net(x, task, drop_noise=noise)
    loss = crit(out, y)
    loss.backward()
    torch.cuda.synchronize()
    return out.detach().clone(), float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}


def main():
    from mdil_ss_b200.erfnet_RA_parallel import Net
    from mdil_ss_b200.losses import CrossEntropyLoss2d
    h = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    classes = [20, 20, 27]
    net = Net(classes, 3, 2)
    net.load_state_dict(make_sd(classes, 30, 31), strict=True)
    net = net.to(DEV).train()
    gen = torch.Generator().manual_seed(600)
    x = torch.rand(b, 3, h, w, generator=gen).to(DEV)
    for task in (0, 2):
        y = torch.randint(0, classes[task], (b, h, w), generator=gen).to(DEV)
        crit = CrossEntropyLoss2d(torch.tensor((oracle.WEIGHT_CITY, oracle.WEIGHT_BDD, oracle.WEIGHT_IDD)[task]).to(DEV))
        torch.manual_seed(77)
        noise = [None if t is None else t.to(DEV) for t in oracle.make_dropout_noise(b, True)]
        o0, l0, g0 = run(net, crit, x, y, task, noise)
        o2, l2, g2 = run(net, crit, x, y, task, noise)
        with poisoned():
            o1, l1, g1 = run(net, crit, x, y, task, noise)
        print(f"{h}x{w} b{b} task {task}: loss clean {l0:.7f} / clean again {l2:.7f} / poisoned {l1:.7f}; logits NaN {int(torch.isnan(o1).sum())}"
              f" max|d| {float((o1 - o0).abs().nan_to_num(0).max()):.2e}")
        bad = 0
        for n in g0:
            nan = int(torch.isnan(g1[n]).sum())
            d1 = float((g1[n] - g0[n]).norm().nan_to_num(0) / (g0[n].norm() + 1e-30))
            d2 = float((g2[n] - g0[n]).norm() / (g0[n].norm() + 1e-30))
            if nan or d1 > 1e-4 or d2 > 1e-4:
                bad += 1
                print(f"   {n:48s} NaN {nan:6d}/{g0[n].numel():6d}  poisoned-vs-clean {d1:.2e}  clean-vs-clean {d2:.2e}  |g| {float(g0[n].norm()):.2e}")
        print(f"   {bad} of {len(g0)} gradient tensors affected")


if __name__ == "__main__":
    main()

"""Debug aid (GPU box): the multi-task iteration of tests/test_gpu_net.py::test_multitask_iteration_matches_oracle,
visit by visit -- per-visit gradient of one tensor against the oracle's, and the per-element Adam movement.

    python tools/debug_multitask.py [tensor-name] [repeats]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from _util import make_sd, oracle  # noqa: E402

NAME = sys.argv[1] if len(sys.argv) > 1 else "encoder.layers.3.conv3x1_1.bias"
REPEATS = int(sys.argv[2]) if len(sys.argv) > 2 else 4
DEV = "cuda"


def oracle_visits(sd, batches, weights, noises, lr=5e-4):
    nb = len(batches)
    names = oracle.param_names(sd)
    state = {n: {} for n in names}
    out = []
    for ind, (images, labels) in enumerate(batches):
        work = oracle._with_grad(sd, names)
        logits = oracle.net_forward(work, images, ind, True, noises[ind])
        loss = oracle.cross_entropy2d(logits, labels[:, 0], weights[ind])
        grads = torch.autograd.grad(loss, [work[n] for n in names], allow_unused=True)
        got = {n: g for n, g in zip(names, grads) if g is not None}
        before = sd[NAME].clone()
        with torch.no_grad():
            enc = [n for n in names if "encoder" in n and n in got]
            dec = [n for n in names if "decoder" in n and n in got]
            oracle.adam_step([sd[n] for n in enc], [got[n] for n in enc], [state[n] for n in enc], lr / nb)
            oracle.adam_step([sd[n] for n in dec], [got[n] for n in dec], [state[n] for n in dec], lr)
        out.append((got[NAME].clone(), (sd[NAME] - before).clone(), float(loss)))
    return out


def main():
    from mdil_ss_b200.erfnet_RA_parallel import Net
    from mdil_ss_b200.train_step import MultiTaskTrainer
    classes = [20, 20, 27]
    sd0 = make_sd(classes, 30, 31)
    gen = torch.Generator().manual_seed(600)
    batches = [(torch.rand(2, 3, 32, 64, generator=gen), torch.randint(0, c, (2, 1, 32, 64), generator=gen)) for c in classes]
    weights = [torch.tensor(w) for w in (oracle.WEIGHT_CITY, oracle.WEIGHT_BDD, oracle.WEIGHT_IDD)]
    torch.manual_seed(77)
    noises = [oracle.make_dropout_noise(2, True) for _ in classes]
    ref = oracle_visits(oracle.clone_sd(sd0), batches, weights, noises)
    torch.set_printoptions(precision=3, linewidth=200, sci_mode=True)
    for rep in range(REPEATS):
        net = Net(classes, 3, 2)
        net.load_state_dict(sd0, strict=True)
        net = net.to(DEV)
        tr = MultiTaskTrainer(net, [w.to(DEV) for w in weights])
        p = dict(net.named_parameters())[NAME]
        net.train()
        total = torch.zeros_like(sd0[NAME], dtype=torch.float64)
        for ind, (images, labels) in enumerate(batches):
            nz = [None if t is None else t.to(DEV) for t in noises[ind]]
            out = net(images.to(DEV), ind, drop_noise=nz)
            tr.optimizer.zero_grad()
            loss = tr.criteria[ind](out, labels[:, 0].to(DEV))
            loss.backward()
            g = p.grad.detach().cpu().clone()
            before = p.detach().cpu().clone()
            tr.optimizer.step()
            d = p.detach().cpu() - before
            g_ref, d_ref, l_ref = ref[ind]
            gerr = (g - g_ref).abs()
            derr = (d - d_ref).abs().flatten()
            total += d.double()
            worst = derr.argmax().item()
            print(f"rep {rep} visit {ind}: loss {float(loss):.6f}/{l_ref:.6f}  |g| ref [{float(g_ref.abs().min()):.2e},"
                  f" {float(g_ref.abs().max()):.2e}]  max|dg| {float(gerr.max()):.2e}  rel L2 "
                  f"{float((g - g_ref).norm() / g_ref.norm()):.2e}  max|dd| {float(derr.max()):.2e} at {worst}: g "
                  f"{float(g.flatten()[worst]):.3e}/{float(g_ref.flatten()[worst]):.3e} d {float(d.flatten()[worst]):.3e}/"
                  f"{float(d_ref.flatten()[worst]):.3e}  n(|dd|>2e-5) {int((derr > 2e-5).sum())}")
        ref_total = sum(r[1].double() for r in ref)
        print(f"rep {rep}: sum|delta| {float(total.abs().sum()):.6f} vs oracle {float(ref_total.abs().sum()):.6f}")


if __name__ == "__main__":
    main()

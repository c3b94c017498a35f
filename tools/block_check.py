"""Numbers behind the block-test tolerances: for a few nb1d block cases print, per tensor, the max-norm relative error,
the fraction of elements off by more than 1e-3 of the tensor's max and the relative L2 error against the CPU oracle.
Run with MDIL_PAIR_IMPL=tc3|ffma or MDIL_H3_FMT=fp16|bf16 to A/B the arithmetic."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
from _util import oracle
from test_gpu_blocks import _randomize, _sd_cpu, _grads_by_name, _oracle_grads
from mdil_ss_b200 import erfnet_RA_parallel as M

CASES = [(128, 16, True, 2, 64, 128, 0.3), (64, 1, True, 6, 128, 256, 0.03), (128, 2, True, 6, 64, 128, 0.3)]
if len(sys.argv) > 1:
    CASES = [tuple(float(v) if "." in v else int(v) for v in arg.split(",")) for arg in sys.argv[1:]]
for (C, dil, rap, N, H, W, pdrop) in CASES:
    rap = bool(rap)
    torch.manual_seed(1)
    M.current_task = 1 if rap else 0
    mod = M.non_bottleneck_1d_RAP(C, pdrop, dil, 2) if rap else M.non_bottleneck_1d(C, pdrop, dil)
    _randomize(mod, 3)
    sd = _sd_cpu(mod, "blk")
    mod = mod.cuda().train()
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(N, C, H, W, generator=g))
    dy = torch.randn(N, C, H, W, generator=g)
    noise = torch.empty(N, C, 1, 1).bernoulli_(1 - pdrop, generator=g).div_(1 - pdrop) if pdrop > 0 else None
    names = [k for k in sd if not ("running" in k or "num_batches" in k)]
    for n in names:
        sd[n].requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    yo = oracle.nb1d(sd, "blk", xo, dil, True, 1 if rap else None, noise)
    go = _oracle_grads(dict(sd, __x=xo), names + ["__x"], (yo * dy).sum())
    xd = x.cuda().requires_grad_(True)
    yd = mod(xd, noise.cuda() if noise is not None else None)
    (yd * dy.cuda()).sum().backward()
    torch.cuda.synchronize()
    gd = _grads_by_name(mod)
    gd["__x"] = xd.grad.cpu()
    print(f"--- C={C} d={dil} rap={rap} N={N} {H}x{W} impl={os.environ.get('MDIL_PAIR_IMPL','h3')} fmt={os.environ.get('MDIL_H3_FMT','auto')}")
    def stat(a, b):
        a = a.double().cpu(); b = b.double()
        diff = (a - b).abs(); ref = b.abs().max().item()
        return diff.max().item() / ref, float((diff > 1e-3 * ref).double().mean()), float(diff.norm() / b.norm())
    print("   y      max %.2e frac %.2e l2 %.2e" % stat(yd.detach(), yo.detach()))
    for n, ref in go.items():
        key = n[len("blk."):] if n != "__x" else n
        if key.endswith("bias") and "bn" not in key:
            continue
        print("   %-22s max %.2e frac %.2e l2 %.2e" % ((key,) + stat(gd[key], ref)))

"""Localise a parity failure of the nb1d block: compare every intermediate of forward and backward with a CPU
restatement (debug tool, not a test)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch, torch.nn.functional as F
from mdil_ss_b200 import erfnet_RA_parallel as M, functional as Fn
from test_gpu_blocks import _randomize

def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(1e-30, b.abs().max()))

def run(C, dil, rap, N, H, W, pdrop):
    torch.manual_seed(1)
    M.current_task = 1 if rap else 0
    mod = M.non_bottleneck_1d_RAP(C, pdrop, dil, 2) if rap else M.non_bottleneck_1d(C, pdrop, dil)
    _randomize(mod, 3)
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    for k, v in sd.items():
        if v.dtype.is_floating_point and "running" not in k:
            v.requires_grad_(True)
    mod = mod.cuda().train()
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(N, C, H, W, generator=g)); dy = torch.randn(N, C, H, W, generator=g)
    noise = torch.empty(N, C, 1, 1).bernoulli_(1 - pdrop, generator=g).div_(1 - pdrop) if pdrop > 0 else None
    t = ".1" if rap else ""
    bn1 = ("bns_1.1" if rap else "bn1"); bn2 = ("bns_2.1" if rap else "bn2")
    xo = x.clone().requires_grad_(True)
    a = F.relu(F.conv2d(xo, sd["conv3x1_1.weight"], sd["conv3x1_1.bias"], padding=(1, 0)))
    p = F.conv2d(a, sd["conv1x3_1.weight"], sd["conv1x3_1.bias"], padding=(0, 1))
    if rap: p = p + F.conv2d(xo, sd["parallel_conv_1.1.weight"], sd["parallel_conv_1.1.bias"])
    q = F.batch_norm(p, None, None, sd[bn1 + ".weight"], sd[bn1 + ".bias"], True, 0.1, 1e-3)
    r = F.relu(q)
    c = F.relu(F.conv2d(r, sd["conv3x1_2.weight"], sd["conv3x1_2.bias"], padding=(dil, 0), dilation=(dil, 1)))
    s = F.conv2d(c, sd["conv1x3_2.weight"], sd["conv1x3_2.bias"], padding=(0, dil), dilation=(1, dil))
    if rap: s = s + F.conv2d(r, sd["parallel_conv_2.1.weight"], sd["parallel_conv_2.1.bias"])
    z = F.batch_norm(s, None, None, sd[bn2 + ".weight"], sd[bn2 + ".bias"], True, 0.1, 1e-3)
    if noise is not None: z = z * noise
    y = F.relu(z + xo)
    inter = [a, p, q, c, s]
    for tns in inter: tns.retain_grad()
    (y * dy).sum().backward()
    Fn.DEBUG_KEEP = []
    xd = x.cuda().requires_grad_(True)
    yd = mod(xd, noise.cuda() if noise is not None else None)
    ga, gp, gc, gs, stats, packed = yd.grad_fn.internal
    nhwc = lambda t: t.permute(0, 3, 1, 2)
    print(f"--- C={C} dil={dil} rap={rap} N={N} {H}x{W}")
    print("fwd  y", rel(yd, y), " a", rel(nhwc(ga), a), " p", rel(nhwc(gp), p), " c", rel(nhwc(gc), c), " s", rel(nhwc(gs), s))
    (yd * dy.cuda()).sum().backward()
    ws = Fn.DEBUG_KEEP[0]
    Tb = N * H * W * C * 4
    off = 7168 if C == 128 else None
    base = 0
    def take(nbytes):
        nonlocal base
        base = (base + 255) // 256 * 256
        o = base; base += nbytes; return o
    take(4 * C * 8); take(3 * C * 4); take(3 * C * 4)     # sums [2][2][C] fp64, coef2, coef1 (api.cu: mdil_nb1d_bwd)
    o1 = take(Tb); o2 = take(Tb); o3 = take(Tb)
    view = lambda o: ws[o:o + Tb].view(torch.float32).view(N, H, W, C).permute(0, 3, 1, 2)
    # dp = grad wrt p ; da' = grad wrt (pre-relu a)*mask = a.grad * (a>0) ; dq = q.grad
    stop = int(os.environ.get("MDIL_DEBUG_STOP", "0"))
    if stop == 0:
        print("bwd  dp(T1)", rel(view(o1), p.grad), " da'(T2)", rel(view(o2), a.grad * (a > 0)), " dq(T3)", rel(view(o3), q.grad), " dx", rel(xd.grad, xo.grad))
        for (nm, prm) in mod.named_parameters():
            if prm.grad is not None and nm in sd and sd[nm].grad is not None:
                print(f"     d{nm:28s} {rel(prm.grad, sd[nm].grad):.3e}   (max ref {float(sd[nm].grad.abs().max()):.3e})")
    else:
        print(f"bwd stop={stop} ds(T1)", rel(view(o1), s.grad), " dc'(T2)", rel(view(o2), c.grad * (c > 0)), " dq(T3)", rel(view(o3), q.grad))
        d = (view(o1).cpu() - s.grad).abs()
        print("   ds err per image:", [float(d[i].max()) for i in range(N)], " per-channel max (first 8):", [round(float(d[:, ch].max()), 4) for ch in range(8)])
        d3 = (view(o3).cpu() - q.grad).abs()
        bad = (d3 > 1e-3)
        print("   bad count", int(bad.sum()), "cols:", sorted(set(bad.nonzero()[:, 3].tolist()))[:140])
        print("   bad chans:", sorted(set(bad.nonzero()[:, 1].tolist()))[:140])
        got = view(o3).cpu(); idx = bad.nonzero()[:6]
        for (n_, c_, h_, w_) in idx.tolist():
            print("     at", (n_, c_, h_, w_), "got", float(got[n_, c_, h_, w_]), "want", float(q.grad[n_, c_, h_, w_]), "q", float(q[n_, c_, h_, w_]))
        print("   dq err per image:", [float(d3[i].max()) for i in range(N)], " rows with err>1e-3:", sorted(set((d3 > 1e-3).nonzero()[:, 2].tolist()))[:40])
    return

if __name__ == "__main__":
    import itertools
    for cfg in [(128, 16, True, 2, 4, 8, 0.0), (128, 2, True, 2, 4, 8, 0.0), (64, 1, True, 2, 8, 16, 0.0), (128, 4, True, 2, 16, 32, 0.0), (64, 1, False, 2, 8, 16, 0.0)]:
        run(*cfg)

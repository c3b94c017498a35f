"""GPU: every block kernel (forward, backward, BatchNorm buffer updates) against the oracle on the same seeded
inputs, through the product modules -> ctypes -> C ABI.  Tolerance: north star's 1e-3 relative fp32 (max-norm);
observed errors are ~1e-6."""
import os

import pytest
import torch

from _util import assert_close, oracle

pytestmark = pytest.mark.gpu
TOL = 1e-3
DEV = "cuda"


def _randomize(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 / (p[0].numel() ** 0.5)))
            elif "bn" in n and n.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)
        for n, b in mod.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.2)
            elif n.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)


def _sd_cpu(mod, prefix):
    return {f"{prefix}.{k}": v.detach().cpu().clone() for k, v in mod.state_dict().items()}


def _grads_by_name(mod):
    return {n: p.grad.detach().cpu() for n, p in mod.named_parameters() if p.grad is not None}


def _oracle_grads(sd, names, loss):
    grads = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
    return {n: g for n, g in zip(names, grads) if g is not None}


NB1D_CASES = [
    # C, dil, rap, N, H, W, dropout p
    (16, 1, False, 2, 12, 20, 0.0),
    (16, 1, False, 1, 40, 70, 0.0),
    (64, 1, True, 2, 9, 37, 0.03),
    (64, 1, False, 1, 16, 32, 0.0),
    (128, 2, True, 2, 8, 16, 0.3),
    (128, 4, True, 1, 16, 32, 0.3),
    (128, 8, True, 2, 11, 19, 0.3),
    (128, 16, True, 1, 16, 32, 0.3),
    (128, 16, True, 2, 64, 128, 0.3),
    (128, 1, True, 1, 7, 5, 0.0),
    # adapter-off blocks WITH dropout and dilation: the encoder of the multi-task joint model (models/erfnet_multi_task.py:83-92)
    (64, 1, False, 2, 9, 37, 0.03),
    (128, 4, False, 1, 16, 32, 0.3),
    (128, 16, False, 2, 11, 19, 0.3),
    # BASELINE configs[1] block shapes (batch 6 at 512x1024 -> 128x256 @ C=64, 64x128 @ C=128): many tiles per persistent
    # CTA, so the accumulator / operand-buffer recycling of the pipelined tensor-core kernel and the full-K weight
    # gradients are compared with numbers
    (64, 1, True, 6, 128, 256, 0.03),
    (128, 2, True, 6, 64, 128, 0.3),
    (128, 16, True, 6, 64, 128, 0.3),
    (16, 1, False, 6, 256, 512, 0.0),
]


@pytest.mark.parametrize("C,dil,rap,N,H,W,pdrop", NB1D_CASES)
@pytest.mark.parametrize("train", [True, False])
def test_nb1d_block(C, dil, rap, N, H, W, pdrop, train):
    from mdil_ss_b200 import erfnet_RA_parallel as M
    torch.manual_seed(1)
    M.current_task = 1 if rap else 0
    mod = M.non_bottleneck_1d_RAP(C, pdrop, dil, 2) if rap else M.non_bottleneck_1d(C, pdrop, dil)
    _randomize(mod, 3)
    sd = _sd_cpu(mod, "blk")
    mod = mod.to(DEV).train(train)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, C, H, W, generator=g)
    x = torch.relu(x)  # block inputs are post-ReLU activations in the network
    dy = torch.randn(N, C, H, W, generator=g)
    noise = None
    if train and pdrop > 0:
        noise = torch.empty(N, C, 1, 1).bernoulli_(1 - pdrop, generator=g).div_(1 - pdrop)
    # ---- oracle (CPU)
    names = [k for k in sd if not ("running" in k or "num_batches" in k)]
    for n in names:
        sd[n].requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    task = 1 if rap else None
    yo = oracle.nb1d(sd, "blk", xo, dil, train, task, noise)
    # ---- CUDA
    xd = x.to(DEV).requires_grad_(True)
    yd = mod(xd, noise.to(DEV) if noise is not None else None)
    assert_close(yd, yo, TOL, "y")
    if not train:
        return
    go = _oracle_grads(dict(sd, __x=xo), names + ["__x"], (yo * dy).sum())
    (yd * dy.to(DEV)).sum().backward()
    # large tensors: allow the rare ReLU-mask near-tie flip (see _util.assert_close).  One flipped mask element moves
    # 3 taps x C entries of dx and a whole row of a weight gradient, so the FRACTION of entries off by more than 1e-3 of
    # the tensor's max is a poor measure (measured with tools/block_check.py, identical for the 3xTF32 kernel, the
    # fp16-split and the bf16-split kernel, i.e. set by the conditioning of the block and not by the arithmetic: up to
    # 1.1e-2 for conv3x1_1.weight, 1.6e-3 .. 3.0e-3 for dx); the bound that matters is the tensor's relative L2 error,
    # <= 5e-3 (measured 0.8e-3 .. 2.0e-3; tensors behind the last ReLU of the block: 3e-6)
    out = 2e-2 if N * H * W * C >= (1 << 18) else 0.0
    # biases that feed a train-mode BatchNorm have mathematically zero gradients: what is compared is the rounding
    # noise of a sum over N*H*W pixels, hence an absolute tolerance that grows with the pixel count
    bias_atol = max(1e-3, 1e-6 * N * H * W)
    assert_close(xd.grad, go["__x"], TOL, "dx", outliers=out)
    gd = _grads_by_name(mod)
    for n, ref in go.items():
        if n == "__x":
            continue
        key = n[len("blk."):]
        assert key in gd, f"missing gradient for {key}"
        if key.endswith("bias") and ("conv1x3" in key or "parallel_conv" in key):
            # feeds a train-mode BatchNorm: mathematically zero, both sides hold rounding noise of a sum over N*H*W pixels
            assert float(gd[key].abs().max()) <= bias_atol and float(ref.abs().max()) <= bias_atol, key
        elif out > 0.0 and ref.numel() < 4096:
            # bias / BatchNorm vectors of 64-128 entries: one entry is 1-2 % of the tensor, only the L2 bound is meaningful
            l2 = float((gd[key].double() - ref.double()).norm() / ref.double().norm())
            assert l2 <= 5e-3, f"{key}: relative L2 {l2:.2e}"
        else:
            assert_close(gd[key], ref, TOL, key, atol=1e-5, outliers=out)
    # other-domain parameters receive no gradient
    if rap:
        assert "parallel_conv_1.0.weight" not in gd and "bns_2.0.weight" not in gd
    # BatchNorm buffers
    after = mod.state_dict()
    for k, v in after.items():
        if "running" in k:
            assert_close(v, sd[f"blk.{k}"], 1e-4, k, atol=1e-6)
        if "num_batches" in k:
            assert int(v) == int(sd[f"blk.{k}"])


DOWN_CASES = [(3, 16, 2, 16, 24), (3, 16, 1, 64, 128), (16, 64, 2, 10, 18), (64, 128, 2, 8, 12), (64, 128, 1, 32, 64)]


@pytest.mark.parametrize("cin,cout,N,H,W", DOWN_CASES)
@pytest.mark.parametrize("train", [True, False])
def test_downsampler(cin, cout, N, H, W, train):
    from mdil_ss_b200 import erfnet_RA_parallel as M
    torch.manual_seed(2)
    M.current_task = 1
    mod = M.DownsamplerBlock(cin, cout, 2)
    _randomize(mod, 4)
    sd = _sd_cpu(mod, "blk")
    mod = mod.to(DEV).train(train)
    g = torch.Generator().manual_seed(6)
    x = torch.rand(N, cin, H, W, generator=g)
    dy = torch.randn(N, cout, H // 2, W // 2, generator=g)
    names = [k for k in sd if not ("running" in k or "num_batches" in k)]
    for n in names:
        sd[n].requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    yo = oracle.downsampler(sd, "blk", xo, 1, train)
    xd = x.to(DEV).requires_grad_(cin != 3)
    yd = mod(xd)
    assert_close(yd, yo, TOL, "y")
    if not train:
        return
    go = _oracle_grads(dict(sd, __x=xo), names + ["__x"], (yo * dy).sum())
    (yd * dy.to(DEV)).sum().backward()
    if cin != 3:
        assert_close(xd.grad, go["__x"], TOL, "dx")
    gd = _grads_by_name(mod)
    for n, ref in go.items():
        if n == "__x":
            continue
        assert_close(gd[n[len("blk."):]], ref, TOL, n, atol=1e-4)
    for k, v in mod.state_dict().items():
        if "running" in k:
            assert_close(v, sd[f"blk.{k}"], 1e-4, k, atol=1e-6)


@pytest.mark.parametrize("cin,cout,N,H,W", [(128, 64, 2, 6, 10), (64, 16, 1, 9, 14), (128, 64, 1, 16, 32)])
@pytest.mark.parametrize("train", [True, False])
def test_upsampler(cin, cout, N, H, W, train):
    from mdil_ss_b200 import erfnet_RA_parallel as M
    torch.manual_seed(3)
    mod = M.UpsamplerBlock(cin, cout)
    _randomize(mod, 5)
    sd = _sd_cpu(mod, "blk")
    mod = mod.to(DEV).train(train)
    g = torch.Generator().manual_seed(7)
    x = torch.relu(torch.randn(N, cin, H, W, generator=g))
    dy = torch.randn(N, cout, 2 * H, 2 * W, generator=g)
    names = [k for k in sd if not ("running" in k or "num_batches" in k)]
    for n in names:
        sd[n].requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    yo = oracle.upsampler(sd, "blk", xo, train)
    xd = x.to(DEV).requires_grad_(True)
    yd = mod(xd)
    assert_close(yd, yo, TOL, "y")
    if not train:
        return
    go = _oracle_grads(dict(sd, __x=xo), names + ["__x"], (yo * dy).sum())
    (yd * dy.to(DEV)).sum().backward()
    assert_close(xd.grad, go["__x"], TOL, "dx")
    gd = _grads_by_name(mod)
    for n, ref in go.items():
        if n != "__x":
            assert_close(gd[n[len("blk."):]], ref, TOL, n, atol=1e-4)


# even W: vectorised forward + one-pass fused backward (64-pixel tiles: W = 70 and 130 end in partial tiles; 27 and 32
# classes use the wider mma column tilings); odd W: the scalar kernels
@pytest.mark.parametrize("ccls,N,H,W", [(20, 2, 8, 12), (27, 1, 16, 32), (20, 1, 33, 65), (32, 1, 6, 70), (20, 2, 5, 130),
                                         (2, 1, 4, 64)])
def test_output_conv(ccls, N, H, W):
    import torch.nn.functional as F
    from mdil_ss_b200.functional import OutConvFn
    g = torch.Generator().manual_seed(8)
    x = torch.randn(N, 16, H, W, generator=g).requires_grad_(True)
    w = (torch.randn(16, ccls, 2, 2, generator=g) * 0.3).requires_grad_(True)
    b = (torch.randn(ccls, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(N, ccls, 2 * H, 2 * W, generator=g)
    yo = F.conv_transpose2d(x, w, b, stride=2)  # decoder_forward's last line in the oracle
    (yo * dy).sum().backward()
    xd, wd, bd = (t.detach().to(DEV).requires_grad_(True) for t in (x, w, b))
    yd = OutConvFn.apply(xd, wd, bd)
    assert yd.is_contiguous() and tuple(yd.shape) == (N, ccls, 2 * H, 2 * W)
    assert_close(yd, yo, TOL, "logits")
    (yd * dy.to(DEV)).sum().backward()
    assert_close(xd.grad, x.grad, TOL, "dx")
    assert_close(wd.grad, w.grad, TOL, "dw")
    assert_close(bd.grad, b.grad, TOL, "db")


@pytest.mark.parametrize("c,N,H,W", [(20, 2, 16, 32), (27, 1, 9, 13), (5, 3, 7, 10), (32, 1, 4, 6)])
def test_cross_entropy2d_two_phase(c, N, H, W):
    """CrossEntropyLoss2d (train_new_task_step2.py:84-92) against torch's NLLLoss(log_softmax): loss-only forward, gradient
    recomputed in the backward pass and scaled by an upstream factor; odd H*W (scalar path) and an ignored class (weight
    0) included; a second backward through a retained graph gives the same gradient."""
    import torch.nn.functional as F
    from mdil_ss_b200.losses import CrossEntropyLoss2d
    g = torch.Generator().manual_seed(40 + c)
    logits = (3.0 * torch.randn(N, c, H, W, generator=g)).requires_grad_(True)
    labels = torch.randint(0, c, (N, H, W), generator=g)
    wts = torch.rand(c, generator=g) + 0.1
    wts[c - 1] = 0.0
    ref = F.nll_loss(F.log_softmax(logits, 1), labels, wts)
    (2.5 * ref).backward()
    ld = logits.detach().to(DEV).requires_grad_(True)
    loss = CrossEntropyLoss2d(wts).to(DEV)(ld, labels.to(DEV))
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    (2.5 * loss).backward(retain_graph=True)
    assert_close(ld.grad, logits.grad, 1e-4, "dlogits")
    first = ld.grad.clone()
    ld.grad = None
    (2.5 * loss).backward()
    assert torch.equal(ld.grad, first)


def test_cpu_tensor_is_rejected():
    from mdil_ss_b200 import erfnet_RA_parallel as M
    mod = M.non_bottleneck_1d(16, 0, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(torch.zeros(1, 16, 4, 4))


def test_round1_kernels_behind_the_switches_still_match():
    """The warp-level / FFMA sampler kernels, the fp32 backward workspaces and the two-kernel head backward are the product
    path for shapes the tcgen05 kernels do not cover and the A/B references of DESIGN.md: the block tests must stay green
    with every switch thrown (the switches are read once per process, hence the subprocess)."""
    import subprocess
    import sys
    env = dict(os.environ, MDIL_CONV_TC="0", MDIL_WGRAD_GATHER="0", MDIL_S16="0", MDIL_HEAD_FUSED="0", MDIL_PREPACK="0")
    sel = "downsampler or upsampler or output_conv or (nb1d_block and 64)"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", sel], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]

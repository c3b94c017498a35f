"""Shared helpers of the test-suite (test infrastructure: may import oracle/)."""
import contextlib
import os
import sys

import numpy as np
import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = os.environ.get("MDIL_REFERENCE", "/root/reference")

from oracle import erfnet_rap_oracle as oracle  # noqa: E402


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def make_sd(classes, init_seed, bn_seed):
    """The state_dict the fixtures were generated with: reference-order default init under `init_seed`
    (oracle.init_state_dict restates the constructors) + deterministic non-trivial BatchNorm state."""
    sd = oracle.init_state_dict(classes, len(classes), seed=init_seed)
    return oracle.perturb_bn_(sd, seed=bn_seed)


def pretrained_sd(npz):
    """State dict of tests/golden/pretrained_eval.npz (the reference's shipped trained weights, bf16-rounded, stored as
    16-bit patterns; tests/golden/make_golden_pretrained.py) for Net([20], 1, 0)."""
    import collections
    sd = collections.OrderedDict()
    for i, k in enumerate(str(s) for s in npz["keys"]):
        a = npz[f"w{i}"]
        if a.dtype == np.uint16:
            sd[k] = torch.from_numpy((a.astype(np.int32) << 16)).view(torch.float32).clone()
        else:
            sd[k] = torch.from_numpy(np.asarray(a)).clone()
    return sd


def noise_list(npz, prefix, n_layers=15):
    out = []
    for i in range(n_layers):
        k = f"{prefix}{i}"
        out.append(torch.from_numpy(npz[k]) if k in npz.files else None)
    return out


@contextlib.contextmanager
def poisoned_empty():
    """While active, every CUDA buffer created with torch.empty / torch.empty_like (the activations, saved tensors and
    workspaces the host layer hands to the C-ABI uninitialised) is pre-filled with NaN bit patterns, so a kernel that
    reads a location nothing wrote shows up as NaN downstream."""
    e, el = torch.empty, torch.empty_like

    def fill(t):
        if t.is_cuda and t.numel():
            if t.dtype == torch.uint8:
                t.fill_(0xFF)
            elif t.is_floating_point():
                t.fill_(float("nan"))
        return t

    torch.empty = lambda *a, **k: fill(e(*a, **k))
    torch.empty_like = lambda *a, **k: fill(el(*a, **k))
    try:
        yield
    finally:
        torch.empty, torch.empty_like = e, el


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def assert_close(a, b, tol, what="", atol=0.0, outliers=0.0, l2_tol=None):
    """max|a-b| <= tol * max|b| + atol (max-norm relative error; atol covers mathematically-zero tensors such as
    the gradient of a conv bias feeding a train-mode BatchNorm).

    ``outliers`` > 0 (gradient checks on large tensors only): a ReLU pre-activation within fp32 rounding of zero
    legitimately flips its mask under any reordering of the fp32 sums, which changes the gradient locally by O(1);
    up to that FRACTION of the elements (<= 1e-3 in every caller) may then exceed the bound, while the relative L2
    error of the whole tensor must stay <= ``l2_tol`` (default 5 * tol)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if a.numel() == 0:
        return 0.0
    diff = (a - b).abs()
    err = diff.max().item()
    ref = b.abs().max().item()
    bound = tol * ref + atol
    if outliers > 0.0 and err > bound:
        l2_tol = 5 * tol if l2_tol is None else l2_tol
        frac = float((diff > bound).double().mean())
        l2 = float(diff.norm() / max(1e-30, float(b.norm())))
        assert frac <= outliers and l2 <= l2_tol, \
            f"{what}: {frac:.2e} of the elements exceed {tol:.1e} relative (allowed {outliers:.1e}), rel L2 {l2:.2e} (allowed {l2_tol:.1e})"
        return err / ref if ref > 0 else err
    assert err <= bound, f"{what}: max abs error {err:.3e} (ref max {ref:.3e}) exceeds {tol:.1e} relative + {atol:.1e}"
    return err / ref if ref > 0 else err


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / max(1e-300, float(b.norm())))

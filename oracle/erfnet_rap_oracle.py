"""CPU oracle for the MDIL-SS hot path (TEST INFRASTRUCTURE — not product code).

A functional, state_dict-driven restatement in plain fp32 PyTorch-on-CPU of
the reference's ERFNet-with-parallel-residual-adapters forward, its two
losses and its optimiser step.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this file; the product path (``mdil_ss_b200``) never does and fails loudly if
its CUDA library is missing.

Parity pin: the reference has no tests or golden vectors of its own
("parity unpinned" by any reference test).  The oracle is
therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build
container: ``tests/golden/make_golden.py`` imports ``/root/reference`` and
writes the committed fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against them (the generating
scripts ``tests/golden/make_golden*.py`` are the live comparison with the
imported reference; they run wherever ``/root/reference`` exists).

Every function cites the reference lines it follows (paths relative to the
reference repo root).  Tensors are NCHW fp32, labels int64, exactly as the
reference sees them.  Arithmetic lives in ATen (oneDNN on CPU), the same
third-party dependency the reference calls.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-3       # models/erfnet_RA_parallel.py:19,36,44,77,86,157
BN_MOMENTUM = 0.1   # torch.nn.BatchNorm2d default, never overridden by the reference

# Encoder schedule: (kind, channels, dropout p, dilation) — models/erfnet_RA_parallel.py:126-141
ENCODER_LAYERS: List[Tuple[str, int, float, int]] = (
    [("down", 64, 0.0, 0)]
    + [("rap", 64, 0.03, 1)] * 5
    + [("down", 128, 0.0, 0)]
    + [("rap", 128, 0.3, d) for _ in range(2) for d in (2, 4, 8, 16)]
)
# Decoder schedule — models/erfnet_RA_parallel.py:171-177
DECODER_LAYERS: List[Tuple[str, int]] = [("up", 64), ("nb", 64), ("nb", 64), ("up", 16), ("nb", 16), ("nb", 16)]

# class-weight literals: train_new_task_step2.py:121-135 (last class zeroed = ignore)
WEIGHT_CITY = [2.8159904084894922, 6.9874672455551075, 3.7901719017455604, 9.94305485286704, 9.77037625072462,
               9.511470001589007, 10.310780572569994, 10.025305236316246, 4.6341256102158805, 9.561389195953845,
               7.869695292372276, 9.518873463871952, 10.374050047877898, 6.662394711556909, 10.26054487392723,
               10.28786101490449, 10.289883605859952, 10.405463349170795, 10.138502340710136, 0.0]
WEIGHT_BDD = [3.6525147483016243, 8.799815287822142, 4.781908267406055, 10.034828238618045, 9.5567865464289,
              9.645099012085169, 10.315292989325766, 10.163473632969513, 4.791692009441432, 9.556915153488912,
              4.142994047786311, 10.246903827488143, 10.47145010979545, 6.006704177894196, 9.60620532303246,
              9.964959813857726, 10.478333987902301, 10.468010534454706, 10.440929141422366, 0.0]
WEIGHT_IDD = [3.235635601598852, 6.76221624390441, 9.458242359884549, 9.446818215454014, 9.947040673126763,
              9.789672819856547, 9.476665808564432, 10.465565126694731, 9.59189547383129, 7.637805282159825,
              8.990899026692638, 9.26222234098628, 10.265657138809514, 9.386517631614392, 8.357391489170013,
              9.910382864314824, 10.389977663948363, 8.997422571963602, 10.418070541191673, 10.483262606962834,
              9.511436923349441, 7.597725385711079, 6.1734896019878205, 9.787631041755187, 3.9178330193378708,
              4.417448652936843, 0.0]

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- BN
def _bn(sd: SD, prefix: str, x: torch.Tensor, train: bool) -> torch.Tensor:
    """nn.BatchNorm2d(C, eps=1e-3) — SURVEY appendix B1.  In train mode the
    running buffers in ``sd`` are updated in place (momentum 0.1, unbiased var)
    and num_batches_tracked is incremented, as torch does."""
    if train:
        sd[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"],
                        training=train, momentum=BN_MOMENTUM, eps=BN_EPS)


# ----------------------------------------------------------------------- blocks
def downsampler(sd: SD, prefix: str, x: torch.Tensor, task: int, train: bool) -> torch.Tensor:
    """DownsamplerBlock.forward — models/erfnet_RA_parallel.py:21-25."""
    conv = F.conv2d(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], stride=2, padding=1)
    pool = F.max_pool2d(x, 2, stride=2)
    out = torch.cat([conv, pool], 1)
    out = _bn(sd, f"{prefix}.bn_ini.{task}", out, train)
    return F.relu(out)


def nb1d(sd: SD, prefix: str, x: torch.Tensor, dil: int, train: bool,
         task: Optional[int], drop_noise: Optional[torch.Tensor]) -> torch.Tensor:
    """non_bottleneck_1d_RAP.forward (task is an int; models/erfnet_RA_parallel.py:90-113)
    and non_bottleneck_1d.forward (task is None; :48-64).

    ``drop_noise`` is the Dropout2d noise tensor [N,C,1,1] already divided by
    (1-p) or None when dropout is inactive (p == 0 or eval)."""
    rap = task is not None
    a = F.relu(F.conv2d(x, sd[prefix + ".conv3x1_1.weight"], sd[prefix + ".conv3x1_1.bias"], padding=(1, 0)))
    u = F.conv2d(a, sd[prefix + ".conv1x3_1.weight"], sd[prefix + ".conv1x3_1.bias"], padding=(0, 1))
    if rap:
        u = u + F.conv2d(x, sd[f"{prefix}.parallel_conv_1.{task}.weight"], sd[f"{prefix}.parallel_conv_1.{task}.bias"])
        q = _bn(sd, f"{prefix}.bns_1.{task}", u, train)
    else:
        q = _bn(sd, f"{prefix}.bn1", u, train)
    r = F.relu(q)
    c = F.relu(F.conv2d(r, sd[prefix + ".conv3x1_2.weight"], sd[prefix + ".conv3x1_2.bias"],
                        padding=(dil, 0), dilation=(dil, 1)))
    v = F.conv2d(c, sd[prefix + ".conv1x3_2.weight"], sd[prefix + ".conv1x3_2.bias"],
                 padding=(0, dil), dilation=(1, dil))
    if rap:
        v = v + F.conv2d(r, sd[f"{prefix}.parallel_conv_2.{task}.weight"], sd[f"{prefix}.parallel_conv_2.{task}.bias"])
        z = _bn(sd, f"{prefix}.bns_2.{task}", v, train)
    else:
        z = _bn(sd, f"{prefix}.bn2", v, train)
    if drop_noise is not None:
        z = z * drop_noise
    return F.relu(z + x)


def upsampler(sd: SD, prefix: str, x: torch.Tensor, train: bool) -> torch.Tensor:
    """UpsamplerBlock.forward — models/erfnet_RA_parallel.py:159-162."""
    out = F.conv_transpose2d(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"],
                             stride=2, padding=1, output_padding=1)
    return F.relu(_bn(sd, prefix + ".bn", out, train))


# -------------------------------------------------------------------- whole net
def make_dropout_noise(n: int, train: bool, generator: Optional[torch.Generator] = None) -> List[Optional[torch.Tensor]]:
    """One Dropout2d noise tensor per encoder layer, in layer order, drawn with
    the calls F.dropout2d makes (bernoulli_(1-p).div_(1-p) on an [N,C,1,1]
    buffer).  None where the reference skips dropout."""
    out: List[Optional[torch.Tensor]] = []
    for kind, ch, p, _ in ENCODER_LAYERS:
        if kind == "rap" and train and p != 0:
            noise = torch.empty(n, ch, 1, 1).bernoulli_(1 - p, generator=generator).div_(1 - p)
            out.append(noise)
        else:
            out.append(None)
    return out


def encoder_forward(sd: SD, x: torch.Tensor, task: int, train: bool,
                    drop_noise: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
    """Encoder.forward — models/erfnet_RA_parallel.py:143-149."""
    out = downsampler(sd, "encoder.initial_block", x, task, train)
    for i, (kind, _ch, _p, dil) in enumerate(ENCODER_LAYERS):
        prefix = f"encoder.layers.{i}"
        if kind == "down":
            out = downsampler(sd, prefix, out, task, train)
        else:
            noise = drop_noise[i] if drop_noise is not None else None
            out = nb1d(sd, prefix, out, dil, train, task, noise)
    return out


def decoder_forward(sd: SD, x: torch.Tensor, task: int, train: bool) -> torch.Tensor:
    """Decoder.forward — models/erfnet_RA_parallel.py:182-190."""
    out = x
    for i, (kind, _ch) in enumerate(DECODER_LAYERS):
        prefix = f"decoder.{task}.layers.{i}"
        if kind == "up":
            out = upsampler(sd, prefix, out, train)
        else:
            out = nb1d(sd, prefix, out, 1, train, None, None)
    return F.conv_transpose2d(out, sd[f"decoder.{task}.output_conv.weight"], sd[f"decoder.{task}.output_conv.bias"],
                              stride=2)


def net_forward(sd: SD, x: torch.Tensor, task: int, train: bool,
                drop_noise: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
    """Net.forward(input, task) — models/erfnet_RA_parallel.py:207-212."""
    return decoder_forward(sd, encoder_forward(sd, x, task, train, drop_noise), task, train)


# ----------------------------------------------------------------------- losses
def cross_entropy2d(logits: torch.Tensor, target: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """CrossEntropyLoss2d — train_new_task_step2.py:84-92 (NLLLoss2d(weight) of
    log_softmax(dim=1)): -sum_i w[y_i] logp[i, y_i] / sum_i w[y_i] (appendix B6)."""
    logp = logits - torch.logsumexp(logits, dim=1, keepdim=True)
    picked = logp.gather(1, target.unsqueeze(1)).squeeze(1)
    w = weight[target]
    return -(w * picked).sum() / w.sum()


def kd_loss(student_logits: torch.Tensor, teacher_logits: torch.Tensor) -> torch.Tensor:
    """Output distillation — train_new_task_step2.py:241,296-297:
    KLDivLoss()(softmax(student), softmax(teacher)) with the default
    reduction='mean' over every element and probabilities (not log-probs) as
    the input: mean(xlogy(T, T) - T * S)."""
    s = torch.softmax(student_logits, dim=1)
    t = torch.softmax(teacher_logits, dim=1)
    return (torch.xlogy(t, t) - t * s).mean()


# -------------------------------------------------------------------- optimiser
def adam_step(params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], state: List[dict], lr: float,
              betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-4) -> None:
    """torch.optim.Adam as the drivers configure it — train_new_task_step2.py:237-239:
    L2 weight decay folded into the gradient, bias-corrected moments."""
    b1, b2 = betas
    for p, g, st in zip(params, grads, state):
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p)
            st["exp_avg_sq"] = torch.zeros_like(p)
        st["step"] += 1
        t = st["step"]
        g = g + weight_decay * p
        st["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
        st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(st["exp_avg"], denom, value=-lr / bc1)


def poly_lr_factor(epoch: int, num_epochs: int) -> float:
    """lambda1 — train_new_task_step2.py:244."""
    return pow((1 - ((epoch - 1) / num_epochs)), 0.9)


# name predicates — train_new_task_step2.py:95-105
def is_shared(n: str) -> bool:
    return "encoder" in n and "parallel_conv" not in n and "bn" not in n


def is_ds_curr(n: str, current_task: int) -> bool:
    if "decoder.{}".format(current_task) in n:
        return True
    if "encoder" in n and ("bn" in n or "parallel_conv" in n):
        return ".{}.weight".format(current_task) in n or ".{}.bias".format(current_task) in n
    return False


def param_names(sd: SD) -> List[str]:
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def trainable_names_incremental(sd: SD, current_task: int) -> List[str]:
    """Freezing policy of steps 2/3 — train_new_task_step2.py:205-215: every
    decoder != t frozen; encoder bn/parallel_conv frozen unless '.t.' in name."""
    out = []
    for n in param_names(sd):
        if "decoder" in n:
            if f"decoder.{current_task}" in n:
                out.append(n)
        elif "encoder" in n and ("bn" in n or "parallel_conv" in n):
            if f".{current_task}.weight" in n or f".{current_task}.bias" in n:
                out.append(n)
        else:
            out.append(n)
    return out


# ------------------------------------------------------------------ train steps
def _with_grad(sd: SD, names: Sequence[str]) -> SD:
    work = dict(sd)
    for n in names:
        work[n] = sd[n].detach().requires_grad_(True)
    return work


def step1_iteration(sd: SD, images: torch.Tensor, labels: torch.Tensor, weight: torch.Tensor, task: int = 0,
                    drop_noise=None, opt_state: Optional[List[dict]] = None, lr: float = 5e-4,
                    apply_update: bool = True):
    """One iteration of train_RAPFT_step1.py:287-305: fwd(train) -> CE -> backward -> Adam over all params.
    ``labels`` is [N,1,H,W] int64 as the loader yields; the loss sees labels[:,0].
    Returns (loss, logits, grads by name)."""
    names = param_names(sd)
    work = _with_grad(sd, names)
    logits = net_forward(work, images, task, True, drop_noise)
    loss = cross_entropy2d(logits, labels[:, 0], weight)
    grads = torch.autograd.grad(loss, [work[n] for n in names], allow_unused=True)
    gdict = {n: (g if g is not None else torch.zeros_like(sd[n])) for n, g in zip(names, grads)}
    if apply_update:
        state = opt_state if opt_state is not None else [dict() for _ in names]
        with torch.no_grad():
            adam_step([sd[n] for n in names], [gdict[n] for n in names], state, lr)
    return loss.detach(), logits.detach(), gdict


def step2_iteration(sd: SD, sd_old: SD, images: torch.Tensor, labels: torch.Tensor, weight: torch.Tensor,
                    task: int, lambdac: float = 0.1, drop_noise_t=None, drop_noise_prev=None,
                    opt_state: Optional[List[dict]] = None, lr_scale: float = 1.0, apply_update: bool = True):
    """One iteration of train_new_task_step2.py:271-306: student fwd(t) and fwd(t-1) in train mode, teacher
    fwd(t-1) in eval mode, CE + lambda*KD, one backward, 2-group Adam (shared encoder convs at 5e-6, :229-239)."""
    names = trainable_names_incremental(sd, task)
    work = _with_grad(sd, names)
    out_t = net_forward(work, images, task, True, drop_noise_t)
    out_prev = net_forward(work, images, task - 1, True, drop_noise_prev)
    with torch.no_grad():
        out_teacher = net_forward(sd_old, images, task - 1, False, None)
    ce = cross_entropy2d(out_t, labels[:, 0], weight)
    kd = kd_loss(out_prev, out_teacher)
    total = ce + lambdac * kd
    grads = torch.autograd.grad(total, [work[n] for n in names], allow_unused=True)
    gdict = {n: (g if g is not None else torch.zeros_like(sd[n])) for n, g in zip(names, grads)}
    if apply_update:
        shared = [n for n in names if is_shared(n)]
        ds = [n for n in names if is_ds_curr(n, task)]
        state = opt_state if opt_state is not None else [dict() for _ in shared + ds]
        with torch.no_grad():
            adam_step([sd[n] for n in shared], [gdict[n] for n in shared], state[:len(shared)], 5e-6 * lr_scale)
            adam_step([sd[n] for n in ds], [gdict[n] for n in ds], state[len(shared):], 5e-4 * lr_scale)
    return ce.detach(), kd.detach(), out_t.detach(), gdict


def step3_iteration(sd: SD, sd_old: SD, images: torch.Tensor, labels: torch.Tensor, weight: torch.Tensor,
                    task: int = 2, lambdac: float = 0.1, noises=None, opt_state: Optional[dict] = None,
                    lr_scale: float = 1.0):
    """One iteration of train_new_task_step3.py:301-356: student fwd(t) -> CE -> backward -> optimiser step; then
    student fwd(t-1), fwd(t-2) and teacher fwd(t-1), fwd(t-2) (the teacher is never put in eval mode, :301),
    lambda * (KD_{t-1} + KD_{t-2}) -> backward -> second optimiser step.  ``noises`` = the five Dropout2d streams in
    call order.  As torch.optim.Adam does after ``zero_grad()`` (grads set to None), parameters that receive no
    gradient in the KD step (the domain-t tensors) are left untouched by the second step.
    Returns (ce, kd, logits_t, CE-step grads by name, KD-step grads by name)."""
    noises = noises if noises is not None else [None] * 5
    names = trainable_names_incremental(sd, task)
    shared = [n for n in names if is_shared(n)]
    ds = [n for n in names if is_ds_curr(n, task)]
    state = opt_state if opt_state is not None else {}
    for n in names:
        state.setdefault(n, {})

    def update(gdict):
        with torch.no_grad():
            sh = [n for n in shared if n in gdict]
            dd = [n for n in ds if n in gdict]
            adam_step([sd[n] for n in sh], [gdict[n] for n in sh], [state[n] for n in sh], 5e-6 * lr_scale)
            adam_step([sd[n] for n in dd], [gdict[n] for n in dd], [state[n] for n in dd], 5e-4 * lr_scale)

    work = _with_grad(sd, names)
    out_t = net_forward(work, images, task, True, noises[0])
    ce = cross_entropy2d(out_t, labels[:, 0], weight)
    grads = torch.autograd.grad(ce, [work[n] for n in names], allow_unused=True)
    g_ce = {n: g for n, g in zip(names, grads) if g is not None}
    update(g_ce)
    work = _with_grad(sd, names)
    out_p1 = net_forward(work, images, task - 1, True, noises[1])
    out_p0 = net_forward(work, images, task - 2, True, noises[2])
    with torch.no_grad():
        old_p1 = net_forward(sd_old, images, task - 1, True, noises[3])
        old_p0 = net_forward(sd_old, images, task - 2, True, noises[4])
    kd = lambdac * (kd_loss(out_p1, old_p1) + kd_loss(out_p0, old_p0))
    grads = torch.autograd.grad(kd, [work[n] for n in names], allow_unused=True)
    g_kd = {n: g for n, g in zip(names, grads) if g is not None}
    update(g_kd)
    return ce.detach(), kd.detach(), out_t.detach(), g_ce, g_kd


def multitask_iteration(sd: SD, batches, weights, noises=None, opt_state: Optional[dict] = None, lr: float = 5e-4,
                        only: Optional[Sequence[int]] = None):
    """One iteration of train_multi_task.py:244-265 over the RAP network: for every dataset i in turn, fwd(task i) ->
    CE_i -> backward -> Adam step (encoder tensors at lr / nb_tasks, decoder tensors at lr, :209-218); tensors without a
    gradient in a visit are skipped by that visit's step.  Returns the list of losses.  ``only`` restricts the call to
    those dataset visits (the optimiser state carries over through ``opt_state``), so a test can compare visit by visit."""
    nb = len(batches)
    names = param_names(sd)
    state = opt_state if opt_state is not None else {}
    for n in names:
        state.setdefault(n, {})
    losses = []
    for ind, (images, labels) in enumerate(batches):
        if only is not None and ind not in only:
            continue
        work = _with_grad(sd, names)
        logits = net_forward(work, images, ind, True, None if noises is None else noises[ind])
        loss = cross_entropy2d(logits, labels[:, 0], weights[ind])
        grads = torch.autograd.grad(loss, [work[n] for n in names], allow_unused=True)
        got = {n: g for n, g in zip(names, grads) if g is not None}
        with torch.no_grad():
            enc = [n for n in names if "encoder" in n and n in got]
            dec = [n for n in names if "decoder" in n and n in got]
            adam_step([sd[n] for n in enc], [got[n] for n in enc], [state[n] for n in enc], lr / nb)
            adam_step([sd[n] for n in dec], [got[n] for n in dec], [state[n] for n in dec], lr)
        losses.append(loss.detach())
    return losses


# ------------------------------------------------------- constructor restatement
def init_state_dict(num_classes: Sequence[int] = (20,), nb_tasks: int = 1, seed: Optional[int] = None) -> SD:
    """Restates Net.__init__ (models/erfnet_RA_parallel.py:195-205 and the block constructors :14-19, :68-88,
    :153-157, :166-180): the same torch.nn layers created in the same order, so that under the same
    torch.manual_seed the default initialisation consumes the RNG identically and the resulting state_dict has the
    reference's keys, order, shapes and values."""
    import collections
    import torch.nn as nn
    if seed is not None:
        torch.manual_seed(seed)
    sd: SD = collections.OrderedDict()

    def put(prefix: str, mod) -> None:
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    def down(prefix: str, cin: int, cout: int) -> None:
        put(prefix + ".conv", nn.Conv2d(cin, cout - cin, (3, 3), stride=2, padding=1, bias=True))
        for t in range(nb_tasks):
            put(f"{prefix}.bn_ini.{t}", nn.BatchNorm2d(cout, eps=BN_EPS))

    def block(prefix: str, ch: int, dil: int, rap: bool) -> None:
        put(prefix + ".conv3x1_1", nn.Conv2d(ch, ch, (3, 1), padding=(1, 0)))
        put(prefix + ".conv1x3_1", nn.Conv2d(ch, ch, (1, 3), padding=(0, 1)))
        if rap:
            for t in range(nb_tasks):
                put(f"{prefix}.parallel_conv_1.{t}", nn.Conv2d(ch, ch, 1))
            for t in range(nb_tasks):
                put(f"{prefix}.bns_1.{t}", nn.BatchNorm2d(ch, eps=BN_EPS))
        else:
            put(prefix + ".bn1", nn.BatchNorm2d(ch, eps=BN_EPS))
        put(prefix + ".conv3x1_2", nn.Conv2d(ch, ch, (3, 1), padding=(dil, 0), dilation=(dil, 1)))
        put(prefix + ".conv1x3_2", nn.Conv2d(ch, ch, (1, 3), padding=(0, dil), dilation=(1, dil)))
        if rap:
            for t in range(nb_tasks):
                put(f"{prefix}.parallel_conv_2.{t}", nn.Conv2d(ch, ch, 1))
            for t in range(nb_tasks):
                put(f"{prefix}.bns_2.{t}", nn.BatchNorm2d(ch, eps=BN_EPS))
        else:
            put(prefix + ".bn2", nn.BatchNorm2d(ch, eps=BN_EPS))

    down("encoder.initial_block", 3, 16)
    cin = 16
    for i, (kind, ch, _p, dil) in enumerate(ENCODER_LAYERS):
        if kind == "down":
            down(f"encoder.layers.{i}", cin, ch)
            cin = ch
        else:
            block(f"encoder.layers.{i}", ch, dil, True)
    for t in range(nb_tasks):
        cin = 128
        for i, (kind, ch) in enumerate(DECODER_LAYERS):
            prefix = f"decoder.{t}.layers.{i}"
            if kind == "up":
                put(prefix + ".conv", nn.ConvTranspose2d(cin, ch, 3, stride=2, padding=1, output_padding=1, bias=True))
                put(prefix + ".bn", nn.BatchNorm2d(ch, eps=BN_EPS))
                cin = ch
            else:
                block(prefix, ch, 1, False)
        put(f"decoder.{t}.output_conv", nn.ConvTranspose2d(16, num_classes[t], 2, stride=2, padding=0,
                                                           output_padding=0, bias=True))
    return sd


def perturb_bn_(sd: SD, seed: int = 7) -> SD:
    """Test helper (no reference counterpart): give every BatchNorm non-trivial affine parameters and running
    statistics, deterministically, so eval-mode BN is not an identity and every per-domain copy differs."""
    g = torch.Generator().manual_seed(seed)
    for k in sd:
        if ".bn" in k:
            v = sd[k]
            if k.endswith(".weight") or k.endswith("running_var"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif k.endswith(".bias") or k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.2)
    return sd


def clone_sd(sd: SD) -> SD:
    return type(sd)((k, v.detach().clone()) for k, v in sd.items())


def multitask_sd_as_rap(sd_mt: SD, nb_tasks: int) -> SD:
    """models/erfnet_multi_task.py restated through the RAP restatement: its Net (:150-163) is the RAP network with
    ONE BatchNorm per position shared by all domains (``bn`` / ``bn1`` / ``bn2``, :18,36,44 -> ``bn_ini.t`` /
    ``bns_1.t`` / ``bns_2.t`` for every t, the same tensors) and no adapters (zero ``parallel_conv``), with the same
    encoder Dropout2d (:83-92).  The returned dict ALIASES the multi-task tensors, so train-mode running-statistic
    updates and autograd leaves are shared with ``sd_mt``."""
    import collections
    out: SD = collections.OrderedDict()
    for k, v in sd_mt.items():
        if k.startswith("decoder."):
            out[k] = v
            continue
        done = False
        for src, dst in ((".bn1.", ".bns_1."), (".bn2.", ".bns_2."), (".bn.", ".bn_ini.")):
            if src in k:
                for t in range(nb_tasks):
                    out[k.replace(src, f"{dst}{t}.")] = v
                done = True
                break
        if not done:
            out[k] = v
    for i, (kind, ch, _p, _d) in enumerate(ENCODER_LAYERS):
        if kind == "rap":
            for which in (1, 2):
                for t in range(nb_tasks):
                    out[f"encoder.layers.{i}.parallel_conv_{which}.{t}.weight"] = torch.zeros(ch, ch, 1, 1)
                    out[f"encoder.layers.{i}.parallel_conv_{which}.{t}.bias"] = torch.zeros(ch)
    return out


# ------------------------------------------------------------------------ metric
def iou_add_batch(pred: torch.Tensor, gt: torch.Tensor, n_classes: int, ignore_index: int):
    """iouEval.addBatch — iouEval.py:21-70, for index inputs [N,1,H,W]: one-hot both, drop the ignore channel,
    tp = sum x*y, fp = sum x*(1-y-ignores), fn = sum (1-x)*y.  Returns float64 (tp, fp, fn) per class."""
    x1 = torch.zeros(pred.size(0), n_classes, pred.size(2), pred.size(3)).scatter_(1, pred, 1).float()
    y1 = torch.zeros(gt.size(0), n_classes, gt.size(2), gt.size(3)).scatter_(1, gt, 1).float()
    if ignore_index != -1:
        ignores = y1[:, ignore_index].unsqueeze(1)
        x1 = x1[:, :ignore_index]
        y1 = y1[:, :ignore_index]
    else:
        ignores = 0
    tp = (x1 * y1).sum(dim=(0, 2, 3)).double()
    fp = (x1 * (1 - y1 - ignores)).sum(dim=(0, 2, 3)).double()
    fn = ((1 - x1) * y1).sum(dim=(0, 2, 3)).double()
    return tp, fp, fn


def iou_from_counts(tp: torch.Tensor, fp: torch.Tensor, fn: torch.Tensor):
    """iouEval.getIoU — iouEval.py:72-77."""
    iou = tp / (tp + fp + fn + 1e-15)
    return torch.mean(iou), iou

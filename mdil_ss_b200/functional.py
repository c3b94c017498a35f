"""torch.autograd.Function wrappers around the C ABI of libmdil_b200.so.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every FLOP of the hot
path runs in the sm_100a kernels.  Internal activations are logical NCHW tensors whose memory is
NHWC (channels_last strides), so the reference's shape conventions hold everywhere.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L

BN_EPS = 1e-3        # models/erfnet_RA_parallel.py:19,36,44,77,86,157
BN_MOMENTUM = 0.1    # nn.BatchNorm2d default
DEBUG_KEEP = None    # tools/debug_nb1d.py sets this to a list to inspect the backward workspace


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what}: mdil_ss_b200 runs on CUDA (sm_100a) tensors only — there is no CPU fallback")
    if x.dtype != torch.float32:
        raise RuntimeError(f"{what}: fp32 tensors required, got {x.dtype}")


def is_nhwc(x: torch.Tensor) -> bool:
    n, c, h, w = x.shape
    return x.stride() == (h * w * c, 1, w * c, c)


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """Logical [N,C,H,W] tensor whose memory is dense NHWC."""
    if is_nhwc(x):
        return x
    return x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def empty_nhwc(n: int, c: int, h: int, w: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty((n, h, w, c), device=like.device, dtype=torch.float32).permute(0, 3, 1, 2)


def grad_target(param: torch.Tensor, need: bool):
    """Where a backward kernel writes a parameter's gradient: (buffer, value returned to autograd).

    Parameters whose ``.grad`` lives in a freshly zeroed flat gradient buffer (parallel.FlatAdam.zero_grad marks them
    ``_mdil_grad_fresh``) take their FIRST gradient of the step directly in that buffer (the kernels overwrite) and
    autograd gets None: no temporary, no accumulation kernel.  Any later contribution in the same step (a weight used
    by two forwards, as in steps 2/3) goes through a temporary that autograd adds as usual."""
    if not need:
        return None, None
    g = param.grad if param.is_leaf else None      # nn.DataParallel replicas hold non-leaf views of the parameters
    if getattr(param, "_mdil_grad_fresh", False) and g is not None and g.is_contiguous() and g.shape == param.shape \
            and g.device == param.device:
        param._mdil_grad_fresh = False
        return g, None
    t = torch.empty_like(param)
    return t, t


def _bn_struct(w, b, rm, rv, nbt=None) -> L.BnParams:
    return L.BnParams(_ptr(w), _ptr(b), _ptr(rm), _ptr(rv), _ptr(nbt))


class PackedCache:
    """Kernel-friendly copies of a block's weights, refreshed when a parameter's version or storage changes.

    ``tensor._version`` does not move under every in-place update (``dist.broadcast(p)``, ``p.data.mul_()``,
    ``dist.all_reduce(p.data)`` leave it unchanged): code that rewrites weights that way must call
    :func:`invalidate_packed` on the module (``parallel.broadcast_module`` and the checkpoint helpers do)."""

    def __init__(self) -> None:
        self._entries = {}

    def invalidate(self) -> None:
        self._entries.clear()

    def get(self, key, tensors: Sequence[torch.Tensor], nfloats: int, pack_fn) -> torch.Tensor:
        sig = tuple((t.data_ptr(), t._version) for t in tensors)
        key = (key, tensors[0].device)
        ent = self._entries.get(key)
        if ent is None or ent[0] != sig or ent[1].device != tensors[0].device:
            buf = ent[1] if ent is not None and ent[1].device == tensors[0].device and ent[1].numel() == nfloats \
                else torch.empty(nfloats, device=tensors[0].device, dtype=torch.float32)
            pack_fn(buf)
            self._entries[key] = (sig, buf)
            return buf
        return ent[1]


def invalidate_packed(module: torch.nn.Module) -> None:
    """Drop every packed weight copy held under ``module`` (call after any weight update that does not go through
    autograd-visible in-place ops or FlatAdam: ``.data`` writes, collectives on parameters, custom EMA / clipping)."""
    for m in module.modules():
        cache = getattr(m, "_cache", None)
        if isinstance(cache, PackedCache):
            cache.invalidate()


# ======================================================================================= nb1d
class Nb1dConfig:
    __slots__ = ("dil", "has_adapter", "train", "bn1_buffers", "bn2_buffers", "cache", "cache_key", "grad")

    def __init__(self, dil, has_adapter, train, bn1_buffers, bn2_buffers, cache, cache_key):
        self.dil, self.has_adapter, self.train = dil, has_adapter, train
        self.grad = torch.is_grad_enabled()
        self.bn1_buffers, self.bn2_buffers = bn1_buffers, bn2_buffers
        self.cache, self.cache_key = cache, cache_key


def _nb1d_weights(params, cfg: Nb1dConfig) -> L.Nb1dWeights:
    (w31_1, b31_1, w13_1, b13_1, w31_2, b31_2, w13_2, b13_2, bn1_w, bn1_b, bn2_w, bn2_b) = params[:12]
    wp1 = bp1 = wp2 = bp2 = None
    if cfg.has_adapter:
        wp1, bp1, wp2, bp2 = params[12:16]
    w = L.Nb1dWeights()
    for name, t in (("w31_1", w31_1), ("b31_1", b31_1), ("w13_1", w13_1), ("b13_1", b13_1), ("w31_2", w31_2),
                    ("b31_2", b31_2), ("w13_2", w13_2), ("b13_2", b13_2), ("wp1", wp1), ("bp1", bp1), ("wp2", wp2),
                    ("bp2", bp2)):
        setattr(w, name, _ptr(t))
    w.bn1 = _bn_struct(bn1_w, bn1_b, *cfg.bn1_buffers)
    w.bn2 = _bn_struct(bn2_w, bn2_b, *cfg.bn2_buffers)
    return w


def prepack_nb1d(cache: PackedCache, cache_key, has_adapter: bool, params) -> None:
    """Refresh the packed weights of a nb1d block ahead of its forward (same cache key, tensors and pack call as
    Nb1dFn.forward, which then finds them current).  The packing launch depends only on the channel counts; it goes to the
    current stream -- erfnet_RA_parallel.Net runs it on a side stream under the first layers of the step."""
    lib = L.lib()
    w0 = params[0]
    c = int(w0.shape[0])
    cfg = Nb1dConfig(1, has_adapter, False, (None, None, None), (None, None, None), cache, cache_key)
    desc = L.Nb1dDesc(1, 4, 4, c, 1, int(has_adapter), 0, 0, BN_EPS, BN_MOMENTUM)
    with torch.cuda.device_of(w0):
        wts = _nb1d_weights(params, cfg)
        conv_params = list(params[:8:2]) + (list(params[12:16:2]) if has_adapter else [])

        def pack(buf):
            L.check(lib.mdil_nb1d_pack(C.byref(desc), C.byref(wts), buf.data_ptr(), _stream()), "mdil_nb1d_pack")

        cache.get(cache_key, conv_params, int(lib.mdil_nb1d_packed_floats(c)), pack)


def prepack_down(cache: PackedCache, cache_key, conv_w, cin: int) -> None:
    """As prepack_nb1d for DownFn (the packed layouts depend on the channel counts and the input's row stride only)."""
    lib = L.lib()
    cout = int(conv_w.shape[0]) + cin
    ldin = cin if cin % 4 == 0 else 4
    desc = L.DownDesc(1, 4, 4, cin, cout, ldin, 0, 0, BN_EPS, BN_MOMENTUM)
    with torch.cuda.device_of(conv_w):
        def pack(buf):
            L.check(lib.mdil_down_pack(C.byref(desc), conv_w.data_ptr(), buf.data_ptr(), _stream()), "mdil_down_pack")

        cache.get(cache_key, [conv_w], int(lib.mdil_down_packed_floats(C.byref(desc))), pack)


def prepack_up(cache: PackedCache, cache_key, conv_w) -> None:
    """As prepack_nb1d for UpFn."""
    lib = L.lib()
    cin, cout = int(conv_w.shape[0]), int(conv_w.shape[1])
    desc = L.UpDesc(1, 4, 4, cin, cout, 0, 0, BN_EPS, BN_MOMENTUM)
    with torch.cuda.device_of(conv_w):
        def pack(buf):
            L.check(lib.mdil_up_pack(C.byref(desc), conv_w.data_ptr(), buf.data_ptr(), _stream()), "mdil_up_pack")

        cache.get(cache_key, [conv_w], int(lib.mdil_up_packed_floats(C.byref(desc))), pack)


class Nb1dFn(torch.autograd.Function):
    """non_bottleneck_1d_RAP.forward / non_bottleneck_1d.forward (models/erfnet_RA_parallel.py:90-113, 48-64).

    params: w31_1,b31_1,w13_1,b13_1,w31_2,b31_2,w13_2,b13_2, bn1_w,bn1_b,bn2_w,bn2_b [, wp1,bp1,wp2,bp2]
    """

    @staticmethod
    def forward(ctx, x, drop, cfg: Nb1dConfig, *params):
        _require_cuda(x, "nb1d")
        lib = L.lib()
        x = to_nhwc(x)
        n, c, h, w = x.shape
        need_grad = cfg.grad and any(ctx.needs_input_grad)
        save = bool(cfg.train and need_grad)
        desc = L.Nb1dDesc(n, h, w, c, cfg.dil, int(cfg.has_adapter), int(cfg.train), int(save), BN_EPS, BN_MOMENTUM)
        with torch.cuda.device_of(x):
            wts = _nb1d_weights(params, cfg)
            conv_params = list(params[:8:2]) + (list(params[12:16:2]) if cfg.has_adapter else [])

            def pack(buf):
                L.check(lib.mdil_nb1d_pack(C.byref(desc), C.byref(wts), buf.data_ptr(), _stream()), "mdil_nb1d_pack")

            packed = cfg.cache.get(cfg.cache_key, conv_params, int(lib.mdil_nb1d_packed_floats(c)), pack)
            y = empty_nhwc(n, c, h, w, x)
            p = torch.empty((n, h, w, c), device=x.device, dtype=torch.float32)
            s = torch.empty_like(p)
            a = torch.empty_like(p) if save else None
            cc = torch.empty_like(p) if save else None
            stats = torch.empty((8, c), device=x.device, dtype=torch.float32)
            saved = L.Nb1dSaved(_ptr(a), _ptr(p), _ptr(cc), _ptr(s), _ptr(stats))
            ws_bytes = int(lib.mdil_nb1d_fwd_workspace_bytes(C.byref(desc)))
            ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
            if drop is not None:
                drop = drop.reshape(n, c).contiguous()
            L.check(lib.mdil_nb1d_fwd(C.byref(desc), x.data_ptr(), C.byref(wts), packed.data_ptr(), _ptr(drop),
                                      y.data_ptr(), C.byref(saved), ws.data_ptr(), ws_bytes, _stream()), "mdil_nb1d_fwd")
        ctx.cfg = cfg
        ctx.train = bool(cfg.train)
        if need_grad:
            ctx.param_objs = params
            ctx.save_for_backward(x, y, drop if drop is not None else torch.empty(0, device=x.device), *params)
            ctx.internal = (a, p, cc, s, stats, packed)
            ctx.has_drop = drop is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        if not ctx.train:
            raise RuntimeError("mdil_ss_b200: backward through an eval-mode (running-statistics) block is not implemented")
        lib = L.lib()
        cfg = ctx.cfg
        x, y, drop, *params = ctx.saved_tensors
        a, p, cc, s, stats, packed = ctx.internal
        drop = drop if ctx.has_drop else None
        dy = to_nhwc(dy)
        n, c, h, w = x.shape
        desc = L.Nb1dDesc(n, h, w, c, cfg.dil, int(cfg.has_adapter), 1, 1, BN_EPS, BN_MOMENTUM)
        needs = ctx.needs_input_grad[3:]
        with torch.cuda.device_of(x):
            wts = _nb1d_weights(params, cfg)
            grads = [None] * len(params)
            rets = [None] * len(params)
            objs = ctx.param_objs
            # weight/bias pairs are produced together
            pairs = [(0, 1), (2, 3), (4, 5), (6, 7)] + ([(12, 13), (14, 15)] if cfg.has_adapter else [])
            for wi, bi in pairs:
                if needs[wi] or needs[bi]:
                    grads[wi], rets[wi] = grad_target(objs[wi], True) if needs[wi] else (torch.empty_like(params[wi]), None)
                    grads[bi], rets[bi] = grad_target(objs[bi], True) if needs[bi] else (torch.empty_like(params[bi]), None)
            for i in (8, 9, 10, 11):
                if needs[i]:
                    grads[i], rets[i] = grad_target(objs[i], True)
            g = L.Nb1dGrads()
            names = ["w31_1", "b31_1", "w13_1", "b13_1", "w31_2", "b31_2", "w13_2", "b13_2", "bn1_w", "bn1_b", "bn2_w",
                     "bn2_b", "wp1", "bp1", "wp2", "bp2"]
            for i, name in enumerate(names[:len(params)]):
                setattr(g, name, _ptr(grads[i]))
            dx = empty_nhwc(n, c, h, w, x)
            saved = L.Nb1dSaved(_ptr(a), _ptr(p), _ptr(cc), _ptr(s), _ptr(stats))
            ws_bytes = int(lib.mdil_nb1d_bwd_workspace_bytes(C.byref(desc)))
            ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
            L.check(lib.mdil_nb1d_bwd(C.byref(desc), dy.data_ptr(), x.data_ptr(), y.data_ptr(), C.byref(wts),
                                      packed.data_ptr(), _ptr(drop), C.byref(saved), dx.data_ptr(), C.byref(g),
                                      ws.data_ptr(), ws_bytes, _stream()), "mdil_nb1d_bwd")
            if DEBUG_KEEP is not None:
                DEBUG_KEEP.append(ws)
        return (dx if ctx.needs_input_grad[0] else None, None, None, *rets)


# ======================================================================================= down / up
class SampConfig:
    __slots__ = ("train", "bn_buffers", "cache", "cache_key", "grad")

    def __init__(self, train, bn_buffers, cache, cache_key):
        self.train, self.bn_buffers, self.cache, self.cache_key = train, bn_buffers, cache, cache_key
        self.grad = torch.is_grad_enabled()


class DownFn(torch.autograd.Function):
    """DownsamplerBlock.forward (models/erfnet_RA_parallel.py:21-25). params: conv_w, conv_b, bn_w, bn_b."""

    @staticmethod
    def forward(ctx, x, cfg: SampConfig, conv_w, conv_b, bn_w, bn_b):
        _require_cuda(x, "downsampler")
        lib = L.lib()
        n, cin, h, w = x.shape
        if h % 2 or w % 2:
            raise RuntimeError("downsampler: H and W must be even")
        cout = conv_w.shape[0] + cin
        need_grad = cfg.grad and any(ctx.needs_input_grad)
        with torch.cuda.device_of(x):
            if cin % 4 != 0:
                # network input: NCHW image (3 channels) -> NHWC padded to 4 channels
                xin = x.contiguous()
                x4 = torch.empty((n, h, w, 4), device=x.device, dtype=torch.float32)
                L.check(lib.mdil_nchw_to_nhwc4(xin.data_ptr(), x4.data_ptr(), n, cin, h, w, _stream()), "mdil_nchw_to_nhwc4")
                xk, ldin = x4, 4
            else:
                xk, ldin = to_nhwc(x), cin
            desc = L.DownDesc(n, h, w, cin, cout, ldin, int(cfg.train), int(need_grad), BN_EPS, BN_MOMENTUM)

            def pack(buf):
                L.check(lib.mdil_down_pack(C.byref(desc), conv_w.data_ptr(), buf.data_ptr(), _stream()), "mdil_down_pack")

            packed = cfg.cache.get(cfg.cache_key, [conv_w], int(lib.mdil_down_packed_floats(C.byref(desc))), pack)
            oh, ow = h // 2, w // 2
            u = torch.empty((n, oh, ow, cout), device=x.device, dtype=torch.float32)
            y = empty_nhwc(n, cout, oh, ow, x)
            stats = torch.empty((4, cout), device=x.device, dtype=torch.float32)
            ws_bytes = 2 * cout * 8 + 1024
            ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
            bn = _bn_struct(bn_w, bn_b, *cfg.bn_buffers)
            L.check(lib.mdil_down_fwd(C.byref(desc), xk.data_ptr(), packed.data_ptr(), conv_b.data_ptr(), C.byref(bn),
                                      u.data_ptr(), stats.data_ptr(), y.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                    "mdil_down_fwd")
        ctx.cfg, ctx.train = cfg, bool(cfg.train)
        if need_grad:
            ctx.param_objs = (conv_w, conv_b, bn_w, bn_b)
            ctx.save_for_backward(xk, y, conv_w, conv_b, bn_w, bn_b)
            ctx.internal = (u, stats, packed, (n, cin, h, w, cout, ldin))
        return y

    @staticmethod
    def backward(ctx, dy):
        if not ctx.train:
            raise RuntimeError("mdil_ss_b200: backward through an eval-mode (running-statistics) block is not implemented")
        lib = L.lib()
        xk, y, conv_w, conv_b, bn_w, bn_b = ctx.saved_tensors
        u, stats, packed, (n, cin, h, w, cout, ldin) = ctx.internal
        dy = to_nhwc(dy)
        need_x, _, need_w, need_b, need_g, need_be = ctx.needs_input_grad
        with torch.cuda.device_of(dy):
            desc = L.DownDesc(n, h, w, cin, cout, ldin, 1, 1, BN_EPS, BN_MOMENTUM)
            dx = empty_nhwc(n, cin, h, w, dy) if need_x else None
            pw, pb, pg, pbe = ctx.param_objs
            dw = db = rw = rb = None
            if need_w or need_b:
                dw, rw = grad_target(pw, True) if need_w else (torch.empty_like(conv_w), None)
                db, rb = grad_target(pb, True) if need_b else (torch.empty_like(conv_b), None)
            dg, rg = grad_target(pg, need_g)
            dbe, rbe = grad_target(pbe, need_be)
            ws_bytes = int(lib.mdil_down_workspace_bytes(C.byref(desc)))
            ws = torch.empty(ws_bytes, device=dy.device, dtype=torch.uint8)
            bn = _bn_struct(bn_w, bn_b, *ctx.cfg.bn_buffers)
            L.check(lib.mdil_down_bwd(C.byref(desc), dy.data_ptr(), xk.data_ptr(), u.data_ptr(), y.data_ptr(),
                                      stats.data_ptr(), packed.data_ptr(), C.byref(bn), _ptr(dx), _ptr(dw), _ptr(db),
                                      _ptr(dg), _ptr(dbe), ws.data_ptr(), ws_bytes, _stream()), "mdil_down_bwd")
        return dx, None, rw, rb, rg, rbe


class UpFn(torch.autograd.Function):
    """UpsamplerBlock.forward (models/erfnet_RA_parallel.py:159-162). params: conv_w [Cin,Cout,3,3], conv_b, bn_w, bn_b."""

    @staticmethod
    def forward(ctx, x, cfg: SampConfig, conv_w, conv_b, bn_w, bn_b):
        _require_cuda(x, "upsampler")
        lib = L.lib()
        x = to_nhwc(x)
        n, cin, h, w = x.shape
        cout = conv_w.shape[1]
        need_grad = cfg.grad and any(ctx.needs_input_grad)
        with torch.cuda.device_of(x):
            desc = L.UpDesc(n, h, w, cin, cout, int(cfg.train), int(need_grad), BN_EPS, BN_MOMENTUM)

            def pack(buf):
                L.check(lib.mdil_up_pack(C.byref(desc), conv_w.data_ptr(), buf.data_ptr(), _stream()), "mdil_up_pack")

            packed = cfg.cache.get(cfg.cache_key, [conv_w], int(lib.mdil_up_packed_floats(C.byref(desc))), pack)
            u = torch.empty((n, 2 * h, 2 * w, cout), device=x.device, dtype=torch.float32)
            y = empty_nhwc(n, cout, 2 * h, 2 * w, x)
            stats = torch.empty((4, cout), device=x.device, dtype=torch.float32)
            ws_bytes = 2 * cout * 8 + 1024
            ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
            bn = _bn_struct(bn_w, bn_b, *cfg.bn_buffers)
            L.check(lib.mdil_up_fwd(C.byref(desc), x.data_ptr(), packed.data_ptr(), conv_b.data_ptr(), C.byref(bn),
                                    u.data_ptr(), stats.data_ptr(), y.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                    "mdil_up_fwd")
        ctx.cfg, ctx.train = cfg, bool(cfg.train)
        if need_grad:
            ctx.param_objs = (conv_w, conv_b, bn_w, bn_b)
            ctx.save_for_backward(x, y, conv_w, conv_b, bn_w, bn_b)
            ctx.internal = (u, stats, packed, (n, cin, h, w, cout))
        return y

    @staticmethod
    def backward(ctx, dy):
        if not ctx.train:
            raise RuntimeError("mdil_ss_b200: backward through an eval-mode (running-statistics) block is not implemented")
        lib = L.lib()
        x, y, conv_w, conv_b, bn_w, bn_b = ctx.saved_tensors
        u, stats, packed, (n, cin, h, w, cout) = ctx.internal
        dy = to_nhwc(dy)
        need_x, _, need_w, need_b, need_g, need_be = ctx.needs_input_grad
        with torch.cuda.device_of(dy):
            desc = L.UpDesc(n, h, w, cin, cout, 1, 1, BN_EPS, BN_MOMENTUM)
            dx = empty_nhwc(n, cin, h, w, dy) if need_x else None
            pw, pb, pg, pbe = ctx.param_objs
            dw = db = rw = rb = None
            if need_w or need_b:
                dw, rw = grad_target(pw, True) if need_w else (torch.empty_like(conv_w), None)
                db, rb = grad_target(pb, True) if need_b else (torch.empty_like(conv_b), None)
            dg, rg = grad_target(pg, need_g)
            dbe, rbe = grad_target(pbe, need_be)
            ws_bytes = int(lib.mdil_up_workspace_bytes(C.byref(desc)))
            ws = torch.empty(ws_bytes, device=dy.device, dtype=torch.uint8)
            bn = _bn_struct(bn_w, bn_b, *ctx.cfg.bn_buffers)
            L.check(lib.mdil_up_bwd(C.byref(desc), dy.data_ptr(), x.data_ptr(), u.data_ptr(), y.data_ptr(),
                                    stats.data_ptr(), packed.data_ptr(), C.byref(bn), _ptr(dx), _ptr(dw), _ptr(db),
                                    _ptr(dg), _ptr(dbe), ws.data_ptr(), ws_bytes, _stream()), "mdil_up_bwd")
        return dx, None, rw, rb, rg, rbe


# ======================================================================================= output conv
class OutConvFn(torch.autograd.Function):
    """Decoder.output_conv = ConvTranspose2d(16, C, 2, stride=2) (models/erfnet_RA_parallel.py:179-180,188).
    x NHWC [N,16,H,W] -> logits NCHW-contiguous [N,C,2H,2W]."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _require_cuda(x, "output_conv")
        lib = L.lib()
        x = to_nhwc(x)
        n, cin, h, w = x.shape
        if cin != 16 or weight.shape[0] != 16 or tuple(weight.shape[2:]) != (2, 2):
            raise RuntimeError("output_conv: expected ConvTranspose2d(16, C, 2, stride=2)")
        ccls = weight.shape[1]
        with torch.cuda.device_of(x):
            logits = torch.empty((n, ccls, 2 * h, 2 * w), device=x.device, dtype=torch.float32)
            L.check(lib.mdil_outconv_fwd(x.data_ptr(), weight.data_ptr(), _ptr(bias), logits.data_ptr(), n, h, w, ccls,
                                         _stream()), "mdil_outconv_fwd")
        ctx.param_objs = (weight, bias)
        ctx.save_for_backward(x, weight, bias)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = L.lib()
        x, weight, bias = ctx.saved_tensors
        n, cin, h, w = x.shape
        ccls = weight.shape[1]
        dlogits = dlogits.contiguous()
        need_x, need_w, need_b = ctx.needs_input_grad
        with torch.cuda.device_of(x):
            dx = empty_nhwc(n, cin, h, w, x) if need_x else None
            dw, rw = grad_target(ctx.param_objs[0], need_w)
            db, rb = grad_target(ctx.param_objs[1], need_b) if bias is not None else (None, None)
            L.check(lib.mdil_outconv_bwd(dlogits.data_ptr(), x.data_ptr(), weight.data_ptr(), _ptr(dx), _ptr(dw), _ptr(db),
                                         n, h, w, ccls, _stream()), "mdil_outconv_bwd")
        return dx, rw, rb


# ======================================================================================= losses
class CrossEntropy2dFn(torch.autograd.Function):
    """CrossEntropyLoss2d.forward (train_new_task_step2.py:84-92): NLLLoss(weight)(log_softmax(logits, 1), target).
    Two passes over the logits: the forward reads them once for the loss sums; the backward recomputes the softmax and
    writes the logit gradient already scaled by grad_out / sum(w) (mdil_ce2d_bwd)."""

    @staticmethod
    def forward(ctx, logits, target, weight, group_norm=None):
        _require_cuda(logits, "cross_entropy2d")
        lib = L.lib()
        logits = logits.contiguous()
        n, c, h, w = logits.shape
        if target.dim() == 4 and target.shape[1] == 1:
            target = target[:, 0]
        if tuple(target.shape) != (n, h, w) or target.dtype != torch.int64:
            raise RuntimeError("cross_entropy2d: target must be int64 [N,H,W]")
        target = target.contiguous()
        with torch.cuda.device_of(logits):
            if weight is None:
                weight = torch.ones(c, device=logits.device, dtype=torch.float32)
            weight = weight.to(device=logits.device, dtype=torch.float32).contiguous()
            loss = torch.empty((), device=logits.device, dtype=torch.float32)
            acc = torch.empty(2, device=logits.device, dtype=torch.float64)
            L.check(lib.mdil_ce2d_fwd_bwd(logits.data_ptr(), target.data_ptr(), weight.data_ptr(), n, c, h, w,
                                          loss.data_ptr(), acc.data_ptr(), None, _stream()), "mdil_ce2d_fwd_bwd")
            ctx.world = 1
            if group_norm is not None:
                # nn.DataParallel computes the loss on the gathered logits: sum(w*nll) / sum(w) over the GLOBAL batch
                # (SURVEY 8e).  One all-reduce of the two fp64 accumulators makes every rank return that value and
                # scale its logit gradient by world / sum_global(w) (the optimiser averages the ranks' gradients).
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size(group_norm or None) > 1:
                    grp = group_norm or None
                    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=grp)
                    ctx.world = dist.get_world_size(grp)
                    loss = (acc[0] / acc[1]).to(torch.float32)
        ctx.save_for_backward(logits, target, weight, acc)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        lib = L.lib()
        logits, target, weight, acc = ctx.saved_tensors
        n, c, h, w = logits.shape
        with torch.cuda.device_of(logits):
            g = (grad_out.to(dtype=torch.float32) * float(ctx.world)).contiguous()
            dlogits = torch.empty_like(logits)
            L.check(lib.mdil_ce2d_bwd(logits.data_ptr(), target.data_ptr(), weight.data_ptr(), n, c, h, w, acc.data_ptr(),
                                      g.data_ptr(), dlogits.data_ptr(), _stream()), "mdil_ce2d_bwd")
        return dlogits, None, None, None


class OutputKDFn(torch.autograd.Function):
    """KLDivLoss()(softmax(student,1), softmax(teacher,1)) exactly as train_new_task_step2.py:296-297 calls it
    (probabilities, not log-probabilities, as the input; reduction 'mean' over every element)."""

    @staticmethod
    def forward(ctx, student, teacher):
        _require_cuda(student, "output_kd")
        lib = L.lib()
        student = student.contiguous()
        teacher = teacher.detach().contiguous()
        if student.shape != teacher.shape:
            raise RuntimeError("output_kd: student/teacher logits must have the same shape")
        n, c, h, w = student.shape
        with torch.cuda.device_of(student):
            loss = torch.empty((), device=student.device, dtype=torch.float32)
            acc = torch.empty(1, device=student.device, dtype=torch.float64)
            need = ctx.needs_input_grad[0]
            dstudent = torch.empty_like(student) if need else None
            L.check(lib.mdil_kd_fwd_bwd(student.data_ptr(), teacher.data_ptr(), n, c, h, w, loss.data_ptr(),
                                        acc.data_ptr(), _ptr(dstudent), _stream()), "mdil_kd_fwd_bwd")
        ctx.internal = dstudent
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lib = L.lib()
        dstudent = ctx.internal
        if dstudent is None:
            if ctx.needs_input_grad[0]:
                raise RuntimeError("output_kd: the fused student gradient is single-use (scaled in place); a second "
                                   "backward through the same loss (retain_graph=True) is not supported")
            return None, None
        ctx.internal = None
        with torch.cuda.device_of(dstudent):
            g = grad_out.to(dtype=torch.float32).contiguous()
            L.check(lib.mdil_scale_by_device_scalar(dstudent.data_ptr(), dstudent.numel(), g.data_ptr(), _stream()),
                    "mdil_scale_by_device_scalar")
        return dstudent, None


# ======================================================================================= validation
def argmax_confusion(logits: torch.Tensor, labels: Optional[torch.Tensor] = None):
    """outputs.max(1)[1] plus, when labels are given, the C x C confusion matrix conf[gt, pred] that
    iouEval.addBatch (iouEval.py:21-70) reduces to tp/fp/fn."""
    _require_cuda(logits, "argmax_confusion")
    lib = L.lib()
    logits = logits.contiguous()
    n, c, h, w = logits.shape
    with torch.cuda.device_of(logits):
        pred = torch.empty((n, h, w), device=logits.device, dtype=torch.int64)
        conf = None
        if labels is not None:
            if labels.dim() == 4:
                labels = labels[:, 0]
            labels = labels.contiguous()
            conf = torch.zeros((c, c), device=logits.device, dtype=torch.int64)
        L.check(lib.mdil_argmax_confusion(logits.data_ptr(), _ptr(labels), n, c, h, w, pred.data_ptr(), _ptr(conf),
                                          _stream()), "mdil_argmax_confusion")
    return pred, conf

"""ctypes binding of libmdil_b200.so (the C ABI declared in include/mdil_b200.h).

There is NO fallback: if the library is missing or the device is not a B200-class (sm_100) GPU the
product path raises.  Nothing here imports or calls ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmdil_b200.so")

c_float_p = C.c_void_p  # device pointers travel as integers

EXPORTS = [
    "mdil_version", "mdil_last_error_string", "mdil_device_supported", "mdil_nchw_to_nhwc4",
    "mdil_nb1d_packed_floats", "mdil_nb1d_fwd_workspace_bytes", "mdil_nb1d_bwd_workspace_bytes",
    "mdil_nb1d_pack", "mdil_nb1d_fwd", "mdil_nb1d_bwd",
    "mdil_down_packed_floats", "mdil_down_workspace_bytes", "mdil_down_pack", "mdil_down_fwd", "mdil_down_bwd",
    "mdil_up_packed_floats", "mdil_up_workspace_bytes", "mdil_up_pack", "mdil_up_fwd", "mdil_up_bwd",
    "mdil_outconv_fwd", "mdil_outconv_bwd",
    "mdil_ce2d_fwd_bwd", "mdil_ce2d_bwd", "mdil_ce2d_scale", "mdil_kd_fwd_bwd", "mdil_scale_by_device_scalar",
    "mdil_argmax_confusion", "mdil_adam_step", "mdil_launch_count", "mdil_profile_begin", "mdil_profile_end",
    "mdil_cotransform", "mdil_adam_step_dev",
]


class BnParams(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p)]


class Nb1dDesc(C.Structure):
    _fields_ = [("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("dil", C.c_int),
                ("has_adapter", C.c_int), ("train", C.c_int), ("save", C.c_int), ("eps", C.c_float),
                ("momentum", C.c_float)]


class Nb1dWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w31_1", "b31_1", "w13_1", "b13_1", "w31_2", "b31_2", "w13_2", "b13_2",
                                          "wp1", "bp1", "wp2", "bp2")] + [("bn1", BnParams), ("bn2", BnParams)]


class Nb1dSaved(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("a", "p", "c", "s", "stats")]


class Nb1dGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w31_1", "b31_1", "w13_1", "b13_1", "w31_2", "b31_2", "w13_2", "b13_2",
                                          "wp1", "bp1", "wp2", "bp2", "bn1_w", "bn1_b", "bn2_w", "bn2_b")]


class DownDesc(C.Structure):
    _fields_ = [("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
                ("ldin", C.c_int), ("train", C.c_int), ("save", C.c_int), ("eps", C.c_float), ("momentum", C.c_float)]


class UpDesc(C.Structure):
    _fields_ = [("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
                ("train", C.c_int), ("save", C.c_int), ("eps", C.c_float), ("momentum", C.c_float)]


_lib = None
_lock = threading.Lock()


def _declare(lib) -> None:
    vp, i, sz, f = C.c_void_p, C.c_int, C.c_size_t, C.c_float
    P = C.POINTER
    lib.mdil_version.restype = C.c_char_p
    lib.mdil_last_error_string.restype = C.c_char_p
    lib.mdil_device_supported.argtypes = [i]
    lib.mdil_nchw_to_nhwc4.argtypes = [vp, vp, i, i, i, i, vp]
    lib.mdil_nb1d_packed_floats.argtypes = [i]
    lib.mdil_nb1d_packed_floats.restype = sz
    for name in ("mdil_nb1d_fwd_workspace_bytes", "mdil_nb1d_bwd_workspace_bytes"):
        getattr(lib, name).argtypes = [P(Nb1dDesc)]
        getattr(lib, name).restype = sz
    lib.mdil_nb1d_pack.argtypes = [P(Nb1dDesc), P(Nb1dWeights), vp, vp]
    lib.mdil_nb1d_fwd.argtypes = [P(Nb1dDesc), vp, P(Nb1dWeights), vp, vp, vp, P(Nb1dSaved), vp, sz, vp]
    lib.mdil_nb1d_bwd.argtypes = [P(Nb1dDesc), vp, vp, vp, P(Nb1dWeights), vp, vp, P(Nb1dSaved), vp, P(Nb1dGrads),
                                  vp, sz, vp]
    for name in ("mdil_down_packed_floats", "mdil_down_workspace_bytes"):
        getattr(lib, name).argtypes = [P(DownDesc)]
        getattr(lib, name).restype = sz
    lib.mdil_down_pack.argtypes = [P(DownDesc), vp, vp, vp]
    lib.mdil_down_fwd.argtypes = [P(DownDesc), vp, vp, vp, P(BnParams), vp, vp, vp, vp, sz, vp]
    lib.mdil_down_bwd.argtypes = [P(DownDesc), vp, vp, vp, vp, vp, vp, P(BnParams), vp, vp, vp, vp, vp, vp, sz, vp]
    for name in ("mdil_up_packed_floats", "mdil_up_workspace_bytes"):
        getattr(lib, name).argtypes = [P(UpDesc)]
        getattr(lib, name).restype = sz
    lib.mdil_up_pack.argtypes = [P(UpDesc), vp, vp, vp]
    lib.mdil_up_fwd.argtypes = [P(UpDesc), vp, vp, vp, P(BnParams), vp, vp, vp, vp, sz, vp]
    lib.mdil_up_bwd.argtypes = [P(UpDesc), vp, vp, vp, vp, vp, vp, P(BnParams), vp, vp, vp, vp, vp, vp, sz, vp]
    lib.mdil_outconv_fwd.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
    lib.mdil_outconv_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp]
    lib.mdil_ce2d_fwd_bwd.argtypes = [vp, vp, vp, i, i, i, i, vp, vp, vp, vp]
    lib.mdil_ce2d_bwd.argtypes = [vp, vp, vp, i, i, i, i, vp, vp, vp, vp]
    lib.mdil_ce2d_scale.argtypes = [vp, sz, vp, vp, vp]
    lib.mdil_kd_fwd_bwd.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp]
    lib.mdil_scale_by_device_scalar.argtypes = [vp, sz, vp, vp]
    lib.mdil_argmax_confusion.argtypes = [vp, vp, i, i, i, i, vp, vp, vp]
    lib.mdil_profile_end.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int), i]
    lib.mdil_adam_step.argtypes = [vp, vp, vp, vp, sz, f, f, f, f, f, i, f, vp]
    lib.mdil_adam_step_dev.argtypes = [vp, vp, vp, vp, sz, vp, f, f, f, f, f, vp]
    lib.mdil_cotransform.argtypes = [vp, vp, i, i, i, i, i, vp, i, vp, i, vp, vp, vp, i, vp, vp, vp]
    non_int = {"mdil_version", "mdil_last_error_string", "mdil_launch_count", "mdil_nb1d_packed_floats",
               "mdil_nb1d_fwd_workspace_bytes", "mdil_nb1d_bwd_workspace_bytes", "mdil_down_packed_floats",
               "mdil_down_workspace_bytes", "mdil_up_packed_floats", "mdil_up_workspace_bytes"}
    lib.mdil_launch_count.restype = C.c_ulonglong
    for name in EXPORTS:
        if name not in non_int:
            getattr(lib, name).restype = C.c_int


def lib():
    """The loaded library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m mdil_ss_b200.build` "
                        "(mdil_ss_b200 has no CPU or PyTorch fallback for the hot path)")
                handle = C.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mdil_last_error_string()
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def launches() -> int:
    """Kernels launched by this library in this process so far (bench.py's gpu_launches claim)."""
    return int(lib().mdil_launch_count())

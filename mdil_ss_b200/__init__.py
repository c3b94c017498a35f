"""mdil_ss_b200 — B200 (sm_100a) implementation of the MDIL-SS hot path.

Public surface (mirrors the reference's):
    from mdil_ss_b200.erfnet_RA_parallel import Net          # models/erfnet_RA_parallel.py:194
    from mdil_ss_b200.losses import CrossEntropyLoss2d, OutputKD
    from mdil_ss_b200.functional import argmax_confusion
    from mdil_ss_b200.data import DevicePrefetcher            # H2D copies of the next batch overlap the current step
The CUDA library (libmdil_b200.so) is loaded lazily on first use; there is no fallback path.
"""
from . import _lib  # noqa: F401

__all__ = ["erfnet_RA_parallel", "losses", "functional", "parallel", "iou", "data"]
__version__ = "0.1"

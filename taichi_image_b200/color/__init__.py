"""Colour helpers (reference: color/__init__.py).  Host-side (numpy/torch) versions of the small
per-pixel functions the reference defines as ``ti.func``; the device kernels inline their own copies
(csrc/common.cuh ``rgb_gray``)."""
from __future__ import annotations

import numpy as np
import torch

from .yuv_420 import rgb_yuv420_image, yuv420_rgb_image, split_yuv_420, YCrCb_T_bgr, bgr_T_YCrCb  # noqa: F401  (color/__init__.py:4)

GRAY_WEIGHTS = (0.299, 0.587, 0.114)     # color/__init__.py:7-10


def _dot_last(x, w):
    if isinstance(x, torch.Tensor):
        return (x * torch.tensor(w, dtype=x.dtype, device=x.device)).sum(-1)
    return (np.asarray(x) * np.asarray(w, dtype=np.asarray(x).dtype)).sum(-1)


def rgb_gray(rgb):
    """0.299 R + 0.587 G + 0.114 B (color/__init__.py:7-10)"""
    return _dot_last(rgb, GRAY_WEIGHTS)


def bgr_gray(bgr):
    """color/__init__.py:12-15"""
    return _dot_last(bgr, GRAY_WEIGHTS[::-1])

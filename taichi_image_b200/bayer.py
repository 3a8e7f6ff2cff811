"""CFA mosaic and Malvar-He-Cutler demosaic (reference: bayer.py).

``rgb_to_bayer`` / ``bayer_to_rgb`` keep the reference signatures; ``bayer_to_rgb_kernel`` returns a
callable ``f(bayer, out)`` on pre-allocated CUDA tensors (bench/bayer.py:38).  Kernels:
``b200isp_rgb_to_bayer`` / ``b200isp_bayer_to_rgb`` (csrc/demosaic.cu, csrc/stream_engine.cuh).
"""
from __future__ import annotations

import ctypes
import enum
from typing import Optional

import numpy as np
import torch

from . import _lib, types
from .dtypes import as_dtype
from .util import cache


class BayerPattern(enum.Enum):          # bayer.py:75-83
    RGGB = 0
    GRBG = 1
    GBRG = 2
    BGGR = 3

    @property
    def pixel_order(self):
        return pixel_orders[self]


pixel_orders = {                        # bayer.py:85-90
    BayerPattern.RGGB: (0, 1, 1, 2),
    BayerPattern.GRBG: (1, 0, 2, 1),
    BayerPattern.GBRG: (1, 2, 0, 1),
    BayerPattern.BGGR: (2, 1, 1, 0),
}

kernel_patterns = {                     # bayer.py:92-97
    BayerPattern.RGGB: (0, 1, 2, 3),
    BayerPattern.GBRG: (1, 0, 3, 2),
    BayerPattern.GRBG: (2, 3, 0, 1),
    BayerPattern.BGGR: (3, 2, 1, 0),
}

_DIAMOND = [(-2, 0), (-1, -1), (-1, 0), (-1, 1), (0, -2), (0, -1), (0, 0), (0, 1), (0, 2),
            (1, -1), (1, 0), (1, 1), (2, 0)]


def _diamond13(a, bc, def_):
    (b, c), (d, e, f) = bc, def_
    return (a, b, c, b, d, e, f, e, d, b, c, b, a)


def make_bayer_kernels():
    """The four 13-tap site kernels of bayer.py:30-55 as ((d0, d1), (wR, wG, wB)) tuples.  The device
    code evaluates them through shared partial sums (csrc/stream_engine.cuh); this table is the
    documentation / test view of the same weights."""
    g_rb = _diamond13(-2, (0, 4), (-2, 4, 8))
    r_g1 = _diamond13(-2, (-2, 8), (1, 0, 10))
    r_g2 = _diamond13(1, (-2, 0), (-2, 8, 10))
    rb_br = _diamond13(-3, (4, 0), (-3, 0, 12))
    ident = _diamond13(0, (0, 0), (0, 0, 16))
    sites = ((ident, g_rb, rb_br), (r_g1, ident, r_g2), (r_g2, ident, r_g1), (rb_br, g_rb, ident))
    return tuple(tuple(zip(_DIAMOND, zip(*s))) for s in sites)


bayer_kernels = make_bayer_kernels()


def _ccm_arg(correct_colors):
    if correct_colors is None:
        return None
    m = np.asarray(correct_colors, dtype=np.float64).reshape(-1)
    assert m.size == 9, "correct_colors must be a 3x3 matrix"
    return (ctypes.c_float * 9)(*[float(x) for x in m])


@cache
def bayer_to_rgb_kernel(pattern: BayerPattern, correct_colors: Optional[tuple] = None, in_dtype=types.u8, out_dtype=None):
    """bayer.py:179-190 -- returns f(bayer, out) on CUDA tensors (H, W) -> (H, W, 3)."""
    in_dtype = as_dtype(in_dtype)
    out_dtype = in_dtype if out_dtype is None else as_dtype(out_dtype)
    ccm = _ccm_arg(correct_colors)

    def f(bayer: torch.Tensor, out: torch.Tensor):
        _lib.require_cuda(bayer, "bayer_to_rgb")
        _lib.require_cuda(out, "bayer_to_rgb")
        assert as_dtype(bayer.dtype) is in_dtype and as_dtype(out.dtype) is out_dtype
        assert bayer.is_contiguous() and out.is_contiguous()
        h, w = bayer.shape
        assert tuple(out.shape) == (h, w, 3)
        with torch.cuda.device(bayer.device):
            _lib.check(_lib.lib.b200isp_bayer_to_rgb(bayer.data_ptr(), in_dtype.code, out.data_ptr(), out_dtype.code,
                                                     h, w, pattern.value, ccm, _lib.stream_ptr(bayer.device)), "bayer_to_rgb")
    return f


def rgb_to_bayer(image, pattern: BayerPattern = BayerPattern.RGGB):
    """bayer.py:193-198"""
    assert image.ndim == 3 and image.shape[2] == 3, "image must be RGB"
    dev, restore = types.to_device(image)
    h, w = dev.shape[:2]
    dt = types.ti_type(image)
    # the reference leaves an odd trailing row/column uninitialised; zero it here
    bayer = (torch.empty if (h % 2 == 0 and w % 2 == 0) else torch.zeros)((h, w), dtype=dev.dtype, device=dev.device)
    if h and w:
        with torch.cuda.device(dev.device):
            _lib.check(_lib.lib.b200isp_rgb_to_bayer(dev.data_ptr(), bayer.data_ptr(), dt.code, h, w, pattern.value,
                                                     _lib.stream_ptr(dev.device)), "rgb_to_bayer")
    return restore(bayer)


def bayer_to_rgb(bayer, pattern: BayerPattern = BayerPattern.RGGB, correct_colors: Optional[np.ndarray] = None, dtype=None,
                 method: str = "malvar"):
    """bayer.py:202-219.  ``method="bilinear"`` (EXTENSION, no reference counterpart): 3x3 bilinear interpolation with
    the same border rule (mean of the in-bounds neighbours of each colour), ``b200isp_bayer_to_rgb_bilinear``."""
    assert method in ("malvar", "bilinear")
    assert bayer.ndim == 2, "image must be mono bayer"
    assert bayer.shape[0] % 2 == 0 and bayer.shape[1] % 2 == 0, "image must be even size"
    dev, restore = types.to_device(bayer)
    in_dtype = types.ti_type(bayer)
    out_dtype = in_dtype if dtype is None else as_dtype(dtype)
    rgb = torch.empty(tuple(dev.shape) + (3,), dtype=out_dtype.torch, device=dev.device)
    if correct_colors is not None:
        correct_colors = tuple(np.asarray(correct_colors, dtype=np.float64).flatten().tolist())
    if dev.numel() and method == "bilinear":
        ccm = None if correct_colors is None else (ctypes.c_float * 9)(*correct_colors)
        with torch.cuda.device(dev.device):
            _lib.check(_lib.lib.b200isp_bayer_to_rgb_bilinear(dev.data_ptr(), in_dtype.code, rgb.data_ptr(), out_dtype.code,
                                                              dev.shape[0], dev.shape[1], pattern.value, ccm,
                                                              _lib.stream_ptr(dev.device)), "bayer_to_rgb_bilinear")
    elif dev.numel():
        bayer_to_rgb_kernel(pattern, correct_colors, in_dtype, out_dtype)(dev, rgb)
    return restore(rgb)

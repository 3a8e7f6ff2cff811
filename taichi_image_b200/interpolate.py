"""Bilinear / area resize and rotate-flip transforms (reference: interpolate.py).

Kernels: ``b200isp_resize_bilinear``, ``b200isp_resize_area`` (extension), ``b200isp_transform``
(csrc/resize.cu).  Deviations from the reference, all listed in SURVEY 2.5: ``resize_bilinear`` without
an explicit scale uses per-axis scales (Q9), ``transverse`` is the real anti-transpose (Q10),
``none`` returns the input without a copy (Q12), input and output dtypes may differ (Q8).
"""
from __future__ import annotations

from enum import Enum

import torch

from . import _lib, types
from .dtypes import as_dtype


class ImageTransform(Enum):              # interpolate.py:9-17
    none = 'none'
    rotate_90 = 'rotate_90'
    rotate_180 = 'rotate_180'
    rotate_270 = 'rotate_270'
    transpose = 'transpose'
    flip_horiz = 'flip_horiz'
    flip_vert = 'flip_vert'
    transverse = 'transverse'


_TRANSFORM_CODE = {ImageTransform.none: 0, ImageTransform.rotate_90: 1, ImageTransform.rotate_180: 2,
                   ImageTransform.rotate_270: 3, ImageTransform.transpose: 4, ImageTransform.flip_horiz: 5,
                   ImageTransform.flip_vert: 6, ImageTransform.transverse: 7}
_SWAPS = (ImageTransform.rotate_90, ImageTransform.rotate_270, ImageTransform.transpose, ImageTransform.transverse)


def transformed_size(size, transform: ImageTransform):
    """interpolate.py:112-117 (transverse also swaps: SURVEY Q10)"""
    a, b = size
    return (b, a) if transform in _SWAPS else (a, b)


def transform(src, transform: ImageTransform):
    """interpolate.py:119-125.  rotate_90 is clockwise: dst[r, c] = src[H-1-c, r]."""
    assert src.ndim == 3 and src.shape[2] == 3, "image must be (H, W, 3)"
    if transform == ImageTransform.none:
        return src
    dev, restore = types.to_device(src)
    h, w = dev.shape[:2]
    hd, wd = transformed_size((h, w), transform)
    dst = torch.empty((hd, wd, 3), dtype=dev.dtype, device=dev.device)
    if dev.numel():
        with torch.cuda.device(dev.device):
            _lib.check(_lib.lib.b200isp_transform(dev.data_ptr(), dst.data_ptr(), types.ti_type(src).code, h, w,
                                                  _TRANSFORM_CODE[transform], _lib.stream_ptr(dev.device)), "transform")
    return restore(dst)


def _scale_pair(scale, size, src_shape):
    if scale is None:
        return (size[1] / src_shape[0], size[0] / src_shape[1])     # (row, col) = (h_out/H, w_out/W)
    if isinstance(scale, (int, float)):
        return (float(scale), float(scale))
    s = [float(x) for x in scale]
    assert len(s) == 2
    return (s[0], s[1])


def resize_bilinear(src, size, scale=None, dtype=None):
    """interpolate.py:128-139.  ``size`` = (width, height); ``scale`` scalar or (row, col) pair."""
    assert src.ndim == 3 and src.shape[2] == 3, "image must be (H, W, 3)"
    in_dtype = types.ti_type(src)
    dtype = in_dtype if dtype is None else as_dtype(dtype)
    dev, restore = types.to_device(src)
    w_out, h_out = int(size[0]), int(size[1])
    sr, sc = _scale_pair(scale, (w_out, h_out), dev.shape[:2])
    dst = torch.empty((h_out, w_out, 3), dtype=dtype.torch, device=dev.device)
    if dst.numel():
        with torch.cuda.device(dev.device):
            _lib.check(_lib.lib.b200isp_resize_bilinear(dev.data_ptr(), in_dtype.code, dev.shape[0], dev.shape[1],
                                                        dst.data_ptr(), dtype.code, h_out, w_out, sr, sc,
                                                        _lib.stream_ptr(dev.device)), "resize_bilinear")
    return restore(dst)


def resize_width(src, width: int, dtype=None):
    """interpolate.py:141-145"""
    h, w = src.shape[:2]
    scale = width / w
    return resize_bilinear(src, (width, int(h * scale)), scale, dtype)


def scale_bilinear(src, scale, dtype=None):
    """interpolate.py:147-151"""
    h, w = src.shape[:2]
    return resize_bilinear(src, (int(w * scale), int(h * scale)), scale, dtype=dtype)


def resize_area(src, size, dtype=None):
    """EXTENSION (north_star "bilinear/area resize"; the reference has no area filter): exact box
    filter over the source footprint of every output pixel.  ``size`` = (width, height)."""
    assert src.ndim == 3 and src.shape[2] == 3, "image must be (H, W, 3)"
    in_dtype = types.ti_type(src)
    dtype = in_dtype if dtype is None else as_dtype(dtype)
    dev, restore = types.to_device(src)
    w_out, h_out = int(size[0]), int(size[1])
    dst = torch.empty((h_out, w_out, 3), dtype=dtype.torch, device=dev.device)
    if dst.numel():
        with torch.cuda.device(dev.device):
            _lib.check(_lib.lib.b200isp_resize_area(dev.data_ptr(), in_dtype.code, dev.shape[0], dev.shape[1],
                                                    dst.data_ptr(), dtype.code, h_out, w_out,
                                                    _lib.stream_ptr(dev.device)), "resize_area")
    return restore(dst)

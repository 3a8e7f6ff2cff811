"""Planar YUV 4:2:0 encode / decode (reference: color/yuv_420.py).

Same functions and layout as the reference: ``rgb_yuv420_image(src, dtype)`` returns one ``(3H/2, W)`` array -- H rows
of Y followed by the two ``(H/2, W/2)`` chroma planes (``split_yuv_420``) -- and ``yuv420_rgb_image`` inverts it.  The
matrices are the reference's module-level constants, including its convention of applying them to the BGR-swizzled
pixel (yuv_420.py:24-34).  Kernels: ``b200isp_rgb_yuv420`` / ``b200isp_yuv420_rgb`` (csrc/yuv420.cu).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib, types
from ..dtypes import as_dtype

YCrCb_T_bgr = np.array([[0.299, 0.587, 0.114],          # yuv_420.py:12-16
                        [-0.168736, -0.331264, 0.5],
                        [0.5, -0.418688, -0.081312]], dtype=np.float64)
# yuv_420.py:18: inverse of the f32 matrix, evaluated once on the host
bgr_T_YCrCb = np.linalg.inv(YCrCb_T_bgr.astype(np.float32).astype(np.float64)).astype(np.float32)


def _mat9(m):
    return (ctypes.c_float * 9)(*[float(v) for v in np.asarray(m, dtype=np.float32).reshape(-1)])


def split_yuv_420(yuv):
    """yuv_420.py:95-101 -> (y, uv[2, H/2, W/2], (width, height))"""
    height = yuv.shape[0] * 2 // 3
    width = yuv.shape[1]
    y = yuv[:height]
    uv = yuv[height:].reshape(2, height // 2, width // 2)
    return y, uv, (width, height)


def rgb_yuv420_image(src, dtype=None):
    """yuv_420.py:104-118"""
    assert src.ndim == 3 and src.shape[2] == 3, "image must be HxWx3"
    out_dtype = types.ti_type(src) if dtype is None else as_dtype(dtype)
    t, restore = types.to_device(src)
    height, width, _ = t.shape
    assert height % 2 == 0 and width % 2 == 0, "image must be even size"
    yuv = torch.zeros(((height * 3) // 2, width), dtype=out_dtype.torch, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib.b200isp_rgb_yuv420(t.data_ptr(), types.ti_type(t).code, yuv.data_ptr(), out_dtype.code, height, width,
                                               _mat9(YCrCb_T_bgr), _lib.stream_ptr(t.device)), "rgb_yuv420")
    return restore(yuv)


def yuv420_rgb_image(yuv, dtype=None):
    """yuv_420.py:120-131"""
    assert yuv.ndim == 2 and yuv.shape[0] % 3 == 0, "yuv420 image must be (3H/2, W)"
    out_dtype = types.ti_type(yuv) if dtype is None else as_dtype(dtype)
    t, restore = types.to_device(yuv)
    height, width = t.shape[0] * 2 // 3, t.shape[1]
    assert height % 2 == 0 and width % 2 == 0, "image must be even size"
    rgb = torch.zeros((height, width, 3), dtype=out_dtype.torch, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib.b200isp_yuv420_rgb(t.data_ptr(), types.ti_type(t).code, rgb.data_ptr(), out_dtype.code, height, width,
                                               _mat9(bgr_T_YCrCb), _lib.stream_ptr(t.device)), "yuv420_rgb")
    return restore(rgb)

"""Stand-alone (per image, stateless) tone mapping (reference: tonemap.py).

``tonemap_linear`` = global min/max normalise + gamma (tonemap.py:26-46);
``tonemap_reinhard`` = the five dependent passes of tonemap.py:134-168.  Everything runs on the
device with no host synchronisation: ``b200isp_bounds`` + ``b200isp_linear`` /
``b200isp_reinhard_standalone`` (csrc/tonemap.cu).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, types
from .dtypes import as_dtype, u8
from .util import Bounds


@dataclass
class Metering:                          # tonemap.py:53-63 (host mirror of the 7-float record)
    log_bounds: Bounds
    log_mean: float
    gray_mean: float
    rgb_mean: np.ndarray

    def to_vec(self):
        return metering_to_np(self)


def metering_to_np(x: Metering):         # tonemap.py:66-68
    return np.array([x.log_bounds.min, x.log_bounds.max, x.log_mean, x.gray_mean, *x.rgb_mean])


def metering_from_np(x: np.ndarray):     # tonemap.py:70-72
    return Metering(Bounds(x[0], x[1]), x[2], x[3], np.array([x[4], x[5], x[6]]))


def _check_image(src):
    assert src.ndim == 3 and src.shape[2] == 3, "image must be (H, W, 3)"


def tonemap_linear(src, gamma=1.0, dtype=u8):
    """tonemap.py:41-46"""
    _check_image(src)
    dtype = as_dtype(dtype)
    dev, restore = types.to_device(src)
    in_dtype = types.ti_type(src)
    out = torch.empty(dev.shape, dtype=dtype.torch, device=dev.device)
    n = dev.numel()
    if n:
        with torch.cuda.device(dev.device):
            ws = _lib.workspace(dev.device)
            s = _lib.stream_ptr(dev.device)
            bounds = torch.empty(2, dtype=torch.float32, device=dev.device)
            _lib.check(_lib.lib.b200isp_bounds(dev.data_ptr(), in_dtype.code, n, bounds.data_ptr(), ws.data_ptr(), s), "bounds")
            _lib.check(_lib.lib.b200isp_linear(dev.data_ptr(), in_dtype.code, out.data_ptr(), dtype.code, n,
                                               bounds.data_ptr(), float(gamma), s), "linear")
    return restore(out)


def tonemap_reinhard(src, gamma=1.0, intensity=1.0, light_adapt=1.0, color_adapt=0.0, dtype=u8):
    """tonemap.py:160-168"""
    _check_image(src)
    dtype = as_dtype(dtype)
    dev, restore = types.to_device(src)
    in_dtype = types.ti_type(src)
    out = torch.empty(dev.shape, dtype=dtype.torch, device=dev.device)
    n_px = dev.shape[0] * dev.shape[1]
    if n_px:
        with torch.cuda.device(dev.device):
            # no temp image (tonemap.py:163): the passes recompute from the source (csrc/tonemap.cu)
            _lib.check(_lib.lib.b200isp_reinhard_standalone(
                dev.data_ptr(), in_dtype.code, None, out.data_ptr(), dtype.code, n_px, float(gamma),
                float(intensity), float(light_adapt), float(color_adapt), _lib.workspace(dev.device).data_ptr(),
                _lib.stream_ptr(dev.device)), "tonemap_reinhard")
    return restore(out)

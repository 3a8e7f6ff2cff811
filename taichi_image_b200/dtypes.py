"""Dtype sentinels with the spelling callers of the reference use (``ti.u8``, ``ti.f16`` ...).

The reference passes Taichi dtype objects as ``dtype=`` arguments (packed.py:188, tonemap.py:41,
camera_isp.py:422-423).  There is no Taichi here, so this module supplies objects of the same names;
every API also accepts ``torch.dtype``, ``numpy.dtype`` / numpy scalar types and the strings
``"u8" "u16" "i16" "f16" "f32"`` (or ``"uint8"`` ...).
"""
from __future__ import annotations

import numpy as np
import torch


class DType:
    __slots__ = ("name", "code", "np", "torch", "scale", "itemsize")

    def __init__(self, name, code, np_dtype, torch_dtype, scale):
        self.name, self.code, self.np, self.torch, self.scale = name, code, np.dtype(np_dtype), torch_dtype, scale
        self.itemsize = self.np.itemsize

    def __repr__(self):
        return f"taichi_image_b200.{self.name}"


# codes = b200isp_dtype (include/b200isp.h); scale = types.py:12-18
u8 = uint8 = DType("u8", 0, np.uint8, torch.uint8, 255)
u16 = uint16 = DType("u16", 1, np.uint16, torch.uint16, 65535)
i16 = int16 = DType("i16", 2, np.int16, torch.int16, 32767)
f16 = float16 = DType("f16", 3, np.float16, torch.float16, 1.0)
f32 = float32 = DType("f32", 4, np.float32, torch.float32, 1.0)

ALL = (u8, u16, i16, f16, f32)
_BY_NAME = {d.name: d for d in ALL}
_BY_NAME.update({str(d.np): d for d in ALL})
_BY_TORCH = {d.torch: d for d in ALL}
_BY_NP = {d.np: d for d in ALL}


def as_dtype(x) -> DType:
    """Normalise any accepted dtype spelling to a DType sentinel."""
    if isinstance(x, DType):
        return x
    if isinstance(x, torch.dtype):
        if x in _BY_TORCH:
            return _BY_TORCH[x]
        raise KeyError(f"unsupported torch dtype {x}")
    if isinstance(x, str):
        key = x.replace("torch.", "")
        if key in _BY_NAME:
            return _BY_NAME[key]
        raise KeyError(f"unsupported dtype {x!r}")
    try:
        return _BY_NP[np.dtype(x)]
    except (TypeError, KeyError):
        raise KeyError(f"unsupported dtype {x!r}") from None

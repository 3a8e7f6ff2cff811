"""dtype tables and "allocate like the input" helpers (reference: types.py:12-91).

Differences from the reference: ``torch.uint16`` is accepted and returned.
numpy arrays and CPU tensors are accepted everywhere: they are copied to the current CUDA device,
processed there and copied back (what Taichi's ndarray interop does for host arrays) -- there is no
CPU execution path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import dtypes
from .dtypes import DType, as_dtype, u8, u16, i16, f16, f32  # noqa: F401

scale_factor = {d: d.scale for d in dtypes.ALL}            # types.py:12-18
ti_to_np = {d: d.np.type for d in dtypes.ALL}              # types.py:21-27
ti_to_torch = {d: d.torch for d in dtypes.ALL}             # types.py:29-33 (+ u16, i16)
type_to_ti = {str(d.np): d for d in dtypes.ALL}            # types.py:36-42
torch_to_ti = {str(d.torch): d for d in dtypes.ALL}        # types.py:44-49 (+ torch.uint16)


def ti_type(in_arr) -> DType:
    """types.py:51-57"""
    if isinstance(in_arr, np.ndarray):
        return type_to_ti[str(in_arr.dtype)]
    if isinstance(in_arr, torch.Tensor):
        return torch_to_ti[str(in_arr.dtype)]
    raise ValueError(f"Unsupported input type {type(in_arr)}")


def empty_like(in_arr, shape=None, dtype=None):
    """types.py:59-67"""
    shape = tuple(in_arr.shape if shape is None else shape)
    dtype = ti_type(in_arr) if dtype is None else as_dtype(dtype)
    if isinstance(in_arr, np.ndarray):
        return np.empty(shape, dtype.np)
    if isinstance(in_arr, torch.Tensor):
        return torch.empty(shape, dtype=dtype.torch, device=in_arr.device)
    raise ValueError(f"Unsupported input type {type(in_arr)}")


def zeros_like(in_arr, shape=None, dtype=None):
    """types.py:81-91"""
    out = empty_like(in_arr, shape, dtype)
    if isinstance(out, np.ndarray):
        out[...] = 0
    else:
        out.zero_()
    return out


# ---------------------------------------------------------------- host <-> device staging
def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("taichi_image_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(arr):
    """Return (contiguous CUDA tensor, restore) where restore(t) converts a result back to the caller's
    container kind (numpy array / CPU tensor / CUDA tensor)."""
    if isinstance(arr, np.ndarray):
        ti_type(arr)   # KeyError for unsupported dtypes, like the reference
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(default_device())
        return t, lambda r: r.cpu().numpy()
    if isinstance(arr, torch.Tensor):
        ti_type(arr)
        if arr.is_cuda:
            return arr.contiguous(), lambda r: r
        src_device = arr.device
        return arr.contiguous().to(default_device()), lambda r: r.to(src_device)
    raise ValueError(f"Unsupported input type {type(arr)}")

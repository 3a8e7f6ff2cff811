"""12-bit pack / unpack and raw-16 decode (reference: packed.py), plus the 10-bit MIPI RAW10 extension.

Same entry points as the reference: ``encode12``, ``decode12``, ``decode16`` on numpy arrays or torch
tensors, and the kernel factories ``encode12_kernel`` / ``decode12_kernel`` / ``decode16_kernel``
returning callables ``f(flat_in, flat_out)`` on pre-allocated flat CUDA tensors (camera_isp.py:335-339).
The work is done by ``b200isp_encode12`` / ``b200isp_decode12`` / ``b200isp_decode16`` (csrc/pack.cu).
"""
from __future__ import annotations

import torch

from . import _lib, types
from .dtypes import as_dtype, u8, u16
from .util import cache


@cache
def encode12_kernel(in_type, scaled=False, ids_format=False):
    """packed.py:59-89 -- returns f(values, encoded) on flat CUDA tensors."""
    in_type = as_dtype(in_type)

    def f(values: torch.Tensor, encoded: torch.Tensor):
        _lib.require_cuda(values, "encode12")
        _lib.require_cuda(encoded, "encode12")
        assert as_dtype(values.dtype) is in_type and encoded.dtype == torch.uint8
        assert values.is_contiguous() and encoded.is_contiguous()
        n = values.numel()
        assert encoded.numel() * 2 == n * 3, "encoded must hold 3 bytes per 2 values"
        with torch.cuda.device(values.device):
            _lib.check(_lib.lib.b200isp_encode12(values.data_ptr(), in_type.code, n, encoded.data_ptr(),
                                                 int(scaled), int(ids_format), _lib.stream_ptr(values.device)), "encode12")
    return f


@cache
def decode12_kernel(out_type, scaled=False, ids_format=False):
    """packed.py:122-131 -- returns k(encoded, out) on flat CUDA tensors."""
    out_type = as_dtype(out_type)

    def k(encoded: torch.Tensor, out: torch.Tensor):
        _lib.require_cuda(encoded, "decode12")
        _lib.require_cuda(out, "decode12")
        assert encoded.dtype == torch.uint8 and as_dtype(out.dtype) is out_type
        assert encoded.is_contiguous() and out.is_contiguous()
        n = out.numel()
        assert encoded.numel() * 2 == n * 3, "encoded must hold 3 bytes per 2 values"
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib.b200isp_decode12(encoded.data_ptr(), n, out.data_ptr(), out_type.code,
                                                 int(scaled), int(ids_format), _lib.stream_ptr(out.device)), "decode12")
    return k


@cache
def decode16_kernel(out_type, scaled=False):
    """packed.py:163-172 -- returns k(encoded, out) on flat CUDA tensors."""
    out_type = as_dtype(out_type)

    def k(encoded: torch.Tensor, out: torch.Tensor):
        _lib.require_cuda(encoded, "decode16")
        _lib.require_cuda(out, "decode16")
        assert encoded.dtype == torch.uint8 and as_dtype(out.dtype) is out_type
        assert encoded.is_contiguous() and out.is_contiguous()
        n = out.numel()
        assert encoded.numel() == 2 * n
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib.b200isp_decode16(encoded.data_ptr(), n, out.data_ptr(), out_type.code,
                                                 int(scaled), _lib.stream_ptr(out.device)), "decode16")
    return k


def encode12(values, scaled=False, ids_format=False):
    """packed.py:176-185"""
    shape = tuple(values.shape)
    assert shape[-1] % 2 == 0, f"last dimension must be even for 12-bit encoding got: {shape}"
    dev, restore = types.to_device(values)
    flat = dev.reshape(-1)
    encoded = torch.empty((flat.shape[0] * 3) // 2, dtype=torch.uint8, device=flat.device)
    if flat.numel():
        encode12_kernel(types.ti_type(values), scaled=scaled, ids_format=ids_format)(flat, encoded)
    return restore(encoded.reshape(shape[:-1] + (shape[-1] * 3 // 2,)))


def decode12(values, dtype=u16, scaled=False, ids_format=False):
    """packed.py:188-198"""
    shape = tuple(values.shape)
    assert types.ti_type(values) is u8
    assert shape[-1] % 3 == 0, f"last dimension must be a factor of 3 for 12-bit decoding got: {shape}"
    dtype = as_dtype(dtype)
    dev, restore = types.to_device(values)
    flat = dev.reshape(-1)
    decoded = torch.empty((flat.shape[0] * 2) // 3, dtype=dtype.torch, device=flat.device)
    if flat.numel():
        decode12_kernel(dtype, scaled=scaled, ids_format=ids_format)(flat, decoded)
    return restore(decoded.reshape(shape[:-1] + (shape[-1] * 2 // 3,)))


def repack12_ids(values, out=None):
    """EXTENSION: IDS-layout 12-bit data -> the standard layout, i.e. ``encode12(decode12(values, ids_format=True))``
    (packed.py:36-44 followed by :12-20) without leaving the packed domain.  Used by ``camera_isp`` so that IDS frames
    take the fused sweep.  ``out``: optional preallocated uint8 tensor of the same shape (CUDA)."""
    shape = tuple(values.shape)
    assert types.ti_type(values) is u8
    assert shape[-1] % 3 == 0, f"last dimension must be a factor of 3 for 12-bit data got: {shape}"
    dev, restore = types.to_device(values)
    dev = dev.contiguous()
    res = torch.empty_like(dev) if out is None else out
    assert res.dtype == torch.uint8 and tuple(res.shape) == shape and res.is_contiguous() and res.device == dev.device
    if dev.numel():
        with torch.cuda.device(dev.device):
            _lib.check(_lib.lib.b200isp_repack12_ids(dev.data_ptr(), res.data_ptr(), dev.numel(), _lib.stream_ptr(dev.device)),
                       "repack12_ids")
    return res if out is not None else restore(res)


def decode16(values, dtype=u16, scaled=False, ids_format=False):
    """packed.py:200-210.  ``ids_format`` is accepted and ignored: the reference forwards it to a
    kernel factory that has no such parameter and always raises."""
    shape = tuple(values.shape)
    assert types.ti_type(values) is u8
    assert shape[-1] % 2 == 0, f"last dimension must be a factor of 2 for 16-bit decoding got: {shape}"
    dtype = as_dtype(dtype)
    dev, restore = types.to_device(values)
    flat = dev.reshape(-1)
    decoded = torch.empty(flat.shape[0] // 2, dtype=dtype.torch, device=flat.device)
    if flat.numel():
        decode16_kernel(dtype, scaled=scaled)(flat, decoded)
    return restore(decoded.reshape(shape[:-1] + (shape[-1] // 2,)))


# ---------------------------------------------------------------------------------------------------------------------
# EXTENSION (SURVEY 8f-4 "10-bit packed"; no reference counterpart): MIPI CSI-2 RAW10 -- 5 bytes <-> 4 pixels, bytes 0..3 =
# bits 9..2 of pixels 0..3, byte 4 = their bits 1..0 (pixel 0 in the lowest bit pair).  Same calling conventions as the
# 12-bit functions above, 1023 in place of 4095 (csrc/pack.cu: b200isp_decode10 / b200isp_encode10).
@cache
def encode10_kernel(in_type, scaled=False):
    """returns f(values, encoded) on flat CUDA tensors (the shape of ``encode12_kernel``)"""
    in_type = as_dtype(in_type)

    def f(values: torch.Tensor, encoded: torch.Tensor):
        _lib.require_cuda(values, "encode10")
        _lib.require_cuda(encoded, "encode10")
        assert as_dtype(values.dtype) is in_type and encoded.dtype == torch.uint8
        assert values.is_contiguous() and encoded.is_contiguous()
        n = values.numel()
        assert encoded.numel() * 4 == n * 5, "encoded must hold 5 bytes per 4 values"
        with torch.cuda.device(values.device):
            _lib.check(_lib.lib.b200isp_encode10(values.data_ptr(), in_type.code, n, encoded.data_ptr(), int(scaled),
                                                 _lib.stream_ptr(values.device)), "encode10")
    return f


@cache
def decode10_kernel(out_type, scaled=False):
    """returns k(encoded, out) on flat CUDA tensors (the shape of ``decode12_kernel``)"""
    out_type = as_dtype(out_type)

    def k(encoded: torch.Tensor, out: torch.Tensor):
        _lib.require_cuda(encoded, "decode10")
        _lib.require_cuda(out, "decode10")
        assert encoded.dtype == torch.uint8 and as_dtype(out.dtype) is out_type
        assert encoded.is_contiguous() and out.is_contiguous()
        n = out.numel()
        assert encoded.numel() * 4 == n * 5, "encoded must hold 5 bytes per 4 values"
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib.b200isp_decode10(encoded.data_ptr(), n, out.data_ptr(), out_type.code, int(scaled),
                                                 _lib.stream_ptr(out.device)), "decode10")
    return k


def encode10(values, scaled=False):
    shape = tuple(values.shape)
    assert shape[-1] % 4 == 0, f"last dimension must be a multiple of 4 for 10-bit encoding got: {shape}"
    dev, restore = types.to_device(values)
    flat = dev.reshape(-1)
    encoded = torch.empty((flat.shape[0] * 5) // 4, dtype=torch.uint8, device=flat.device)
    if flat.numel():
        encode10_kernel(types.ti_type(values), scaled=scaled)(flat, encoded)
    return restore(encoded.reshape(shape[:-1] + (shape[-1] * 5 // 4,)))


def decode10(values, dtype=u16, scaled=False):
    shape = tuple(values.shape)
    assert types.ti_type(values) is u8
    assert shape[-1] % 5 == 0, f"last dimension must be a factor of 5 for 10-bit decoding got: {shape}"
    dtype = as_dtype(dtype)
    dev, restore = types.to_device(values)
    flat = dev.reshape(-1)
    decoded = torch.empty((flat.shape[0] * 4) // 5, dtype=dtype.torch, device=flat.device)
    if flat.numel():
        decode10_kernel(dtype, scaled=scaled)(flat, decoded)
    return restore(decoded.reshape(shape[:-1] + (shape[-1] * 4 // 5,)))

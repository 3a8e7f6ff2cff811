"""GPU parity: color/yuv_420.py (planar YUV 4:2:0) against the golden vectors of the reference source and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import isp_oracle as O
from tests.util import rng, random_plane, to_cuda, to_np

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _defined(yuv, name):
    # negative components are undefined behaviour in the reference's integer cast (no lower clamp, yuv_420.py:88)
    return (O.yuv420_rgb_unit(yuv) >= 0) | (name in ("f16", "f32"))


@pytest.mark.parametrize("name", ["u8", "u16", "f16", "f32"])
def test_yuv420_golden(cuda, name):
    from taichi_image_b200 import color
    g = np.load(os.path.join(GOLD, "color.npz"))
    enc = to_np(color.rgb_yuv420_image(to_cuda(g[f"rgb_{name}"])))
    assert enc.dtype == g[f"yuv_{name}"].dtype and np.array_equal(enc, g[f"yuv_{name}"])          # bit-exact
    dec = to_np(color.yuv420_rgb_image(to_cuda(g[f"yuv_{name}"])))
    ok = _defined(g[f"yuv_{name}"], name)
    assert dec.shape == g[f"rgb_back_{name}"].shape and np.array_equal(dec[ok], g[f"rgb_back_{name}"][ok])


def test_yuv420_dtype_conversion_and_numpy(cuda):
    from taichi_image_b200 import color, f32, u8
    g = np.load(os.path.join(GOLD, "color.npz"))
    assert np.array_equal(color.rgb_yuv420_image(g["rgb_u8"], f32), g["yuv_u8_to_f32"])     # numpy in -> numpy out
    assert np.array_equal(color.rgb_yuv420_image(g["rgb_f32"], u8), g["yuv_f32_to_u8"])


@pytest.mark.parametrize("name", ["u8", "u16", "f16", "f32"])
@pytest.mark.parametrize("shape", [(34, 72), (6, 20), (2, 8), (66, 1032)])     # vector kernels (W % 8 == 0) and element-wise
def test_yuv420_shapes_vs_oracle(cuda, name, shape):
    from taichi_image_b200 import color
    rgb = random_plane(rng(81), shape + (3,), name)
    ref = O.rgb_yuv420(rgb)
    enc = to_np(color.rgb_yuv420_image(to_cuda(rgb)))
    assert np.array_equal(np.ascontiguousarray(enc).view(np.uint8), np.ascontiguousarray(ref).view(np.uint8))
    dec = to_np(color.yuv420_rgb_image(to_cuda(ref)))
    ok = _defined(ref, name)
    assert np.array_equal(dec[ok], O.yuv420_rgb(ref)[ok])


@pytest.mark.parametrize("name", ["u8", "u16", "f32"])
def test_yuv420_large_vs_oracle(cuda, name):
    from taichi_image_b200 import color
    r = rng(80)
    rgb = random_plane(r, (128, 192, 3), name)
    enc = to_np(color.rgb_yuv420_image(to_cuda(rgb)))
    ref = O.rgb_yuv420(rgb)
    assert np.array_equal(enc, ref)
    y, uv, (w, h) = color.split_yuv_420(enc)
    assert (w, h) == (192, 128) and y.shape == (128, 192) and uv.shape == (2, 64, 96)
    dec = to_np(color.yuv420_rgb_image(to_cuda(ref)))
    ok = _defined(ref, name)
    assert np.array_equal(dec[ok], O.yuv420_rgb(ref)[ok])
    # grey images survive the round trip (chroma 0.5 exactly representable only approximately: 1 LSB)
    grey = np.repeat(random_plane(r, (32, 48, 1), name), 3, axis=2)
    back = to_np(color.yuv420_rgb_image(color.rgb_yuv420_image(to_cuda(grey))))
    tol = {"u8": 2, "u16": 300, "f32": 5e-3}[name]
    assert np.abs(back.astype(np.float64) - grey.astype(np.float64)).max() <= tol


# ---------------------------------------------------------------- YUV 4:2:0 straight from the fused path
@pytest.mark.parametrize("pattern", ["RGGB", "GBRG"])
@pytest.mark.parametrize("gamma", [1.0, 0.7])
@pytest.mark.parametrize("shape", [(36, 72), (20, 520)])
def test_fused_yuv420_output(cuda, pattern, gamma, shape):
    """process_packed12(yuv420=True) == rgb_yuv420_image(process_packed12(...)) bit for bit, and the RGB it is made
    from is within 1 LSB of the oracle ISP"""
    from taichi_image_b200 import bayer, camera_isp, color
    from tests.util import packed_frame, assert_close_int
    r = rng(82)
    a = camera_isp.Camera16(bayer.BayerPattern[pattern])
    b = camera_isp.Camera16(bayer.BayerPattern[pattern])
    ref = O.ISP("f16", pattern)
    for step in range(2):
        fr = [packed_frame(r, *shape, pattern) for _ in range(3)]
        cu = [to_cuda(f) for f in fr]
        yuv = a.process_packed12(cu, tonemap="reinhard", gamma=gamma, intensity=2.0, light_adapt=0.8, yuv420=True)
        rgb = b.process_packed12(cu, tonemap="reinhard", gamma=gamma, intensity=2.0, light_adapt=0.8)
        exp = ref.tonemap_reinhard([ref.load_packed12(f) for f in fr], gamma=gamma, intensity=2.0, light_adapt=0.8)
        for y, g, e in zip(yuv, rgb, exp):
            assert tuple(y.shape) == (shape[0] * 3 // 2, shape[1]) and y.dtype == g.dtype
            assert np.array_equal(to_np(y), to_np(color.rgb_yuv420_image(g))), f"step {step}"
            assert np.array_equal(to_np(y), O.rgb_yuv420(to_np(g)))
            assert_close_int(to_np(g), e, 1, "rgb behind the yuv")


@pytest.mark.parametrize("dt,tonemap,kw", [("f32", "reinhard", dict()), ("f32", "linear", dict()), ("f16", "linear", dict()),
                                           ("f32", "reinhard", dict(resize_width=48)), ("f16", "reinhard", dict(resize_width=48))])
def test_yuv420_output_on_every_fused_path(cuda, dt, tonemap, kw):
    """yuv420=True where no YUV epilogue exists (Camera32, linear, resizing ISPs): RGB8 sweep into a device scratch + the
    stand-alone conversion kernel -- still bit-identical to rgb_yuv420_image of the RGB result, metrics unchanged"""
    from taichi_image_b200 import bayer, camera_isp, color
    from tests.util import packed_frame
    cls = camera_isp.Camera16 if dt == "f16" else camera_isp.Camera32
    a, b = cls(bayer.BayerPattern.RGGB, **kw), cls(bayer.BayerPattern.RGGB, **kw)
    r = rng(84)
    for step in range(2):
        cu = [to_cuda(packed_frame(r, 64, 96)) for _ in range(2)]
        yuv = a.process_packed12(cu, tonemap=tonemap, gamma=0.9, yuv420=True)
        rgb = b.process_packed12(cu, tonemap=tonemap, gamma=0.9)
        for y, g in zip(yuv, rgb):
            assert tuple(y.shape) == (g.shape[0] * 3 // 2, g.shape[1]) and y.dtype == torch.uint8
            assert np.array_equal(to_np(y), to_np(color.rgb_yuv420_image(g))), f"step {step}"
        assert torch.equal(a.metrics, b.metrics)


def test_fused_yuv420_unsupported_configurations_fail(cuda):
    from taichi_image_b200 import bayer, camera_isp
    from tests.util import packed_frame
    fr = [to_cuda(packed_frame(rng(83), 16, 32))]
    with pytest.raises(AssertionError):
        camera_isp.Camera32(bayer.BayerPattern.RGGB).process_packed12(fr, tonemap="reinhard", yuv420=True, dtype="u16")
    # the C layer refuses as well (no silent RGB output into a YUV-sized buffer)
    from taichi_image_b200 import _lib
    import torch
    isp = camera_isp.Camera32(bayer.BayerPattern.RGGB)
    p = isp._fused_params(fr, "reinhard", isp.dtype, {}, update_metering=True, yuv420=True)
    p.out_dtype = 0
    out = torch.empty((24, 32), dtype=torch.uint8, device="cuda")
    m = torch.zeros(9, dtype=torch.float32, device="cuda")
    rc = _lib.lib.b200isp_process_packed12(_lib.ptr_array(fr), _lib.ptr_array([out]), 1, p, m.data_ptr(),
                                           _lib.workspace(isp.device).data_ptr(), _lib.stream_ptr(isp.device))
    assert rc != 0 and b"YUV" in _lib.lib.b200isp_last_error()

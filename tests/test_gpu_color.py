"""GPU parity: color/yuv_420.py (planar YUV 4:2:0) against the golden vectors of the reference source and the oracle."""
import os

import numpy as np
import pytest

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

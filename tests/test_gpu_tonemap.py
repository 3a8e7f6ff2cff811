"""GPU parity: tonemap.py (stand-alone) and interpolate.py against the oracle."""
import numpy as np
import pytest

from oracle import isp_oracle as O
from tests.util import rng, random_plane, smooth_rgb, to_cuda, to_np, assert_close_int, assert_close_float

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("in_name", ["u8", "u16", "f16", "f32"])
@pytest.mark.parametrize("out_name", ["u8", "u16", "f16", "f32"])
@pytest.mark.parametrize("gamma", [1.0, 0.6])
def test_tonemap_linear(cuda, in_name, out_name, gamma):
    from taichi_image_b200 import tonemap
    img = random_plane(rng(20), (33, 47, 3), in_name)
    got = to_np(tonemap.tonemap_linear(to_cuda(img), gamma=gamma, dtype=out_name))
    ref = O.tonemap_linear(img, gamma, out_name)
    if out_name in ("u8", "u16"):
        assert_close_int(got, ref, 1 if gamma != 1.0 or out_name == "u16" else 0, f"{in_name}->{out_name}")
    else:
        assert_close_float(got, ref, rtol=1e-3, atol=1e-3 if out_name == "f16" else 1e-6)


@pytest.mark.parametrize("in_name", ["u8", "f16", "f32"])
@pytest.mark.parametrize("out_name", ["u8", "u16", "f32"])
@pytest.mark.parametrize("params", [dict(), dict(gamma=0.6, intensity=3.0, light_adapt=0.9, color_adapt=0.2)])
def test_tonemap_reinhard(cuda, in_name, out_name, params):
    from taichi_image_b200 import tonemap
    r = rng(21)
    img = smooth_rgb(r, 48, 64, noise=0.1)
    if in_name == "u8":
        img = (img * 255).astype(np.uint8)
    else:
        img = img.astype(O.NP_DTYPE[in_name])
    got = to_np(tonemap.tonemap_reinhard(to_cuda(img), dtype=out_name, **params))
    ref = O.tonemap_reinhard(img, dtype=out_name, **params)
    if out_name == "u8":
        assert_close_int(got, ref, 1, f"{in_name}->{out_name}")
    elif out_name == "u16":
        assert_close_int(got, ref, 4, f"{in_name}->{out_name}")      # 5 f32 passes with pow: <= 6e-5 relative
    else:
        assert_close_float(got, ref, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("name", ["u8", "f16", "f32"])
@pytest.mark.parametrize("scale", [0.8, 0.469, 1.5])
def test_scale_bilinear(cuda, name, scale):
    from taichi_image_b200 import interpolate
    img = random_plane(rng(22), (40, 56, 3), name)
    got = to_np(interpolate.scale_bilinear(to_cuda(img), scale))
    h, w = img.shape[:2]
    ref = O.resize_bilinear(img, (int(w * scale), int(h * scale)), scale)
    assert got.shape == ref.shape
    if name == "u8":
        assert_close_int(got, ref, 0, "bilinear u8")
    else:
        assert_close_float(got, ref, rtol=1e-3, atol=1e-3 if name == "f16" else 1e-6)


def test_resize_width_and_per_axis(cuda):
    from taichi_image_b200 import interpolate
    img = random_plane(rng(23), (30, 50, 3), "f32")
    got = to_np(interpolate.resize_width(to_cuda(img), 20))
    ref = O.resize_bilinear(img, (20, int(30 * 20 / 50)), 20 / 50)
    assert_close_float(got, ref, rtol=1e-5, atol=1e-6)
    got = to_np(interpolate.resize_bilinear(to_cuda(img), (25, 10)))           # per-axis scale (SURVEY Q9 fix)
    ref = O.resize_bilinear(img, (25, 10), (10 / 30, 25 / 50))
    assert_close_float(got, ref, rtol=1e-5, atol=1e-6)
    got = to_np(interpolate.resize_bilinear(to_cuda(img), (25, 15), 0.5, dtype="u8"))   # in != out dtype (Q8)
    ref = O.resize_bilinear(img, (25, 15), 0.5, "u8")
    assert_close_int(got, ref, 0, "f32->u8")


@pytest.mark.parametrize("name", ["u8", "u16", "f16", "f32"])
@pytest.mark.parametrize("tname", O.TRANSFORMS)
@pytest.mark.parametrize("shape", [(5, 9), (64, 33), (70, 130),                   # element-wise kernel (unaligned row pitch for u8)
                                   (4, 4), (64, 128), (72, 132), (100, 260), (8, 68),    # word kernel: full and partial tiles
                                   (8, 8), (128, 32), (136, 40), (264, 1032), (256, 72)])  # turned path (both extents % 8 == 0): 128 x 32 tiles
def test_transform(cuda, name, tname, shape):
    import torch
    from taichi_image_b200 import interpolate
    img = random_plane(rng(24), shape + (3,), name)
    got = to_np(interpolate.transform(to_cuda(img), interpolate.ImageTransform[tname]))
    ref = O.transform(img, tname)
    assert got.shape == ref.shape and np.array_equal(np.ascontiguousarray(got).view(np.uint8), np.ascontiguousarray(ref).view(np.uint8)), tname
    if tname == "rotate_90":      # clockwise == torch.rot90(k=3)  
        assert np.array_equal(got, torch.rot90(torch.from_numpy(img), 3, (0, 1)).numpy())


def test_resize_area_matches_box_filter(cuda):
    from taichi_image_b200 import interpolate
    img = random_plane(rng(25), (48, 64, 3), "f32")
    got = to_np(interpolate.resize_area(to_cuda(img), (16, 12)))          # integer factor 4: plain block mean
    ref = img.reshape(12, 4, 16, 4, 3).mean(axis=(1, 3))
    assert_close_float(got, ref, rtol=1e-5, atol=1e-6)
    got = to_np(interpolate.resize_area(to_cuda(np.full((30, 50, 3), 0.25, np.float32)), (19, 7)))
    assert_close_float(got, np.full((7, 19, 3), 0.25, np.float32), rtol=1e-5, atol=1e-6)


def _area_ref(img, wd, hd):
    """exact box filter in float64: output (i, j) = coverage-weighted mean of the source over
    [i hs/hd, (i+1) hs/hd) x [j ws/wd, (j+1) ws/wd)  (separable coverage weights)"""
    hs, ws = img.shape[:2]

    def weights(n_src, n_dst):
        m = np.zeros((n_dst, n_src))
        f = n_src / n_dst
        for o in range(n_dst):
            a, b = o * f, min((o + 1) * f, n_src)
            for s in range(int(np.floor(a)), min(int(np.ceil(b)), n_src)):
                m[o, s] = max(0.0, min(b, s + 1) - max(a, s))
        return m / m.sum(axis=1, keepdims=True)

    wr, wc = weights(hs, hd), weights(ws, wd)
    return np.einsum("ir,jc,rck->ijk", wr, wc, img.astype(np.float64))


@pytest.mark.parametrize("size", [(19, 7), (37, 23), (50, 30), (64, 48), (70, 41)])
def test_resize_area_fractional_coverage(cuda, size):
    """EXTENSION: the fractional-coverage branch of area_kernel (csrc/resize.cu) on a RANDOM image: non-integer shrink
    factors in both axes, a factor of exactly 1 and an enlargement, against an exact float64 box filter"""
    from taichi_image_b200 import interpolate
    img = random_plane(rng(26), (30, 50, 3), "f32")
    wd, hd = size
    got = to_np(interpolate.resize_area(to_cuda(img), (wd, hd)))
    assert got.shape == (hd, wd, 3)
    assert_close_float(got, _area_ref(img, wd, hd).astype(np.float32), rtol=2e-5, atol=2e-6)
    u8 = random_plane(rng(27), (30, 50, 3), "u8")
    got8 = to_np(interpolate.resize_area(to_cuda(u8), (wd, hd)))
    assert np.abs(got8.astype(np.int64) - np.trunc(_area_ref(u8, wd, hd) + 1e-9).astype(np.int64)).max() <= 1

"""CPU: the numpy oracle against the golden vectors produced by executing the unmodified reference
sources on oracle/taichi_shim (oracle/gen_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from oracle import isp_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
DT = ("u8", "u16", "i16", "f16", "f32")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_packed_golden():
    g = load("packed")
    x = g["values12"]
    for ids in (0, 1):
        assert same(O.encode12(x, ids_format=bool(ids)), g[f"encode_ids{ids}"])
        assert same(O.decode12(g[f"encode_ids{ids}"], ids_format=bool(ids)), g[f"decode_ids{ids}"])
    assert np.array_equal(g["decode_ids0"], x)            # standard layout round-trips (reference test/packed.py)
    enc = g["encoded_random"]
    for name in DT:
        for ids in (0, 1):
            for scaled in (0, 1):
                assert same(O.decode12(enc, name, bool(scaled), bool(ids)), g[f"decode12_{name}_ids{ids}_scaled{scaled}"]), (name, ids, scaled)
            assert same(O.encode12(g[f"values_{name}"], scaled=True, ids_format=bool(ids)), g[f"encode12_scaled_{name}_ids{ids}"]), (name, ids)
    for name in ("u16", "f16", "f32"):
        for scaled in (0, 1):
            assert same(O.decode16(g["encoded16"], name, bool(scaled)), g[f"decode16_{name}_scaled{scaled}"])


def test_bayer_golden():
    g = load("bayer")
    ccm = g["ccm"].flatten().tolist()
    for name in ("u8", "u16", "f16", "f32"):
        for p in O.PATTERNS:
            assert same(O.rgb_to_bayer(g[f"rgb_{name}"], p), g[f"mosaic_{name}_{p}"])
            assert same(O.bayer_to_rgb(g[f"cfa_{name}"], p), g[f"demosaic_{name}_{p}"]), (name, p)
            assert same(O.bayer_to_rgb(g[f"cfa_{name}"], p, ccm), g[f"demosaic_ccm_{name}_{p}"]), (name, p)
    assert same(O.bayer_to_rgb(g["cfa_2x2"]), g["demosaic_2x2"])
    assert same(O.bayer_to_rgb(g["cfa_u8"], "GBRG", dtype="f32"), g["demosaic_u8_to_f32"])
    assert same(O.bayer_to_rgb(g["cfa_u16"], "GRBG", dtype="u8"), g["demosaic_u16_to_u8"])


def test_tonemap_golden():
    g = load("tonemap")
    for src in ("f32", "u8", "f16"):
        img = g[f"img_{src}"]
        for out in ("u8", "u16", "f16", "f32"):
            for gi, gamma in enumerate((1.0, 0.6)):
                assert same(O.tonemap_linear(img, gamma, out), g[f"linear_{src}_{out}_g{gi}"]), (src, out, gamma)
        for out in ("u8", "u16", "f32"):
            # the reference accumulates the metering sums sequentially in f32, the oracle in f64:
            # integer outputs may move by 1 LSB, floats by ~1e-6
            for key, kw in (("default", {}), ("params", dict(gamma=0.6, intensity=3.0, light_adapt=0.9, color_adapt=0.2))):
                got, ref = O.tonemap_reinhard(img, dtype=out, **kw), g[f"reinhard_{src}_{out}_{key}"]
                if out == "f32":
                    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
                else:
                    assert np.abs(got.astype(np.int64) - ref.astype(np.int64)).max() <= 1, (src, out, key)


def test_interpolate_golden():
    g = load("interpolate")
    for name in ("u8", "f16", "f32"):
        img = g[f"img_{name}"]
        h, w = img.shape[:2]
        for si, s in enumerate((0.8, 0.469, 1.5)):
            assert same(O.resize_bilinear(img, (int(w * s), int(h * s)), s), g[f"scale_{name}_{si}"]), (name, s)
        assert same(O.resize_bilinear(img, (9, int(h * (9 / w))), 9 / w), g[f"width_{name}"])
        for t in O.TRANSFORMS:
            if t != "transverse":
                assert same(O.transform(img, t), g[f"transform_{name}_{t}"]), (name, t)
    assert same(O.transform(g["img_square"], "transverse"), g["transform_square_transverse"])


CONFIGS = {"plain": dict(), "ccm": dict(correct_colors=True), "resize": dict(resize_width=16),
           "stride3": dict(metering_stride=3, moving_alpha=0.3)}
TMS = {"default": dict(), "script": dict(gamma=0.9, intensity=3.0, light_adapt=0.9, color_adapt=0.0),
       "coloradapt": dict(gamma=0.6, intensity=2.0, light_adapt=0.7, color_adapt=0.3),
       "linear": dict(gamma=0.8), "linear1": dict(gamma=1.0)}


def defined_mask(image, metrics, tm):
    """Pixels whose Reinhard value is a non-negative finite number.  Elsewhere (input below the
    sub-sampled moving-average lower bound) the reference casts a negative / NaN float to u8, which is
    undefined behaviour: those pixels are excluded from output comparisons."""
    kw = dict(gamma=1.0, intensity=1.0, light_adapt=1.0, color_adapt=0.0)
    kw.update(tm)
    _, stored, _ = O.isp_reinhard(image, metrics, return_intermediate=True, **kw)
    p = stored.astype(np.float32)
    return (np.isfinite(p) & (p >= 0)).all(axis=-1)


@pytest.mark.parametrize("cam", ["f16", "f32"])
@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_camera_isp_golden(cam, cfg):
    g = load("camera_isp")
    checked = 0
    for tm_name, tm in TMS.items():
        isp = O.ISP(cam, **CONFIGS[cfg])
        for s in range(2):
            frames = [g[f"frame_s{s}_c{c}"] for c in range(2)]
            images = [isp.load_packed12(f) for f in frames]
            if tm_name == "default":
                for c, im in enumerate(images):
                    assert same(im, g[f"{cam}_{cfg}_rgb_s{s}_c{c}"]), (cam, cfg, s, c)
            linear = tm_name.startswith("linear")
            outs = isp.tonemap_linear(images, **tm) if linear else isp.tonemap_reinhard(images, **tm)
            key = f"{cam}_{cfg}_{tm_name}_s{s}"
            assert np.allclose(isp.metrics, g[key + "_metrics"], rtol=2e-6, atol=2e-6), key
            for c, o in enumerate(outs):
                ref = g[key + f"_c{c}"]
                assert o.shape == ref.shape
                mask = np.ones(o.shape[:2], bool) if linear else defined_mask(images[c], isp.metrics, tm)
                d = np.abs(o.astype(np.int64) - ref.astype(np.int64))[mask]
                checked += d.size
                assert d.max() <= 1 and np.count_nonzero(d) <= 0.01 * d.size, (key, c, d.max())
    assert checked > 5000


def test_color_yuv420_golden():
    """color/yuv_420.py.  The decoder has no lower clamp (its tm.clamp arguments are swapped, yuv_420.py:88), so a
    negative component is cast to an unsigned integer -- undefined behaviour in the reference: those pixels are
    excluded for the integer dtypes (the oracle and the product saturate them to 0)."""
    g = load("color")
    for name in ("u8", "u16", "f16", "f32"):
        assert same(O.rgb_yuv420(g[f"rgb_{name}"]), g[f"yuv_{name}"]), name
        got, ref = O.yuv420_rgb(g[f"yuv_{name}"]), g[f"rgb_back_{name}"]
        assert got.shape == ref.shape and got.dtype == ref.dtype
        defined = (O.yuv420_rgb_unit(g[f"yuv_{name}"]) >= 0) | (name in ("f16", "f32"))
        assert defined.mean() > 0.5
        assert np.array_equal(got[defined], ref[defined]), name
    assert same(O.rgb_yuv420(g["rgb_u8"], "f32"), g["yuv_u8_to_f32"])
    assert same(O.rgb_yuv420(g["rgb_f32"], "u8"), g["yuv_f32_to_u8"])


WIDE_TMS = {"script": dict(gamma=0.9, intensity=3.0, light_adapt=0.9, color_adapt=0.0), "linear1": dict(gamma=1.0)}


@pytest.mark.parametrize("cam", ["f16", "f32"])
def test_camera_isp_wide_golden(cam):
    """the 20 x 776 golden case (oracle/gen_golden.py gen_camera_isp_wide): wide enough for interior strips of the
    CUDA sweep -- here it pins the oracle at that size"""
    g = load("camera_isp_wide")
    for tm_name, tm in WIDE_TMS.items():
        isp = O.ISP(cam)
        for s in range(2):
            frames = [g[f"frame_s{s}_c{c}"] for c in range(2)]
            images = [isp.load_packed12(f) for f in frames]
            for c, im in enumerate(images):
                assert same(im, g[f"{cam}_rgb_s{s}_c{c}"]), (cam, s, c)
            linear = tm_name.startswith("linear")
            outs = isp.tonemap_linear(images, **tm) if linear else isp.tonemap_reinhard(images, **tm)
            key = f"{cam}_{tm_name}_s{s}"
            assert np.allclose(isp.metrics, g[key + "_metrics"], rtol=2e-6, atol=2e-6), key
            for c, o in enumerate(outs):
                ref = g[key + f"_c{c}"]
                mask = np.ones(o.shape[:2], bool) if linear else defined_mask(g[f"{cam}_rgb_s{s}_c{c}"], isp.metrics, tm)
                d = np.abs(o.astype(np.int64) - ref.astype(np.int64))[mask]
                assert d.max() <= 1 and np.count_nonzero(d) <= 0.01 * d.size, (key, c, d.max())

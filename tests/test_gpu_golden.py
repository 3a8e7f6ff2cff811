"""GPU parity against the golden vectors of the reference itself (tests/golden/, produced by running the
unmodified reference sources on oracle/taichi_shim).  The oracle is not involved here except to mask the
pixels where the reference's own behaviour is undefined (negative / NaN float -> u8 cast)."""
import os

import numpy as np
import pytest
import torch

from tests.test_oracle_golden import load, CONFIGS, TMS, defined_mask
from tests.util import to_cuda, to_np, assert_close_float

pytestmark = pytest.mark.gpu
DT = ("u8", "u16", "i16", "f16", "f32")
PATTERNS = ("RGGB", "GRBG", "GBRG", "BGGR")


def same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_packed_golden(cuda):
    from taichi_image_b200 import packed
    g = load("packed")
    x = g["values12"]
    for ids in (0, 1):
        e = packed.encode12(to_cuda(x), ids_format=bool(ids))
        assert same(to_np(e), g[f"encode_ids{ids}"])
        assert same(to_np(packed.decode12(e, ids_format=bool(ids))), g[f"decode_ids{ids}"])
    enc = to_cuda(g["encoded_random"])
    for name in DT:
        for ids in (0, 1):
            for scaled in (0, 1):
                got = to_np(packed.decode12(enc, dtype=name, scaled=bool(scaled), ids_format=bool(ids)))
                assert same(got, g[f"decode12_{name}_ids{ids}_scaled{scaled}"]), (name, ids, scaled)
            got = to_np(packed.encode12(to_cuda(g[f"values_{name}"]), scaled=True, ids_format=bool(ids)))
            assert same(got, g[f"encode12_scaled_{name}_ids{ids}"]), (name, ids)
    for name in ("u16", "f16", "f32"):
        for scaled in (0, 1):
            assert same(to_np(packed.decode16(to_cuda(g["encoded16"]), dtype=name, scaled=bool(scaled))), g[f"decode16_{name}_scaled{scaled}"])


def test_bayer_golden(cuda):
    from taichi_image_b200 import bayer
    g = load("bayer")
    for name in ("u8", "u16", "f16", "f32"):
        for p in PATTERNS:
            pat = bayer.BayerPattern[p]
            assert same(to_np(bayer.rgb_to_bayer(to_cuda(g[f"rgb_{name}"]), pat)), g[f"mosaic_{name}_{p}"])
            for key, ccm in ((f"demosaic_{name}_{p}", None), (f"demosaic_ccm_{name}_{p}", g["ccm"])):
                got = to_np(bayer.bayer_to_rgb(to_cuda(g[f"cfa_{name}"]), pat, correct_colors=ccm))
                if name in ("u8", "u16"):
                    assert same(got, g[key]), key          # 10x12 planes: width % 8 != 0 -> literal per-pixel kernel
                else:
                    assert_close_float(got, g[key], rtol=1e-3, atol=1e-3 if name == "f16" else 1e-6, what=key)
    assert same(to_np(bayer.bayer_to_rgb(to_cuda(g["cfa_2x2"]))), g["demosaic_2x2"])
    assert_close_float(to_np(bayer.bayer_to_rgb(to_cuda(g["cfa_u8"]), bayer.BayerPattern.GBRG, dtype="f32")), g["demosaic_u8_to_f32"], 1e-6, 1e-7)
    assert same(to_np(bayer.bayer_to_rgb(to_cuda(g["cfa_u16"]), bayer.BayerPattern.GRBG, dtype="u8")), g["demosaic_u16_to_u8"])


def test_tonemap_golden(cuda):
    from taichi_image_b200 import tonemap
    g = load("tonemap")
    for src in ("f32", "u8", "f16"):
        img = to_cuda(g[f"img_{src}"])
        for out in ("u8", "u16", "f16", "f32"):
            for gi, gamma in enumerate((1.0, 0.6)):
                got, ref = to_np(tonemap.tonemap_linear(img, gamma, out)), g[f"linear_{src}_{out}_g{gi}"]
                if out in ("u8", "u16"):
                    assert np.abs(got.astype(np.int64) - ref.astype(np.int64)).max() <= (0 if gamma == 1.0 else 1), (src, out, gamma)
                else:
                    assert_close_float(got, ref, rtol=1e-3, atol=1e-3 if out == "f16" else 1e-6)
        for out in ("u8", "u16", "f32"):
            for key, kw in (("default", {}), ("params", dict(gamma=0.6, intensity=3.0, light_adapt=0.9, color_adapt=0.2))):
                got, ref = to_np(tonemap.tonemap_reinhard(img, dtype=out, **kw)), g[f"reinhard_{src}_{out}_{key}"]
                if out == "f32":
                    assert_close_float(got, ref, rtol=1e-3, atol=1e-5)
                else:
                    assert np.abs(got.astype(np.int64) - ref.astype(np.int64)).max() <= 1, (src, out, key)     # u8 and u16 (measured: r02_error_histogram)


def test_interpolate_golden(cuda):
    from taichi_image_b200 import interpolate
    g = load("interpolate")
    for name in ("u8", "f16", "f32"):
        img = to_cuda(g[f"img_{name}"])
        for si, s in enumerate((0.8, 0.469, 1.5)):
            got, ref = to_np(interpolate.scale_bilinear(img, s)), g[f"scale_{name}_{si}"]
            assert same(got, ref), (name, s)
        assert same(to_np(interpolate.resize_width(img, 9)), g[f"width_{name}"])
        for t in interpolate.ImageTransform:
            if t != interpolate.ImageTransform.transverse:
                assert same(to_np(interpolate.transform(img, t)), g[f"transform_{name}_{t.value}"]), (name, t)
    assert same(to_np(interpolate.transform(to_cuda(g["img_square"]), interpolate.ImageTransform.transverse)),
                g["transform_square_transverse"])


@pytest.mark.parametrize("cam", ["f16", "f32"])
@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("fused", [False, True])
def test_camera_isp_golden(cuda, cam, cfg, fused):
    """eager (load_packed12 + tonemap_*) and fused (process_packed12) paths against the reference outputs"""
    from taichi_image_b200 import camera_isp, bayer
    g = load("camera_isp")
    cls = camera_isp.Camera16 if cam == "f16" else camera_isp.Camera32
    for tm_name, tm in TMS.items():
        isp = cls(bayer.BayerPattern.RGGB, **CONFIGS[cfg])
        for s in range(2):
            frames = [to_cuda(g[f"frame_s{s}_c{c}"]) for c in range(2)]
            linear = tm_name.startswith("linear")
            if fused:
                outs = isp.process_packed12(frames, tonemap="linear" if linear else "reinhard", **tm)
            else:
                images = [isp.load_packed12(f) for f in frames]
                if tm_name == "default":
                    for c, im in enumerate(images):
                        assert_close_float(to_np(im), g[f"{cam}_{cfg}_rgb_s{s}_c{c}"], rtol=1e-3,
                                           atol=1e-3 if cam == "f16" else 2e-6)
                outs = isp.tonemap_linear(images, **tm) if linear else isp.tonemap_reinhard(images, **tm)
            key = f"{cam}_{cfg}_{tm_name}_s{s}"
            assert_close_float(to_np(isp.metrics), g[key + "_metrics"], rtol=2e-5, atol=2e-6, what=key)
            for c, o in enumerate(outs):
                ref = g[key + f"_c{c}"]
                assert tuple(o.shape) == ref.shape and o.dtype == torch.uint8
                mask = np.ones(ref.shape[:2], bool) if linear else \
                    defined_mask(g[f"{cam}_{cfg}_rgb_s{s}_c{c}"], g[key + "_metrics"], tm)
                d = np.abs(to_np(o).astype(np.int64) - ref.astype(np.int64))[mask]
                assert d.max() <= 1, (key, c, int(d.max()), int(np.count_nonzero(d)))


@pytest.mark.parametrize("cam", ["f16", "f32"])
@pytest.mark.parametrize("rpt", [0, 6])
def test_camera_isp_wide_golden(cuda, cam, rpt):
    """The fused sweep against the reference SOURCE at a width that reaches its interior-strip kind (K_CORE), the
    partial last strip and (rows_per_task = 6) several row chunks per strip: 20 x 776, two time steps x two cameras."""
    from taichi_image_b200 import camera_isp, bayer
    from tests.test_oracle_golden import WIDE_TMS
    g = load("camera_isp_wide")
    cls = camera_isp.Camera16 if cam == "f16" else camera_isp.Camera32
    for tm_name, tm in WIDE_TMS.items():
        isp = cls(bayer.BayerPattern.RGGB)
        for s in range(2):
            frames = [to_cuda(g[f"frame_s{s}_c{c}"]) for c in range(2)]
            if tm_name == "script":
                for c, f in enumerate(frames):
                    assert_close_float(to_np(isp.load_packed12(f)), g[f"{cam}_rgb_s{s}_c{c}"], rtol=1e-3,
                                       atol=1e-3 if cam == "f16" else 2e-6)
            linear = tm_name.startswith("linear")
            outs = isp.process_packed12(frames, tonemap="linear" if linear else "reinhard", rows_per_task=rpt, **tm)
            key = f"{cam}_{tm_name}_s{s}"
            assert_close_float(to_np(isp.metrics), g[key + "_metrics"], rtol=2e-5, atol=2e-6, what=key)
            for c, o in enumerate(outs):
                ref = g[key + f"_c{c}"]
                mask = np.ones(ref.shape[:2], bool) if linear else defined_mask(g[f"{cam}_rgb_s{s}_c{c}"], g[key + "_metrics"], tm)
                d = np.abs(to_np(o).astype(np.int64) - ref.astype(np.int64))[mask]
                assert d.max() <= 1, (key, c, int(d.max()), int(np.count_nonzero(d)))

"""GPU parity of the fused demosaic + bilinear-resize gather (csrc/fused_resize.cu, resize_isp.cuh) behind a resizing
ISP (camera_isp.py:302-315, :371-373; interpolate.py:19-34, :59-66): against the numpy oracle at small sizes (all
patterns, both ISP dtypes, CCM, down- and up-scaling, per-axis target size), against the staged kernels (bit-exact)
and against the C oracle at BASELINE configs[4]'s size (4096x3000 -> width 1920)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import isp_oracle as O
from tests.test_gpu_camera_isp import frames, make_isp
from tests.util import rng, to_cuda, to_np, assert_close_int, assert_close_float

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("pattern", O.PATTERNS)
@pytest.mark.parametrize("kw", [dict(resize_width=40), dict(scale=0.469), dict(scale=1.3), dict(resize_width=79)])
def test_load_packed12_resized(cuda, dt, pattern, kw):
    """ISP.load_packed12 of a resizing ISP = demosaic + resize in one gather: float RGB of the ISP dtype"""
    r = rng(80)
    fr = frames(r, 1, 48, 80, pattern)[0]
    isp, ref = make_isp(dt, bayer_pattern=pattern, **kw), O.ISP(dt, pattern, **kw)
    got, exp = to_np(isp.load_packed12(to_cuda(fr))), ref.load_packed12(fr)
    assert got.shape == exp.shape and got.dtype == exp.dtype
    assert_close_float(got, exp, rtol=1e-3, atol=1e-3 if dt == "f16" else 2e-6, what=f"{dt} {pattern} {kw}")


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("tonemap,out", [("reinhard", "u8"), ("reinhard", "f16"), ("linear", "u16"), ("linear", "u8")])
@pytest.mark.parametrize("ccm", [False, True])
def test_fused_resize_tonemap(cuda, dt, tonemap, out, ccm):
    """process_packed12 on a resizing ISP: metering and tone map on the RESIZED images, two moving-average steps"""
    r = rng(81)
    isp, ref = make_isp(dt, resize_width=56, correct_colors=ccm), O.ISP(dt, resize_width=56, correct_colors=ccm)
    tm = dict(gamma=0.9, intensity=3.0, light_adapt=0.9) if tonemap == "reinhard" else dict(gamma=0.8)
    for step in range(2):
        fr = frames(r, 3, 64, 128)
        got = isp.process_packed12([to_cuda(f) for f in fr], tonemap=tonemap, dtype=out, **tm)
        ims = [ref.load_packed12(f) for f in fr]
        exp = ref.tonemap_reinhard(ims, out_dtype=out, **tm) if tonemap == "reinhard" else ref.tonemap_linear(ims, out_dtype=out, **tm)
        assert_close_float(to_np(isp.metrics), ref.metrics, rtol=1e-4, atol=1e-5, what="metrics")
        for g, e in zip(got, exp):
            assert tuple(g.shape) == e.shape == (28, 56, 3)
            if out == "f16":
                assert_close_float(to_np(g), e, rtol=2e-3, atol=1e-3)
            elif dt == "f16" and out == "u16":
                from tests.util import assert_u16_from_f16_isp
                assert_u16_from_f16_isp(to_np(g), e, (1.0 / 0.8) / float(ref.metrics[1] - ref.metrics[0]), f"{dt} {tonemap}")
            else:
                assert_close_int(to_np(g), e, 1, f"{dt} {tonemap}->{out} step {step}")


@pytest.mark.parametrize("gather", [False, True])
def test_fused_resize_equals_staged_kernels_bit_exact(cuda, gather, monkeypatch):
    """Camera32: the resizing sweep (and the per-output-pixel gather, B200ISP_RESIZE_GATHER=1) evaluate the same exact
    integer sums as the full-resolution sweep and mix with the same per-op rounding as the stand-alone bilinear kernel
    -> bit-identical to load_packed12 (full size) + interpolate.resize_bilinear wherever no tap lies on the 2-pixel
    image frame (there the gather's literal 13-tap path rounds differently in the last bit)"""
    from taichi_image_b200 import interpolate
    monkeypatch.setenv("B200ISP_RESIZE_GATHER", "1" if gather else "0")
    r = rng(82)
    for pattern in O.PATTERNS:
        fr = to_cuda(frames(r, 1, 96, 776, pattern)[0])
        for kw, size, scale in ((dict(resize_width=360), (360, round(96 * 360 / 776)), 360 / 776), (dict(scale=0.77), (round(776 * 0.77), round(96 * 0.77)), 0.77)):
            full = make_isp("f32", bayer_pattern=pattern).load_packed12(fr)
            staged = interpolate.resize_bilinear(full, size, scale)
            fused = make_isp("f32", bayer_pattern=pattern, **kw).load_packed12(fr)
            assert staged.shape == fused.shape
            assert float((staged - fused).abs().max()) <= 2e-6, (pattern, kw)
            assert torch.equal(staged[4:-4, 4:-4], fused[4:-4, 4:-4]), (pattern, kw)


def test_fused_resize_sweep_equals_gather(cuda, monkeypatch):
    """the two fused forms agree everywhere (Camera16 too: same samples, same formulas up to the frame pixels)"""
    r = rng(84)
    fr = [to_cuda(f) for f in frames(r, 2, 120, 1288)]
    for dt in ("f16", "f32"):
        outs = {}
        for gather in ("0", "1"):
            monkeypatch.setenv("B200ISP_RESIZE_GATHER", gather)
            isp = make_isp(dt, resize_width=600)
            outs[gather] = isp.process_packed12(fr, tonemap="reinhard", gamma=0.9, intensity=3.0, light_adapt=0.9, dtype="u8")
        for a, b in zip(outs["0"], outs["1"]):
            assert int((a.int() - b.int()).abs().max()) <= 1


def test_resize_size_per_axis_extension(cuda):
    """EXTENSION: resize_size=(w, h) with per-axis scales (BASELINE configs[4]: 1920x1080 from a 4:3 sensor)"""
    from taichi_image_b200 import interpolate
    r = rng(83)
    fr = to_cuda(frames(r, 1, 60, 80)[0])
    full = make_isp("f32").load_packed12(fr)
    fused = make_isp("f32", resize_size=(48, 27)).load_packed12(fr)
    assert tuple(fused.shape) == (27, 48, 3)
    assert float((fused - interpolate.resize_bilinear(full, (48, 27))).abs().max()) <= 2e-6


@pytest.mark.parametrize("dt,tonemap,out", [("f16", "reinhard", "u8"), ("f32", "reinhard", "u8"), ("f32", "linear", "u16")])
def test_cfg5_size_vs_c_oracle(cuda, dt, tonemap, out):
    """BASELINE configs[4] geometry: 4096x3000 -> resize_width 1920 (1920 x 1406), two frames, two time steps, against
    the C oracle (demosaic -> resize -> joint metering -> tone map)"""
    from tests.test_gpu_fullsize import synth_packed, compare_with_c_oracle
    tm = dict(gamma=0.9, intensity=3.0, light_adapt=0.9, color_adapt=0.0) if tonemap == "reinhard" else dict(gamma=1.0)
    isp = make_isp(dt, resize_width=1920, moving_alpha=0.1)
    scale = 1920 / 4096
    metrics = None
    for step in range(2):
        dev, host = synth_packed(2, 3000, 4096, seed=41 + step)
        got = isp.process_packed12(dev, tonemap=tonemap, dtype=out, **tm)
        exp, metrics = c_oracle.process(host, "RGGB", dt == "f16", out, tonemap, None, tm.get("gamma", 1.0), tm.get("intensity", 1.0),
                                        tm.get("light_adapt", 1.0), tm.get("color_adapt", 0.0), alpha=0.0 if step == 0 else 0.9,
                                        metrics=metrics, resize=((1920, round(3000 * scale)), (scale, scale)))
        assert tuple(got[0].shape) == (1406, 1920, 3)
        np.testing.assert_allclose(isp.metrics.cpu().numpy(), metrics, rtol=2e-5, atol=2e-6)
        compare_with_c_oracle(f"cfg5 {dt} {tonemap} step {step}", got, exp)

"""GPU parity: the fused packed12 sweep with the BILINEAR demosaic (north_star extension, ``ISP(demosaic="bilinear")``)
against the oracle ISP built on ``bayer_to_rgb_bilinear`` -- every CFA pattern, Camera16 / Camera32, load / linear /
Reinhard, all task kinds (narrow, wide, ragged strips), metering from the packed frames, and the staged fallback."""
import numpy as np
import pytest
import torch

from oracle import isp_oracle as O
from tests.util import rng, packed_frame, to_cuda, to_np, assert_close_int, assert_close_float, assert_u16_from_f16_isp
from tests.test_gpu_camera_isp import make_isp, frames, TM

pytestmark = pytest.mark.gpu


def pair(dt, pattern="RGGB", **kw):
    return (make_isp(dt, bayer_pattern=pattern, demosaic="bilinear", **kw),
            O.ISP(dt, pattern, demosaic="bilinear", **{k: v for k, v in kw.items()}))


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("pattern", O.PATTERNS)
@pytest.mark.parametrize("shape", [(32, 48), (30, 36), (4, 8), (12, 264), (20, 520)])   # fused / generic / minimal / 2 / 3 strips
@pytest.mark.parametrize("ccm", [False, True])
def test_load_packed12_bilinear(cuda, dt, pattern, shape, ccm):
    pk = packed_frame(rng(60), *shape, pattern)
    isp, ref = pair(dt, pattern, correct_colors=ccm)
    got = to_np(isp.load_packed12(to_cuda(pk)))
    exp = ref.load_packed12(pk)
    assert got.dtype == exp.dtype and got.shape == exp.shape
    assert_close_float(got, exp, rtol=1e-3, atol=1e-3 if dt == "f16" else 2e-6, what=f"{dt} {pattern} {shape}")


def test_bilinear_differs_from_malvar(cuda):
    """guards against the flag being ignored: the two demosaics disagree on a textured frame"""
    pk = to_cuda(packed_frame(rng(61), 32, 64, smooth=False))
    a = to_np(make_isp("f32", demosaic="bilinear").load_packed12(pk))
    b = to_np(make_isp("f32").load_packed12(pk))
    assert np.abs(a - b).max() > 0.05


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("pattern", O.PATTERNS)
@pytest.mark.parametrize("tm", TM[1:])
def test_fused_reinhard_bilinear(cuda, dt, pattern, tm):
    r = rng(62)
    isp, ref = pair(dt, pattern)
    for step in range(3):
        fr = frames(r, 3, 40, 64, pattern)
        got = isp.process_packed12([to_cuda(f) for f in fr], tonemap="reinhard", **tm)
        exp = ref.tonemap_reinhard([ref.load_packed12(f) for f in fr], **tm)
        for g, e in zip(got, exp):
            assert_close_int(to_np(g), e, 1, f"{dt} {pattern} {tm} step {step}")
        assert_close_float(to_np(isp.metrics), ref.metrics, rtol=1e-4, atol=1e-5, what="metrics")


@pytest.mark.parametrize("dt,out", [("f32", "u8"), ("f32", "u16"), ("f16", "u8"), ("f16", "u16")])
@pytest.mark.parametrize("gamma", [1.0, 0.7])
@pytest.mark.parametrize("ccm", [False, True])
def test_fused_linear_bilinear(cuda, dt, out, gamma, ccm):
    r = rng(63)
    isp, ref = pair(dt, "GRBG", correct_colors=ccm)
    for step in range(2):
        fr = frames(r, 2, 36, 72, "GRBG")
        got = isp.process_packed12([to_cuda(f) for f in fr], tonemap="linear", gamma=gamma, dtype=out)
        exp = ref.tonemap_linear([ref.load_packed12(f) for f in fr], gamma=gamma, out_dtype=out)
        for g, e in zip(got, exp):
            if dt == "f16" and out == "u16":      # one f16 ulp of the intermediate RGB through the tone map's gain
                gain = max(1.0, 1.0 / gamma) / float(ref.metrics[1] - ref.metrics[0])
                assert_u16_from_f16_isp(to_np(g), e, gain, f"{dt}->{out} gamma {gamma}")
            else:
                assert_close_int(to_np(g), e, 1, f"{dt}->{out} gamma {gamma}")


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("shape", [(24, 776), (10, 1032)])       # core + edge + ragged last strip; border tasks dominate
def test_fused_bilinear_wide_frames(cuda, dt, shape):
    r = rng(64)
    isp, ref = pair(dt, "BGGR")
    fr = frames(r, 2, *shape, "BGGR")
    got = isp.process_packed12([to_cuda(f) for f in fr], tonemap="linear", dtype="u8")
    exp = ref.tonemap_linear([ref.load_packed12(f) for f in fr], out_dtype="u8")
    for g, e in zip(got, exp):
        assert_close_int(to_np(g), e, 1, f"{dt} {shape}")
    got = isp.process_packed12([to_cuda(f) for f in fr], tonemap="reinhard", gamma=0.8, dtype="u8")
    exp = ref.tonemap_reinhard([ref.load_packed12(f) for f in fr], gamma=0.8)
    for g, e in zip(got, exp):
        assert_close_int(to_np(g), e, 1, f"{dt} {shape} reinhard")


@pytest.mark.parametrize("stride", [8, 3])        # word-aligned fast sampler / generic per-pixel sampler
def test_bilinear_metering_matches_staged(cuda, stride):
    """metering straight from the packed frames == metering of the demosaiced images (same samples)"""
    r = rng(65)
    fr = frames(r, 3, 40, 56)
    a = make_isp("f32", demosaic="bilinear", metering_stride=stride)
    b = make_isp("f32", demosaic="bilinear", metering_stride=stride)
    ref = O.ISP("f32", demosaic="bilinear", metering_stride=stride)
    for step in range(2):
        a.process_packed12([to_cuda(f) for f in fr], tonemap="linear")
        b.update_metering([b.load_packed12(to_cuda(f)) for f in fr])
        ref.update_metering([ref.load_packed12(f) for f in fr])
        assert_close_float(to_np(a.metrics), to_np(b.metrics), rtol=2e-5, atol=2e-6, what=f"fused vs staged, step {step}")
        assert_close_float(to_np(a.metrics), ref.metrics, rtol=1e-4, atol=1e-5, what=f"vs oracle, step {step}")


def test_bilinear_with_resize_uses_staged_kernels(cuda):
    r = rng(66)
    isp, ref = pair("f16", resize_width=40)
    fr = frames(r, 2, 48, 80)
    got = isp.process_packed12([to_cuda(f) for f in fr], tonemap="reinhard", gamma=0.6)
    exp = ref.tonemap_reinhard([ref.load_packed12(f) for f in fr], gamma=0.6)
    for g, e in zip(got, exp):
        assert tuple(g.shape) == e.shape == (24, 40, 3)
        assert_close_int(to_np(g), e, 1, "bilinear + resize")


def test_unknown_demosaic_is_rejected(cuda):
    from taichi_image_b200 import _lib
    with pytest.raises(AssertionError):
        make_isp("f32", demosaic="nearest")
    isp = make_isp("f32")
    fr = [to_cuda(packed_frame(rng(67), 16, 32))]
    p = isp._fused_params(fr, "linear", isp.dtype, {})
    p.demosaic = 7
    out = torch.empty((16, 32, 3), dtype=torch.float32, device="cuda")
    rc = _lib.lib.b200isp_process_packed12(_lib.ptr_array(fr), _lib.ptr_array([out]), 1, p, 0,
                                           _lib.workspace(isp.device).data_ptr(), _lib.stream_ptr(isp.device))
    assert rc != 0 and b"demosaic" in _lib.lib.b200isp_last_error()

"""GPU: the BASELINE.json configurations at FULL size -- against the C/OpenMP oracle (oracle/c/isp_oracle.c: ~0.1 s per
frame on the box's host cores) on the same seeded inputs, and through size-independent properties: mosaic / demosaic
round trips, constant and linearity properties, cross-checks between the fused sweep and the staged kernels,
determinism."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import isp_oracle as O
from tests.test_gpu_camera_isp import make_isp

pytestmark = pytest.mark.gpu


def packed_from_cfa(cfa12: torch.Tensor) -> torch.Tensor:
    """standard packed12 layout of a (H, W) tensor of 12-bit codes, on the device"""
    from taichi_image_b200 import packed
    return packed.encode12(cfa12.to(torch.uint16))


@pytest.mark.parametrize("pattern", O.PATTERNS)
@pytest.mark.parametrize("dtype", [torch.uint8, torch.uint16])
def test_cfg4_round_trip_8k(cuda, pattern, dtype):
    """BASELINE configs[3]: rgb_to_bayer -> bayer_to_rgb -> rgb_to_bayer is the identity at 7680x4320 (bit-exact:
    the demosaic reproduces every CFA sample)"""
    from taichi_image_b200 import bayer
    g = torch.Generator(device="cuda").manual_seed(7)
    hi = 256 if dtype == torch.uint8 else 65536
    rgb = torch.randint(0, hi, (4320, 7680, 3), generator=g, device="cuda", dtype=torch.int32).to(dtype)
    p = bayer.BayerPattern[pattern]
    b = bayer.rgb_to_bayer(rgb, p)
    back = bayer.rgb_to_bayer(bayer.bayer_to_rgb(b, p), p)
    assert torch.equal(back, b)


@pytest.mark.parametrize("dt", ["f16", "f32"])
def test_cfg2_constant_and_determinism(cuda, dt):
    """BASELINE configs[1] size (5472x3648): a constant sensor image demosaics to the same constant everywhere
    (including the renormalised image frame), and two runs are bit-identical (deterministic reductions)"""
    h, w = 3648, 5472
    isp = make_isp(dt)
    cfa = torch.full((h, w), 2000, dtype=torch.int32, device="cuda")
    frame = packed_from_cfa(cfa)
    rgb = isp.load_packed12(frame)
    ref = np.float32(2000) * np.float32(1.0 / 4095.0)
    ref = np.float32(np.float16(ref)) if dt == "f16" else ref
    assert float((rgb.float() - float(ref)).abs().max()) <= (1e-3 if dt == "f16" else 2e-6)
    g = torch.Generator(device="cuda").manual_seed(11)
    noisy = packed_from_cfa(torch.randint(0, 4096, (h, w), generator=g, device="cuda", dtype=torch.int32))
    a = make_isp(dt).process_packed12([noisy, frame], tonemap="linear", dtype="u16")
    b = make_isp(dt).process_packed12([noisy, frame], tonemap="linear", dtype="u16")
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_cfg1_fused_equals_staged(cuda):
    """BASELINE configs[0] size (4096x3000, Reinhard -> RGB8): the fused sweep and the staged reference-shaped calls
    (load_packed12 -> tonemap_reinhard) agree within 1 LSB and produce the same metrics"""
    h, w = 3000, 4096
    g = torch.Generator(device="cuda").manual_seed(13)
    base = (torch.rand((h // 8, w // 8), generator=g, device="cuda") * 3000 + 200)
    cfa = torch.nn.functional.interpolate(base[None, None], size=(h, w), mode="bilinear")[0, 0].to(torch.int32)
    frame = packed_from_cfa(cfa)
    fused, staged = make_isp("f32"), make_isp("f32")
    y1 = fused.process_packed12([frame], tonemap="reinhard", gamma=0.9, intensity=3.0, light_adapt=0.9)[0]
    y2 = staged.tonemap_reinhard([staged.load_packed12(frame)], gamma=0.9, intensity=3.0, light_adapt=0.9)[0]
    d = (y1.int() - y2.int()).abs()
    assert int(d.max()) <= 1, f"max diff {int(d.max())}"
    np.testing.assert_allclose(fused.metrics.cpu().numpy(), staged.metrics.cpu().numpy(), rtol=1e-4, atol=1e-6)


def test_packed_round_trip_full_frame(cuda):
    """decode12(encode12(x)) == x on a full 4096x3000 frame (reference test/packed.py at full size).  The reference's
    IDS encoder and decoder are NOT inverses: encode puts the low nibble of the first pixel into the HIGH nibble of
    byte 2 (packed.py:47-55), decode reads it from the LOW nibble (packed.py:36-44), so the pair's low nibbles come
    back swapped -- reproduced bit for bit (the golden vectors pin each direction separately)."""
    from taichi_image_b200 import packed, u16
    g = torch.Generator(device="cuda").manual_seed(17)
    x = torch.randint(0, 4096, (3000, 4096), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
    assert torch.equal(packed.decode12(packed.encode12(x), dtype=u16), x)
    y = packed.decode12(packed.encode12(x, ids_format=True), dtype=u16, ids_format=True).int()
    xi = x.int()
    swapped = torch.empty_like(xi)
    swapped[:, 0::2] = (xi[:, 0::2] & ~0xF) | (xi[:, 1::2] & 0xF)
    swapped[:, 1::2] = (xi[:, 1::2] & ~0xF) | (xi[:, 0::2] & 0xF)
    assert torch.equal(y, swapped)


# ---------------------------------------------------------------- BASELINE configs at full size vs the C oracle
def synth_packed(n, h, w, seed, pattern="RGGB"):
    """SURVEY 8d generator on the device: smooth HDR-ish field x channel gains + 2 % noise, mosaiced, 12 bit, packed
    (packed.encode12 is itself bit-exact against the golden vectors).  Returns device frames and host copies."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    yy = torch.arange(h, device="cuda", dtype=torch.float32)[:, None] / h
    xx = torch.arange(w, device="cuda", dtype=torch.float32)[None, :] / w
    order = {"RGGB": (0, 1, 1, 2), "GRBG": (1, 0, 2, 1), "GBRG": (1, 2, 0, 1), "BGGR": (2, 1, 1, 0)}[pattern]
    gains = (0.9, 1.0, 0.7)
    dev, host = [], []
    for i in range(n):
        base = 0.5 + 0.35 * torch.sin(xx * (5.1 + i) + 0.3 * i + 0.01 * seed) * torch.cos(yy * 3.7)
        cfa = torch.empty((h, w), device="cuda")
        for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            cfa[dy::2, dx::2] = base[dy::2, dx::2] * gains[order[k]]
        cfa += 0.05 + 0.02 * (torch.rand((h, w), generator=g, device="cuda") - 0.5)
        f = packed_from_cfa(torch.round(cfa.clamp(0, 1) * 4095).to(torch.int32))
        dev.append(f)
        host.append(f.cpu().numpy())
    return dev, host


def compare_with_c_oracle(what, got, exp, max_lsb=1, max_frac=0.10):
    """max |diff| <= max_lsb and the fraction of differing values (trunc-cast flips at rounding boundaries)"""
    worst, nd, nt = 0, 0, 0
    for gt, e in zip(got, exp):
        d = (gt.to(torch.int32) - torch.from_numpy(e.astype(np.int32)).cuda()).abs()
        worst = max(worst, int(d.max()))
        nd += int((d != 0).sum())
        nt += d.numel()
    print(f"[fullsize] {what}: max |diff| = {worst} LSB, differing = {nd / nt:.3e} of {nt} values")
    assert worst <= max_lsb, f"{what}: max |diff| = {worst} LSB"
    assert nd <= max_frac * nt, f"{what}: {nd / nt:.3f} of the values differ"


FULL = {
    # name: (frames, H, W, ISP dtype, tonemap, out dtype, tone-map settings)
    "cfg1_default": (1, 3000, 4096, "f32", "reinhard", "u8", dict()),
    "cfg1_script": (1, 3000, 4096, "f32", "reinhard", "u8", dict(gamma=0.9, intensity=3.0, light_adapt=0.9, color_adapt=0.0)),
    "cfg1_cam16_script": (1, 3000, 4096, "f16", "reinhard", "u8", dict(gamma=0.9, intensity=3.0, light_adapt=0.9, color_adapt=0.0)),
    "cfg1_cam16_bench": (2, 3000, 4096, "f16", "reinhard", "u8", dict(gamma=0.6)),
    "cfg2": (6, 3648, 5472, "f32", "linear", "u16", dict(gamma=1.0)),
    "cfg3_shard": (6, 3000, 4096, "f32", "reinhard", "u8", dict(gamma=0.9, intensity=3.0, light_adapt=0.9, color_adapt=0.0)),
}


@pytest.mark.parametrize("name", list(FULL))
def test_baseline_config_vs_c_oracle(cuda, name):
    """Every BASELINE configuration at its full size, first frame (metrics = None) and after three moving-average steps
    on changing frames: outputs within 1 LSB of the C oracle, metrics within 2e-5 relative (the reductions run in a
    different, deterministic order)."""
    n, h, w, dt, tonemap, out, tm = FULL[name]
    isp = make_isp(dt, moving_alpha=0.1)
    metrics = None
    for step in range(4):
        dev, host = synth_packed(n, h, w, seed=100 * step + 7)
        got = isp.process_packed12(dev, tonemap=tonemap, dtype=out, **tm)
        if step in (0, 3):
            torch.cuda.synchronize()
        exp, metrics = c_oracle.process(host, "RGGB", dt == "f16", out, tonemap, None, tm.get("gamma", 1.0), tm.get("intensity", 1.0),
                                        tm.get("light_adapt", 1.0), tm.get("color_adapt", 0.0), stride=8,
                                        alpha=0.0 if step == 0 else 0.9, metrics=metrics)
        np.testing.assert_allclose(isp.metrics.cpu().numpy(), metrics, rtol=2e-5, atol=2e-6, err_msg=f"{name} step {step}")
        if step in (0, 3):
            compare_with_c_oracle(f"{name} step {step}", got, exp)


@pytest.mark.parametrize("pattern", ["GRBG", "GBRG", "BGGR"])
@pytest.mark.parametrize("ccm", [False, True])
def test_full_frame_patterns_and_ccm_vs_c_oracle(cuda, pattern, ccm):
    """one 4096x3000 frame for the other three CFA patterns, with and without the colour-correction matrix"""
    isp = make_isp("f32", bayer_pattern=pattern, correct_colors=ccm)
    dev, host = synth_packed(1, 3000, 4096, seed=31, pattern=pattern)
    got = isp.process_packed12(dev, tonemap="reinhard", gamma=0.9, intensity=3.0, light_adapt=0.9)
    exp, metrics = c_oracle.process(host, pattern, False, "u8", "reinhard", (O.DEFAULT_CC * O.DEFAULT_WB) if ccm else None,
                                    0.9, 3.0, 0.9, 0.0)
    np.testing.assert_allclose(isp.metrics.cpu().numpy(), metrics, rtol=2e-5, atol=2e-6)
    compare_with_c_oracle(f"{pattern} ccm={ccm}", got, exp)

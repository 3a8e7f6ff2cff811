"""GPU: the BASELINE.json configurations at FULL size, checked through size-independent properties (the oracle is
too slow there): mosaic / demosaic round trips, constant and linearity properties, cross-checks between the fused
sweep and the staged kernels, determinism."""
import numpy as np
import pytest
import torch

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

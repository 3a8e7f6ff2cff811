"""GPU parity of the 10-bit MIPI RAW10 extension (csrc/pack.cu b200isp_decode10 / b200isp_encode10, packed.decode10 / encode10,
ISP.load_packed10) against the oracle restatement of the same layout -- bit-exact.  No reference counterpart (SURVEY 8f-4:
"10-bit packed (not in reference)"), so this parity is unpinned; the layout itself is pinned by a hand-written vector in
tests/test_oracle_packed10.py."""
import numpy as np
import pytest
import torch

from oracle import isp_oracle as O
from tests.util import rng, random_plane, to_cuda, to_np, assert_close_float

pytestmark = pytest.mark.gpu
DTYPES = ("u8", "u16", "i16", "f16", "f32")


@pytest.mark.parametrize("n", [0, 4, 12, 16, 20, 64, 4096, 100004])
def test_encode_decode_raw(cuda, n):
    from taichi_image_b200 import packed
    x = rng(n).integers(0, 1024, size=n).astype(np.uint16)
    e = packed.encode10(to_cuda(x))
    assert e.dtype == torch.uint8 and e.numel() == n * 5 // 4
    assert np.array_equal(to_np(e), O.encode10(x))
    d = packed.decode10(e)
    assert d.dtype == torch.uint16 and np.array_equal(to_np(d), x)


@pytest.mark.parametrize("name", DTYPES)
@pytest.mark.parametrize("scaled", [False, True])
def test_decode10_dtypes(cuda, name, scaled):
    from taichi_image_b200 import packed, dtypes
    enc = rng(3).integers(0, 256, size=(37, 5 * 52), dtype=np.uint8)
    got = packed.decode10(to_cuda(enc), dtype=getattr(dtypes, name), scaled=scaled)
    exp = O.decode10(enc, name, scaled)
    assert tuple(got.shape) == exp.shape == (37, 4 * 52)
    if name in ("f16", "f32"):
        assert np.array_equal(to_np(got).view(np.uint16 if name == "f16" else np.uint32), exp.view(np.uint16 if name == "f16" else np.uint32))
    else:
        assert np.array_equal(to_np(got), exp)


@pytest.mark.parametrize("name", DTYPES)
def test_encode10_scaled(cuda, name):
    from taichi_image_b200 import packed
    x = random_plane(rng(5), (19, 48), name)
    got = packed.encode10(to_cuda(x), scaled=True)
    assert np.array_equal(to_np(got), O.encode10(x, scaled=True))


def test_scaled_roundtrip_all_codes(cuda):
    """every 10-bit code survives decode (scaled, f32) -> encode (scaled)"""
    from taichi_image_b200 import packed, dtypes
    x = np.arange(1024, dtype=np.uint16)
    e = packed.encode10(to_cuda(x))
    f = packed.decode10(e, dtype=dtypes.f32, scaled=True)
    assert float(f.max()) == 1.0 and float(f.min()) == 0.0
    assert torch.equal(packed.encode10(f, scaled=True), e)


def test_unaligned_views_take_the_scalar_kernels(cuda):
    from taichi_image_b200 import packed
    x = rng(6).integers(0, 1024, size=4 * 301).astype(np.uint16)
    buf = torch.empty(5 * 301 + 1, dtype=torch.uint8, device="cuda")
    enc = O.encode10(x)
    buf[1:] = to_cuda(enc)
    assert np.array_equal(to_np(packed.decode10(buf[1:])), x)


def test_numpy_input_and_shape_errors(cuda):
    from taichi_image_b200 import packed
    x = rng(7).integers(0, 1024, size=(6, 40)).astype(np.uint16)
    e = packed.encode10(x)
    assert isinstance(e, np.ndarray) and e.shape == (6, 50) and np.array_equal(packed.decode10(e), x)
    with pytest.raises(AssertionError):
        packed.encode10(np.zeros(6, np.uint16))
    with pytest.raises(AssertionError):
        packed.decode10(np.zeros(8, np.uint8))


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("pattern", ["RGGB", "GBRG"])
def test_isp_load_packed10(cuda, dt, pattern):
    from taichi_image_b200 import camera_isp, bayer
    isp = getattr(camera_isp, {"f16": "Camera16", "f32": "Camera32"}[dt])(bayer.BayerPattern[pattern])
    ref = O.ISP(dt, pattern)
    raw = O.encode10(rng(8).integers(0, 1024, size=(32, 48)).astype(np.uint16))
    got, exp = isp.load_packed10(to_cuda(raw)), ref.load_packed10(raw)
    assert tuple(got.shape) == exp.shape == (32, 48, 3)
    assert_close_float(to_np(got).astype(np.float32), exp.astype(np.float32), rtol=1e-3, atol=1e-6 if dt == "f32" else 1e-3, what=f"load_packed10 {dt}")
    out = isp.tonemap_reinhard([got], gamma=0.8)
    assert out[0].dtype == torch.uint8 and tuple(out[0].shape) == (32, 48, 3)


@pytest.mark.parametrize("n", [4, 16, 20, 36, 1028])
def test_kernels_stay_inside_their_buffers(cuda, n):
    """guard bands around the encode10 / decode10 outputs (vector kernel of 16 pixels + the 4-pixel tail kernel)"""
    from taichi_image_b200 import packed, dtypes
    x = rng(n).integers(0, 1024, size=n).astype(np.uint16)
    xv = to_cuda(x)
    ebuf = torch.full((n * 5 // 4 + 512,), 0xA5, dtype=torch.uint8, device="cuda")
    enc = ebuf[256:256 + n * 5 // 4]
    packed.encode10_kernel(dtypes.u16)(xv, enc)
    dbuf = torch.full((n + 256,), 0x7777, dtype=torch.int16, device="cuda")
    dec = dbuf.view(torch.uint16)[128:128 + n]
    packed.decode10_kernel(dtypes.u16)(enc, dec)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(enc), O.encode10(x)) and np.array_equal(to_np(dec), x)
    assert bool((ebuf[:256] == 0xA5).all()) and bool((ebuf[256 + n * 5 // 4:] == 0xA5).all())
    assert bool((dbuf[:128] == 0x7777).all()) and bool((dbuf[128 + n:] == 0x7777).all())

"""GPU parity: bayer.py -- mosaic, Malvar demosaic for all four CFA patterns (bit-exact on integers)."""
import numpy as np
import pytest

from oracle import isp_oracle as O
from tests.util import rng, random_plane, to_cuda, to_np, assert_close_float

pytestmark = pytest.mark.gpu
PATTERNS = O.PATTERNS
CCM = (O.DEFAULT_CC * O.DEFAULT_WB).flatten().tolist()


def _pat(name):
    from taichi_image_b200.bayer import BayerPattern
    return BayerPattern[name]


@pytest.mark.parametrize("pattern", PATTERNS)
@pytest.mark.parametrize("name", ["u8", "u16", "f16", "f32"])
def test_rgb_to_bayer(cuda, pattern, name):
    from taichi_image_b200 import bayer
    img = random_plane(rng(7), (18, 26, 3), name)
    got = to_np(bayer.rgb_to_bayer(to_cuda(img), _pat(pattern)))
    assert np.array_equal(got.view(np.uint8), O.rgb_to_bayer(img, pattern).view(np.uint8))


@pytest.mark.parametrize("pattern", PATTERNS)
@pytest.mark.parametrize("name", ["u8", "u16", "i16"])
@pytest.mark.parametrize("shape", [(2, 2), (4, 6), (6, 10), (16, 8), (64, 96), (38, 264), (130, 520)])
def test_demosaic_int_bit_exact(cuda, pattern, name, shape):
    from taichi_image_b200 import bayer
    b = random_plane(rng(hash((pattern, name, shape)) % 2 ** 31), shape, name)
    got = to_np(bayer.bayer_to_rgb(to_cuda(b), _pat(pattern)))
    ref = O.bayer_to_rgb(b, pattern)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.array_equal(got, ref), f"{np.count_nonzero(got != ref)} mismatches, first at {np.argwhere(got != ref)[:4]}"


@pytest.mark.parametrize("pattern", PATTERNS)
@pytest.mark.parametrize("name", ["f16", "f32"])
@pytest.mark.parametrize("shape", [(6, 10), (64, 96), (38, 264)])
def test_demosaic_float(cuda, pattern, name, shape):
    from taichi_image_b200 import bayer
    b = random_plane(rng(11), shape, name)
    got = to_np(bayer.bayer_to_rgb(to_cuda(b), _pat(pattern)))
    ref = O.bayer_to_rgb(b, pattern)
    assert_close_float(got, ref, rtol=1e-3, atol=1e-3 if name == "f16" else 1e-6, what=f"{pattern} {name}")


@pytest.mark.parametrize("pattern", PATTERNS)
@pytest.mark.parametrize("name", ["u8", "u16"])
@pytest.mark.parametrize("shape", [(6, 10), (64, 96)])
def test_demosaic_ccm_int_bit_exact(cuda, pattern, name, shape):
    from taichi_image_b200 import bayer
    b = random_plane(rng(12), shape, name)
    got = to_np(bayer.bayer_to_rgb(to_cuda(b), _pat(pattern), correct_colors=np.array(CCM).reshape(3, 3)))
    ref = O.bayer_to_rgb(b, pattern, CCM)
    assert np.array_equal(got, ref), f"{np.count_nonzero(got != ref)} mismatches"


@pytest.mark.parametrize("in_name,out_name", [("u8", "f32"), ("u16", "u8"), ("u8", "u16"), ("f32", "u8"), ("u16", "f16"), ("f16", "f32")])
@pytest.mark.parametrize("shape", [(6, 10), (64, 96)])
def test_demosaic_mixed_dtypes(cuda, in_name, out_name, shape):
    from taichi_image_b200 import bayer
    b = random_plane(rng(13), shape, in_name)
    got = to_np(bayer.bayer_to_rgb(to_cuda(b), _pat("GRBG"), dtype=out_name))
    ref = O.bayer_to_rgb(b, "GRBG", dtype=out_name)
    if out_name in ("u8", "u16"):
        assert np.array_equal(got, ref)
    else:
        assert_close_float(got, ref, rtol=1e-3, atol=1e-3 if out_name == "f16" else 1e-6)


@pytest.mark.parametrize("pattern", PATTERNS)
@pytest.mark.parametrize("name", ["u8", "u16", "f16", "f32"])
def test_roundtrip_bit_exact(cuda, pattern, name):
    """reference test/bayer.py:56-65 turned into an assertion: re-mosaicing the demosaic returns the CFA"""
    from taichi_image_b200 import bayer
    rgb = random_plane(rng(14), (96, 136, 3), name)
    cfa = bayer.rgb_to_bayer(to_cuda(rgb), _pat(pattern))
    back = bayer.rgb_to_bayer(bayer.bayer_to_rgb(cfa, _pat(pattern)), _pat(pattern))
    assert np.array_equal(to_np(back).view(np.uint8), to_np(cfa).view(np.uint8))


@pytest.mark.parametrize("pattern", PATTERNS)
def test_constant_image_stays_constant(cuda, pattern):
    from taichi_image_b200 import bayer
    b = np.full((32, 40), 77, np.uint8)
    assert np.all(to_np(bayer.bayer_to_rgb(to_cuda(b), _pat(pattern))) == 77)


def test_kernel_factory_and_errors(cuda):
    import torch
    from taichi_image_b200 import bayer, u8
    b = random_plane(rng(15), (16, 24), "u8")
    f = bayer.bayer_to_rgb_kernel(_pat("RGGB"), None, u8, u8)
    out = torch.empty((16, 24, 3), dtype=torch.uint8, device="cuda")
    f(to_cuda(b), out)
    assert np.array_equal(to_np(out), O.bayer_to_rgb(b, "RGGB"))
    with pytest.raises(AssertionError):
        bayer.bayer_to_rgb(np.zeros((3, 4), np.uint8))
    with pytest.raises(AssertionError):
        bayer.rgb_to_bayer(np.zeros((4, 4), np.uint8))


def test_tables_match_reference_layout():
    from taichi_image_b200 import bayer
    for k in range(4):
        assert [tuple(w) for _, w in bayer.bayer_kernels[k]] == [tuple(w) for _, w in O.MALVAR[k]]
        assert [o for o, _ in bayer.bayer_kernels[k]] == O.DIAMOND_OFFSETS


@pytest.mark.parametrize("name", ["u8", "u16", "f32"])
@pytest.mark.parametrize("pattern", O.PATTERNS)
def test_bilinear_demosaic_extension(cuda, name, pattern):
    """EXTENSION (north_star; no reference counterpart): bit-exact against the oracle's restatement of the same rule,
    constant images stay constant, and the CFA samples themselves are reproduced exactly"""
    from taichi_image_b200 import bayer
    r = rng(90)
    cfa = random_plane(r, (38, 52), name)
    got = to_np(bayer.bayer_to_rgb(to_cuda(cfa), bayer.BayerPattern[pattern], method="bilinear"))
    ref = O.bayer_to_rgb_bilinear(cfa, pattern)
    if name == "f32":
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7)
    else:
        assert np.array_equal(got, ref)
    assert np.array_equal(to_np(bayer.rgb_to_bayer(to_cuda(got), bayer.BayerPattern[pattern])), cfa) or name == "f32"
    value = O.NP_DTYPE[name](0.25 if name == "f32" else 77)
    const = np.full((8, 8), value)
    out = to_np(bayer.bayer_to_rgb(to_cuda(const), bayer.BayerPattern[pattern], method="bilinear"))
    assert np.all(out == value)

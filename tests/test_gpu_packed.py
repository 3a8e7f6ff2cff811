"""GPU parity: packed.py (encode12 / decode12 / decode16) -- bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from oracle import isp_oracle as O
from tests.util import rng, random_plane, to_cuda, to_np

pytestmark = pytest.mark.gpu
DTYPES = ("u8", "u16", "i16", "f16", "f32")


def test_roundtrip_like_reference(cuda):
    """reference test/packed.py:6-15: 100 random even lengths < 2000, values < 4096"""
    from taichi_image_b200.packed import encode12, decode12
    r = rng(1)
    for _ in range(100):
        size = int(r.integers(1000)) * 2
        x = r.integers(0, 2 ** 12, size=size).astype(np.uint16)
        assert np.all(decode12(encode12(x)) == x)


@pytest.mark.parametrize("ids", [False, True])
@pytest.mark.parametrize("n", [0, 2, 30, 32, 34, 64, 4096, 100002])
def test_encode_decode_raw(cuda, ids, n):
    from taichi_image_b200 import packed
    x = rng(n).integers(0, 4096, size=n).astype(np.uint16)
    e = packed.encode12(to_cuda(x), ids_format=ids)
    assert np.array_equal(to_np(e), O.encode12(x, ids_format=ids))
    d = packed.decode12(e, ids_format=ids)
    assert d.dtype == torch.uint16
    assert np.array_equal(to_np(d), O.decode12(O.encode12(x, ids_format=ids), ids_format=ids))


@pytest.mark.parametrize("name", DTYPES)
@pytest.mark.parametrize("scaled", [False, True])
@pytest.mark.parametrize("ids", [False, True])
def test_decode12_dtypes(cuda, name, scaled, ids):
    from taichi_image_b200 import packed
    enc = rng(3).integers(0, 256, size=(37, 3 * 50), dtype=np.uint8)
    got = to_np(packed.decode12(to_cuda(enc), dtype=name, scaled=scaled, ids_format=ids))
    ref = O.decode12(enc, name, scaled, ids)
    assert got.shape == ref.shape == (37, 100)
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), f"{name} scaled={scaled} ids={ids}"


@pytest.mark.parametrize("name", DTYPES)
@pytest.mark.parametrize("ids", [False, True])
def test_encode12_scaled(cuda, name, ids):
    from taichi_image_b200 import packed
    r = rng(4)
    x = random_plane(r, (23, 66), name)
    got = to_np(packed.encode12(to_cuda(x), scaled=True, ids_format=ids))
    ref = O.encode12(x, scaled=True, ids_format=ids)
    assert got.shape == ref.shape == (23, 99)
    assert np.array_equal(got, ref)


def test_scaled_roundtrip_all_codes(cuda):
    """decode12(scaled) -> f32 -> encode12(scaled) is the identity on all 4096 codes"""
    from taichi_image_b200 import packed
    x = np.arange(4096, dtype=np.uint16)
    e = O.encode12(x)
    f = packed.decode12(to_cuda(e), dtype="f32", scaled=True)
    e2 = to_np(packed.encode12(f, scaled=True))
    assert np.array_equal(e2, e)


@pytest.mark.parametrize("name", DTYPES)
@pytest.mark.parametrize("scaled", [False, True])
def test_decode16(cuda, name, scaled):
    from taichi_image_b200 import packed
    enc = rng(5).integers(0, 256, size=(11, 2 * 77), dtype=np.uint8)
    got = to_np(packed.decode16(to_cuda(enc), dtype=name, scaled=scaled))
    ref = O.decode16(enc, name, scaled)
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))


def test_numpy_and_cpu_tensor_inputs(cuda):
    from taichi_image_b200 import packed
    x = rng(6).integers(0, 4096, size=(8, 64)).astype(np.uint16)
    e_np = packed.encode12(x)
    assert isinstance(e_np, np.ndarray) and np.array_equal(e_np, O.encode12(x))
    e_t = packed.encode12(torch.from_numpy(x))
    assert isinstance(e_t, torch.Tensor) and e_t.device.type == "cpu" and np.array_equal(e_t.numpy(), O.encode12(x))


def test_shape_errors(cuda):
    from taichi_image_b200 import packed
    with pytest.raises(AssertionError):
        packed.encode12(np.zeros(3, np.uint16))
    with pytest.raises(AssertionError):
        packed.decode12(np.zeros(4, np.uint8))
    with pytest.raises(AssertionError):
        packed.decode12(np.zeros(3, np.uint16))

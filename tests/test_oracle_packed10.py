"""CPU: the oracle's MIPI RAW10 restatement (EXTENSION) against a hand-written vector of the layout --
bytes 0..3 = bits 9..2 of pixels 0..3, byte 4 = their bits 1..0 (pixel 0 in the lowest bit pair) -- and its own round trips."""
import numpy as np

from oracle import isp_oracle as O


def test_layout_vector():
    px = np.array([0x3FF, 0x001, 0x2A5, 0x100, 0, 1, 2, 3], np.uint16)
    exp = np.array([0xFF, 0x00, 0xA9, 0x40, 0b00010111, 0, 0, 0, 0, 0b11100100], np.uint8)
    assert np.array_equal(O.encode10(px), exp)
    assert np.array_equal(O.decode10(exp), px)


def test_roundtrip_and_scaling():
    r = np.random.default_rng(0)
    x = r.integers(0, 1024, size=(7, 64)).astype(np.uint16)
    e = O.encode10(x)
    assert e.shape == (7, 80) and np.array_equal(O.decode10(e), x)
    f = O.decode10(e, "f32", scaled=True)
    assert f.dtype == np.float32 and f.max() <= 1.0 and np.array_equal(O.encode10(f, scaled=True), e)
    assert np.array_equal(O.decode10(e, "u8", scaled=True), (x.astype(np.float32) * np.float32(255.0 / 1023.0)).astype(np.uint8))

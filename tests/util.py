"""Shared helpers for the parity tests (oracle = checker only)."""
import numpy as np
import torch

from oracle import isp_oracle as O

INT = ("u8", "u16", "i16")


def rng(seed=0):
    return np.random.default_rng(seed)


def random_plane(r, shape, name):
    if name == "u8":
        return r.integers(0, 256, size=shape, dtype=np.uint8)
    if name == "u16":
        return r.integers(0, 65536, size=shape).astype(np.uint16)
    if name == "i16":
        return r.integers(0, 32768, size=shape).astype(np.int16)
    return r.random(shape, dtype=np.float32).astype(O.NP_DTYPE[name])


def smooth_rgb(r, h, w, noise=0.02):
    """SURVEY 8d synthetic frame: low-frequency gradients x per-channel gains + uniform noise, clipped."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    base = 0.5 + 0.35 * np.sin(x / max(w, 1) * 5.1 + 0.3) * np.cos(y / max(h, 1) * 3.7)
    img = np.stack([base * 0.9, base * 1.0, base * 0.7], -1) + 0.05
    img = img + noise * (r.random((h, w, 3), dtype=np.float32) - 0.5)
    return np.clip(img, 0, 1).astype(np.float32)


def packed_frame(r, h, w, pattern="RGGB", smooth=True):
    if smooth:
        cfa = O.rgb_to_bayer(smooth_rgb(r, h, w), pattern)
        return O.encode12(cfa, scaled=True)
    return O.encode12(r.integers(0, 4096, size=(h, w)).astype(np.uint16))


def to_cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def to_np(t):
    return t.detach().cpu().numpy()


def assert_close_int(got, ref, max_lsb=0, what=""):
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= max_lsb, f"{what}: max |diff| = {d.max()} LSB (allowed {max_lsb}), {np.count_nonzero(d)} / {d.size} differ"
    return float(np.count_nonzero(d)) / max(d.size, 1)


def assert_close_float(got, ref, rtol=1e-3, atol=1e-6, what=""):
    g, r = got.astype(np.float64), ref.astype(np.float64)
    err = np.abs(g - r) - (atol + rtol * np.abs(r))
    assert err.max() <= 0, f"{what}: max abs err {np.abs(g - r).max():.3e} (rtol {rtol}, atol {atol})"


def assert_u16_from_f16_isp(got, ref, gain, what=""):
    """Camera16 -> u16 (documented deviation, profiles/r02_error_histogram.txt): the ISP stores its intermediates as
    f16; a 1-ulp f32 disagreement in front of such a store (MUFU pow / reciprocal, FMA contraction of the CCM) flips
    the f16 rounding of ~1e-4 of the values by ONE f16 ulp = 2^-11 relative = up to 32 LSB of u16 times the tone
    map's gain.  Asserted: at least 99.7 % of the values within 1 LSB (measured worst 1.2e-3 on the small test frames), the rest within one f16 ulp seen through
    ``gain`` (>= 1: 1 / (max - min) for the linear map, 1 / max_out for Reinhard, times the slope of x^(1/gamma))."""
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    frac = float(np.count_nonzero(d > 1)) / d.size
    bound = int(32.0 * gain) + 2
    assert frac <= 3e-3, f"{what}: {frac:.2e} of the values differ by more than 1 LSB"
    assert d.max() <= bound, f"{what}: max |diff| = {d.max()} LSB > one f16 ulp through the tone map ({bound})"

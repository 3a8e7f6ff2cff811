"""numpy float32 restatement of the reference camera-ISP kernels.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function follows one
reference function (uc-vision/taichi_image 0.3.2, paths relative to
/root/reference/taichi_image) and reproduces its operation order and rounding
points: all arithmetic is float32 (Taichi default_fp), float->int casts
truncate toward zero, ``ti.round`` rounds half away from zero, stores to f16
round to nearest even.  Global float sums (metering) are accumulated in
float64 and rounded once -- the reference's atomics have unspecified order, so
the sum is only defined to ~1e-6 relative anyway.

Pinning: packed round trip (test/packed.py:6-15) and the golden vectors under
tests/golden/ (reference source executed through oracle/taichi_shim).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

# types.py:12-18
SCALE = {"u8": 255, "u16": 65535, "i16": 32767, "f16": 1.0, "f32": 1.0}
NP_DTYPE = {"u8": np.uint8, "u16": np.uint16, "i16": np.int16,
            "f16": np.float16, "f32": np.float32}
NAME_OF = {np.dtype(v): k for k, v in NP_DTYPE.items()}


def dtype_name(arr_or_dtype) -> str:
    dt = arr_or_dtype.dtype if hasattr(arr_or_dtype, "dtype") else np.dtype(arr_or_dtype)
    return NAME_OF[np.dtype(dt)]


def cast_to(x: np.ndarray, name: str) -> np.ndarray:
    """ti.cast(f32 value, dtype): truncation for ints, RNE for f16."""
    if name in ("u8", "u16", "i16"):
        info = np.iinfo(NP_DTYPE[name])
        # reference behaviour is UB outside the range; the product saturates
        #  and NaN -> 0.
        y = np.nan_to_num(np.trunc(x.astype(np.float64)), nan=0.0,
                          posinf=info.max, neginf=info.min)
        return np.clip(y, info.min, info.max).astype(NP_DTYPE[name])
    return x.astype(NP_DTYPE[name])


def round_half_away(x: np.ndarray) -> np.ndarray:
    return np.sign(x) * np.floor(np.abs(x) + f32(0.5))


# --------------------------------------------------------------------------
# packed.py
# --------------------------------------------------------------------------
def encode12(values: np.ndarray, scaled: bool = False, ids_format: bool = False) -> np.ndarray:
    """packed.py:12-20 (standard), :47-55 (IDS), :59-89, :176-185."""
    shape = values.shape
    assert shape[-1] % 2 == 0
    flat = values.reshape(-1)
    if scaled:  # packed.py:66-69
        scale = SCALE[dtype_name(flat)]
        v = flat.astype(f32) * f32(4095.0 / scale)
        p = round_half_away(v).astype(np.int64).astype(np.uint16)
    else:
        p = flat.astype(np.uint16)
    p0 = p[0::2].astype(np.uint32)
    p1 = p[1::2].astype(np.uint32)
    out = np.empty((p0.size, 3), np.uint8)
    if not ids_format:
        out[:, 0] = (p0 & 0xFF).astype(np.uint8)
        out[:, 1] = ((((p1 & 0xF) << 4) | (p0 >> 8)) & 0xFF).astype(np.uint8)
        out[:, 2] = ((p1 >> 4) & 0xFF).astype(np.uint8)
    else:
        out[:, 0] = ((p0 >> 4) & 0xFF).astype(np.uint8)
        out[:, 1] = ((p1 >> 4) & 0xFF).astype(np.uint8)
        out[:, 2] = ((((p0 & 0xF) << 4) | (p1 & 0xF)) & 0xFF).astype(np.uint8)
    return out.reshape(shape[:-1] + (shape[-1] * 3 // 2,))


def decode12_raw(enc: np.ndarray, ids_format: bool = False) -> np.ndarray:
    """packed.py:23-31 (standard), :36-44 (IDS) -> u16 array."""
    b = enc.reshape(-1, 3).astype(np.uint16)
    out = np.empty((b.shape[0], 2), np.uint16)
    if not ids_format:
        out[:, 0] = ((b[:, 1] & 0xF) << 8) | b[:, 0]
        out[:, 1] = (b[:, 2] << 4) | (b[:, 1] >> 4)
    else:
        out[:, 0] = (b[:, 0] << 4) | (b[:, 2] & 0xF)
        out[:, 1] = (b[:, 1] << 4) | (b[:, 2] >> 4)
    return out.reshape(-1)


def decode12(enc: np.ndarray, dtype: str = "u16", scaled: bool = False,
             ids_format: bool = False) -> np.ndarray:
    """packed.py:91-131, :188-198."""
    shape = enc.shape
    assert enc.dtype == np.uint8 and shape[-1] % 3 == 0
    v = decode12_raw(enc, ids_format)
    if scaled:  # packed.py:98-100 -- multiply by the f32 constant scale/4095
        out = cast_to(v.astype(f32) * f32(SCALE[dtype] / 4095.0), dtype)
    else:       # packed.py:103-104 -- arr[i] = value (implicit cast)
        out = v.astype(NP_DTYPE[dtype])
    return out.reshape(shape[:-1] + (shape[-1] * 2 // 3,))


def repack12_ids(enc: np.ndarray) -> np.ndarray:
    """IDS layout -> standard layout (extension of the framework): packed.py:36-44 followed by :12-20"""
    return encode12(decode12_raw(enc, ids_format=True), ids_format=False).reshape(enc.shape)


def decode16(enc: np.ndarray, dtype: str = "u16", scaled: bool = False) -> np.ndarray:
    """packed.py:134-172, :200-210 (the ids_format kwarg bug, SURVEY Q2, is not reproduced)."""
    shape = enc.shape
    assert enc.dtype == np.uint8 and shape[-1] % 2 == 0
    b = enc.reshape(-1, 2).astype(np.uint16)
    v = (b[:, 1] << 8) | b[:, 0]
    if scaled:
        out = cast_to(v.astype(f32) * f32(SCALE[dtype] / 65535.0), dtype)
    else:
        out = v.astype(NP_DTYPE[dtype])
    return out.reshape(shape[:-1] + (shape[-1] // 2,))


# EXTENSION (SURVEY 8f-4 "10-bit packed"; the reference has no 10-bit format -> parity unpinned): MIPI CSI-2 RAW10,
# 5 bytes <-> 4 pixels, bytes 0..3 = bits 9..2 of pixels 0..3, byte 4 = their bits 1..0 (pixel 0 in the lowest bit pair);
# value conventions of packed.py:66-73 / :98-104 with 1023 in place of 4095.
def encode10(values: np.ndarray, scaled: bool = False) -> np.ndarray:
    shape = values.shape
    assert shape[-1] % 4 == 0
    flat = values.reshape(-1)
    if scaled:
        v = flat.astype(f32) * f32(1023.0 / SCALE[dtype_name(flat)])
        p = round_half_away(v).astype(np.int64).astype(np.uint16)
    else:
        p = flat.astype(np.uint16)
    p = (p & 0x3FF).reshape(-1, 4)
    out = np.empty((p.shape[0], 5), np.uint8)
    out[:, :4] = (p >> 2).astype(np.uint8)
    out[:, 4] = ((p[:, 0] & 3) | ((p[:, 1] & 3) << 2) | ((p[:, 2] & 3) << 4) | ((p[:, 3] & 3) << 6)).astype(np.uint8)
    return out.reshape(shape[:-1] + (shape[-1] * 5 // 4,))


def decode10(enc: np.ndarray, dtype: str = "u16", scaled: bool = False) -> np.ndarray:
    shape = enc.shape
    assert enc.dtype == np.uint8 and shape[-1] % 5 == 0
    b = enc.reshape(-1, 5).astype(np.uint16)
    v = np.stack([(b[:, j] << 2) | ((b[:, 4] >> (2 * j)) & 3) for j in range(4)], -1).reshape(-1)
    if scaled:
        out = cast_to(v.astype(f32) * f32(SCALE[dtype] / 1023.0), dtype)
    else:
        out = v.astype(NP_DTYPE[dtype])
    return out.reshape(shape[:-1] + (shape[-1] * 4 // 5,))


# --------------------------------------------------------------------------
# bayer.py
# --------------------------------------------------------------------------
PATTERNS = ("RGGB", "GRBG", "GBRG", "BGGR")           # bayer.py:75-79 (values 0..3)
PIXEL_ORDER = {"RGGB": (0, 1, 1, 2), "GRBG": (1, 0, 2, 1),  # bayer.py:85-90
               "GBRG": (1, 2, 0, 1), "BGGR": (2, 1, 1, 0)}
KERNEL_PATTERN = {"RGGB": (0, 1, 2, 3), "GBRG": (1, 0, 3, 2),  # bayer.py:92-97
                  "GRBG": (2, 3, 0, 1), "BGGR": (3, 2, 1, 0)}


def _expand_quarter(q):
    """kernel.py:3-12 -- quarter kernel [[a],[b,c],[d,e,f]] -> 13 diamond taps
    in the order of bayer.py:15-27: rows -2..2, columns ascending."""
    (a,), (b, c), (d, e, f) = q
    return [a, b, c, b, d, e, f, e, d, b, c, b, a]


DIAMOND_OFFSETS = [(-2, 0), (-1, -1), (-1, 0), (-1, 1), (0, -2), (0, -1), (0, 0),
                   (0, 1), (0, 2), (1, -1), (1, 0), (1, 1), (2, 0)]


def malvar_kernels():
    """bayer.py:30-55: four site kernels, each a list of (offset, (wR,wG,wB))."""
    g_rb = _expand_quarter([(-2,), (0, 4), (-2, 4, 8)])
    r_g1 = _expand_quarter([(-2,), (-2, 8), (1, 0, 10)])
    r_g2 = _expand_quarter([(1,), (-2, 0), (-2, 8, 10)])
    rb_br = _expand_quarter([(-3,), (4, 0), (-3, 0, 12)])
    ident = _expand_quarter([(0,), (0, 0), (0, 0, 16)])
    b_g1, b_g2 = r_g2, r_g1
    sites = [(ident, g_rb, rb_br), (r_g1, ident, b_g1), (r_g2, ident, b_g2), (rb_br, g_rb, ident)]
    return [list(zip(DIAMOND_OFFSETS, zip(*s))) for s in sites]


MALVAR = malvar_kernels()


def rgb_to_bayer(image: np.ndarray, pattern: str = "RGGB") -> np.ndarray:
    """bayer.py:101-112, :193-198."""
    assert image.ndim == 3 and image.shape[2] == 3
    h, w = image.shape[:2]
    out = np.zeros((h, w), image.dtype)
    p1, p2, p3, p4 = PIXEL_ORDER[pattern]
    he, we = h // 2 * 2, w // 2 * 2
    out[0:he:2, 0:we:2] = image[0:he:2, 0:we:2, p1]
    out[0:he:2, 1:we:2] = image[0:he:2, 1:we:2, p2]
    out[1:he:2, 0:we:2] = image[1:he:2, 0:we:2, p3]
    out[1:he:2, 1:we:2] = image[1:he:2, 1:we:2, p4]
    return out


def _ccm_apply(c: np.ndarray, ccm) -> np.ndarray:
    """mat3 @ vec3 in f32, row dot products left to right (bayer.py:152-153)."""
    m = np.asarray(ccm, dtype=np.float64).reshape(3, 3).astype(f32)
    out = np.empty_like(c)
    for r in range(3):
        out[..., r] = (c[..., 0] * m[r, 0] + c[..., 1] * m[r, 1]) + c[..., 2] * m[r, 2]
    return out


def demosaic_unit(bayer: np.ndarray, pattern: str = "RGGB", ccm=None) -> np.ndarray:
    """bayer.py:137-155 -- f32 (H,W,3) in [0,1] before the output scaling."""
    assert bayer.ndim == 2 and bayer.shape[0] % 2 == 0 and bayer.shape[1] % 2 == 0
    h, w = bayer.shape
    in_scale = f32(SCALE[dtype_name(bayer)])
    pad = np.zeros((h + 4, w + 4), f32)
    pad[2:-2, 2:-2] = bayer.astype(f32)
    valid = np.zeros((h + 4, w + 4), bool)
    valid[2:-2, 2:-2] = True
    out = np.zeros((h, w, 3), f32)
    kidx = KERNEL_PATTERN[pattern]
    for slot in range(4):
        r0, c0 = slot & 1, slot >> 1        # slot k = (row&1) + 2*(col&1), bayer.py:162-175
        taps = MALVAR[kidx[slot]]
        hh, ww = h // 2, w // 2
        c = np.zeros((hh, ww, 3), f32)
        t = np.zeros((hh, ww, 3), f32)
        for (d0, d1), wt in taps:
            v = pad[2 + r0 + d0: 2 + r0 + d0 + h: 2, 2 + c0 + d1: 2 + c0 + d1 + w: 2]
            m = valid[2 + r0 + d0: 2 + r0 + d0 + h: 2, 2 + c0 + d1: 2 + c0 + d1 + w: 2]
            wv = np.asarray(wt, f32)
            c = c + np.where(m[..., None], v[..., None] * wv, f32(0))
            t = t + np.where(m[..., None], wv, f32(0))
        c = c / (in_scale * t)
        if ccm is not None:
            c = _ccm_apply(c, ccm)
        out[r0::2, c0::2] = np.clip(c, f32(0), f32(1))
    return out


def bayer_to_rgb(bayer: np.ndarray, pattern: str = "RGGB", ccm=None, dtype: str | None = None) -> np.ndarray:
    """bayer.py:114-177, :202-219: demosaic + cast(v*out_scale) (truncating)."""
    dtype = dtype or dtype_name(bayer)
    c = demosaic_unit(bayer, pattern, ccm)
    return cast_to(c * f32(SCALE[dtype]), dtype)


# --------------------------------------------------------------------------
# util.py / color
# --------------------------------------------------------------------------
def lerp(t, a, b):
    """util.py:82-84."""
    return a + t * (b - a)


def rgb_gray(rgb: np.ndarray) -> np.ndarray:
    """color/__init__.py:7-10 (f32 dot, left to right)."""
    return (rgb[..., 0] * f32(0.299) + rgb[..., 1] * f32(0.587)) + rgb[..., 2] * f32(0.114)


def bounds(image: np.ndarray):
    """util.py:49-60."""
    x = image.astype(f32)
    return f32(x.min()), f32(x.max())


# --------------------------------------------------------------------------
# tonemap.py (stand-alone, per image)
# --------------------------------------------------------------------------
def _pow(x, y):
    with np.errstate(all="ignore"):
        return np.power(x.astype(f32) if hasattr(x, "astype") else f32(x), f32(y)).astype(f32)


def linear_func(image: np.ndarray, bmin, bmax, gamma, scale_factor, dtype: str) -> np.ndarray:
    """tonemap.py:11-17."""
    with np.errstate(all="ignore"):
        inv_range = f32(1) / (f32(bmax) - f32(bmin))
        x = _pow((image.astype(f32) - f32(bmin)) * inv_range, f32(1) / f32(gamma))
        y = np.clip(x, f32(0), f32(1)) * f32(scale_factor)
    return cast_to(y, dtype)


def tonemap_linear(src: np.ndarray, gamma: float = 1.0, dtype: str = "u8") -> np.ndarray:
    """tonemap.py:26-46."""
    bmin, bmax = bounds(src)
    return linear_func(src, bmin, bmax, gamma, SCALE[dtype], dtype)


def metering_standalone(temp: np.ndarray):
    """tonemap.py:77-103 with Bounds(0,1); returns (b_min, b_max, log_mean, gray_mean, rgb_mean)
    where (b_min, b_max) = (log_min, -log_max) as written at :102."""
    scaled = (temp - f32(0)) / (f32(1) - f32(0))
    gray = rgb_gray(scaled)
    log_gray = np.log(np.maximum(gray, f32(1e-4))).astype(f32)
    n = temp.shape[0] * temp.shape[1]
    return (f32(log_gray.min()), f32(-log_gray.max()),
            f32(log_gray.astype(np.float64).sum() / n), f32(gray.astype(np.float64).sum() / n),
            (scaled.astype(np.float64).sum(axis=(0, 1)) / n).astype(f32))


def reinhard_map(scaled: np.ndarray, b_min, b_max, log_mean, gray_mean, rgb_mean,
                 intensity, light_adapt, color_adapt) -> np.ndarray:
    """Shared per-pixel photoreceptor map: tonemap.py:107-131 == camera_isp.py:192-210."""
    with np.errstate(all="ignore"):
        key = (f32(b_max) - f32(log_mean)) / (f32(b_max) - f32(b_min))
        map_key = f32(0.3) + f32(0.7) * _pow(key, 1.4)
        ca, la = f32(color_adapt), f32(light_adapt)
        mean = f32(gray_mean) + ca * (np.asarray(rgb_mean, f32) - f32(gray_mean))   # lerp(ca, mean, rgb_mean)
        gray = rgb_gray(scaled)
        adapt_color = gray[..., None] + ca * (scaled - gray[..., None])
        adapt_mean = mean + la * (adapt_color - mean)
        adapt = _pow(np.exp(-f32(intensity)).astype(f32) * adapt_mean, map_key)
        return scaled * (f32(1.0) / (adapt + scaled))


def tonemap_reinhard(src: np.ndarray, gamma=1.0, intensity=1.0, light_adapt=1.0,
                     color_adapt=0.0, dtype: str = "u8") -> np.ndarray:
    """tonemap.py:134-168 (five dependent passes)."""
    bmin, bmax = bounds(src)
    temp = linear_func(src, bmin, bmax, 1.0, 1.0, "f32")
    stats = metering_standalone(temp)
    temp = reinhard_map(temp, *stats, intensity, light_adapt, color_adapt).astype(f32)
    b2min, b2max = bounds(temp)
    return linear_func(temp, b2min, b2max, gamma, SCALE[dtype], dtype)


# --------------------------------------------------------------------------
# interpolate.py
# --------------------------------------------------------------------------
def resize_bilinear(src: np.ndarray, size, scale, dtype: str | None = None) -> np.ndarray:
    """interpolate.py:19-34, :59-66, :128-139.  size=(w,h); scale scalar or (row, col).
    p = I/scale, top-left aligned, clamp-to-edge, truncating cast."""
    in_name = dtype_name(src)
    dtype = dtype or in_name
    sr, sc = (scale, scale) if np.isscalar(scale) else scale
    h_out, w_out = int(size[1]), int(size[0])
    hs, ws = src.shape[:2]
    rows = np.arange(h_out, dtype=f32) / f32(sr)
    cols = np.arange(w_out, dtype=f32) / f32(sc)
    r1 = rows.astype(np.int32)
    c1 = cols.astype(np.int32)
    fr = (rows - r1.astype(f32))[:, None, None]
    fc = (cols - c1.astype(f32))[None, :, None]
    ra, rb = np.clip(r1, 0, hs - 1), np.clip(r1 + 1, 0, hs - 1)
    ca, cb = np.clip(c1, 0, ws - 1), np.clip(c1 + 1, 0, ws - 1)
    s = src.astype(f32)
    mix = lambda a, b, t: a * (f32(1) - t) + b * t
    y1 = mix(s[ra][:, ca], s[rb][:, ca], fr)      # along dim 0 first (frac.x)
    y2 = mix(s[ra][:, cb], s[rb][:, cb], fr)
    out = mix(y1, y2, fc)
    intensity = f32(SCALE[dtype] / SCALE[in_name])
    return cast_to(out * intensity, dtype)


TRANSFORMS = ("none", "rotate_90", "rotate_180", "rotate_270", "transpose",
              "flip_horiz", "flip_vert", "transverse")


def transform(src: np.ndarray, name: str) -> np.ndarray:
    """interpolate.py:36-56, :93-125.  rotate_90 is clockwise;
    transverse is the fixed anti-transpose."""
    if name == "none":
        return src.copy()
    if name == "rotate_90":       # dst[r,c] = src[H-1-c, r]
        return np.ascontiguousarray(np.rot90(src, 3, (0, 1)))
    if name == "rotate_180":
        return np.ascontiguousarray(src[::-1, ::-1])
    if name == "rotate_270":      # dst[r,c] = src[c, W-1-r]
        return np.ascontiguousarray(np.rot90(src, 1, (0, 1)))
    if name == "transpose":
        return np.ascontiguousarray(np.swapaxes(src, 0, 1))
    if name == "flip_vert":
        return np.ascontiguousarray(src[::-1])
    if name == "flip_horiz":
        return np.ascontiguousarray(src[:, ::-1])
    if name == "transverse":
        return np.ascontiguousarray(np.swapaxes(src, 0, 1)[::-1, ::-1])
    raise ValueError(name)


# --------------------------------------------------------------------------
# camera_isp.py
# --------------------------------------------------------------------------
DEFAULT_CC = np.array([[1.75, -0.25, -0.30], [-0.10, 1.40, -0.30], [-0.05, -0.55, 2.10]])
DEFAULT_WB = np.array([1.8, 1.0, 2.1])


def metering_update(images, prev: np.ndarray, alpha: float, stride: int = 8) -> np.ndarray:
    """camera_isp.py:142-175: joint two-phase metering with the double bounds blend (Q4)."""
    stack = np.stack([im[::stride, ::stride, :] for im in images], 0).astype(f32)
    a = f32(alpha)
    prev = prev.astype(f32)
    mn, mx = f32(stack.min()), f32(stack.max())
    bmin = mn + a * (prev[0] - mn)            # lerp(alpha, new, prev)
    bmax = mx + a * (prev[1] - mx)
    scaled = (stack - bmin) / (bmax - bmin + f32(1e-6))
    gray = rgb_gray(scaled)
    log_gray = np.log(np.maximum(gray, f32(1e-4))).astype(f32)
    n = f32(stack.shape[0] * stack.shape[1] * stack.shape[2])
    stats = np.array([bmin, bmax, log_gray.min(), log_gray.max(),
                      f32(log_gray.astype(np.float64).sum()) / n,
                      f32(gray.astype(np.float64).sum()) / n,
                      *(scaled.astype(np.float64).sum(axis=(0, 1, 2)).astype(f32) / n)], f32)
    return (stats + a * (prev - stats)).astype(f32)


# -- the same update split at its two exchange points (multi-GPU shared exposure): every rank
#    computes a record over its own images, the records are gathered and folded in rank order.
def _meter_stack(images, stride):
    return np.stack([im[::stride, ::stride, :] for im in images], 0).astype(f32)


def metering_phase1(images, stride: int = 8) -> np.ndarray:
    """camera_isp.py:149-154 over this rank's images: {min, max}"""
    stack = _meter_stack(images, stride)
    return np.array([stack.min(), stack.max()], f32)


def metering_fold_bounds(gathered1, alpha, prev):
    """camera_isp.py:156 with the joint min / max of all ranks"""
    g1 = np.asarray(gathered1, f32).reshape(-1, 2)
    a, prev = f32(alpha), np.asarray(prev, f32)
    mn, mx = f32(g1[:, 0].min()), f32(g1[:, 1].max())
    return mn + a * (prev[0] - mn), mx + a * (prev[1] - mx)


def metering_phase2(images, gathered1, alpha, prev, stride: int = 8) -> np.ndarray:
    """camera_isp.py:158-161 over this rank's images w.r.t. the joint blended bounds:
    {log_min, log_max, sum_log, sum_gray, sum_r, sum_g, sum_b, n}"""
    stack = _meter_stack(images, stride)
    bmin, bmax = metering_fold_bounds(gathered1, alpha, prev)
    scaled = (stack - bmin) / (bmax - bmin + f32(1e-6))
    gray = rgb_gray(scaled)
    log_gray = np.log(np.maximum(gray, f32(1e-4))).astype(f32)
    n = stack.shape[0] * stack.shape[1] * stack.shape[2]
    return np.array([log_gray.min(), log_gray.max(), log_gray.astype(np.float64).sum(), gray.astype(np.float64).sum(),
                     *scaled.astype(np.float64).sum(axis=(0, 1, 2)), n], f32)


def metering_finalize(gathered1, gathered2, alpha, prev) -> np.ndarray:
    """camera_isp.py:131-134, :164-166 with the folded records"""
    g2 = np.asarray(gathered2, f32).reshape(-1, 8)
    a, prev = f32(alpha), np.asarray(prev, f32)
    bmin, bmax = metering_fold_bounds(gathered1, alpha, prev)
    sums = g2[:, 2:].astype(np.float64).sum(0).astype(f32)
    n = sums[5]
    stats = np.array([bmin, bmax, g2[:, 0].min(), g2[:, 1].max(), *(sums[:5] / n)], f32)
    return (stats + a * (prev - stats)).astype(f32)


def isp_reinhard(image: np.ndarray, metrics: np.ndarray, gamma, intensity, light_adapt,
                 color_adapt, out_dtype: str = "u8", return_intermediate=False):
    """camera_isp.py:177-218.  `image` is the ISP-dtype RGB; the f16/f32 rounding of the
    stored intermediate (:211) is applied with image.dtype."""
    m = metrics.astype(f32)
    with np.errstate(all="ignore"):
        scaled = (image.astype(f32) - m[0]) / (m[1] - m[0])
        p = reinhard_map(scaled, m[2], m[3], m[4], m[5], m[6:9], intensity, light_adapt, color_adapt).astype(f32)
        max_out = max(f32(1e-6), f32(np.nanmax(p)))
        stored = p.astype(image.dtype)                                   # :211
        q = _pow(stored.astype(f32) / max_out, 1.0 / gamma)              # :217
        # pixels darker than the (sub-sampled, moving-average) lower bound give negative / NaN q; the
        # reference then casts a negative or NaN float to u8 (undefined).  Defined here as 0.
        q = np.where(q > 0, q, f32(0)).astype(f32)
        if out_dtype in ("u8", "u16", "i16"):
            q = np.minimum(q, f32(1))
        out = cast_to(f32(SCALE[out_dtype]) * q, out_dtype)               # :218 (255*p for u8)
    return (out, stored, max_out) if return_intermediate else out


def isp_linear(image: np.ndarray, metrics: np.ndarray, gamma, out_dtype: str = "u8") -> np.ndarray:
    """camera_isp.py:220-227 -> tonemap.py:11-17 with the EMA bounds."""
    m = metrics.astype(f32)
    return linear_func(image, m[0], m[1], gamma, SCALE[out_dtype], out_dtype)


class ISP:
    """camera_isp.py:237-413 on numpy arrays.  dtype 'f16' = Camera16, 'f32' = Camera32.
    Differences from the reference, all listed in SURVEY 2.5: bayer_pattern is honoured
    (Q1), out_dtype may be u8/u16/f16 (extension, formula of tonemap.py:16-17)."""

    def __init__(self, dtype="f32", bayer_pattern="RGGB", scale=None, resize_width=0,
                 moving_alpha=0.1, correct_colors=False, white_balance=DEFAULT_WB,
                 color_correction=DEFAULT_CC, transform="none", metering_stride=8, demosaic="malvar"):
        assert scale is None or resize_width == 0
        assert demosaic in ("malvar", "bilinear")        # "bilinear": extension, bayer_to_rgb_bilinear below
        self.demosaic = demosaic
        self.dtype, self.bayer_pattern = dtype, bayer_pattern
        self.scale, self.resize_width = scale, resize_width
        self.moving_alpha, self.correct_colors = moving_alpha, correct_colors
        self.white_balance, self.color_correction = white_balance, color_correction
        self.transform, self.metering_stride = transform, metering_stride
        self.metrics = None

    @property
    def color_correct_matrix(self):                      # camera_isp.py:360-369
        if not self.correct_colors:
            return None
        cc = np.array(self.color_correction, dtype=np.float64).copy()
        cc[:, :3] *= np.asarray(self.white_balance, dtype=np.float64)
        return cc

    def resize_image(self, image):                       # camera_isp.py:302-315
        h, w = image.shape[:2]
        if self.resize_width > 0:
            s = self.resize_width / w
            return resize_bilinear(image, (self.resize_width, round(h * s)), s)
        if self.scale is not None:
            return resize_bilinear(image, (round(w * self.scale), round(h * self.scale)), self.scale)
        return image

    def _process_image(self, cfa):                       # camera_isp.py:371-373
        ccm = self.color_correct_matrix
        demosaic = bayer_to_rgb_bilinear if self.demosaic == "bilinear" else bayer_to_rgb
        rgb = demosaic(cfa, self.bayer_pattern, None if ccm is None else ccm.flatten().tolist())
        return self.resize_image(rgb)

    def load_packed12(self, data, ids_format=False):     # camera_isp.py:333-340
        return self._process_image(decode12(data, self.dtype, scaled=True, ids_format=ids_format))

    def load_packed16(self, data):                       # camera_isp.py:342-347
        return self._process_image(decode16(data, self.dtype, scaled=True))

    def load_packed10(self, data):                       # EXTENSION (MIPI RAW10), the path of load_packed16
        return self._process_image(decode10(data, self.dtype, scaled=True))

    def load_16u(self, image):                           # camera_isp.py:82-87, :318-321
        return self._process_image((image.astype(f32) / f32(65535.0)).astype(NP_DTYPE[self.dtype]))

    def load_32f(self, image):                           # camera_isp.py:89-93
        return self._process_image(image.astype(NP_DTYPE[self.dtype]))

    def update_metering(self, images):                   # camera_isp.py:376-385
        if self.metrics is None:
            self.metrics = metering_update(images, np.zeros(9, f32), 0.0, self.metering_stride)
        else:
            self.metrics = metering_update(images, self.metrics, 1.0 - self.moving_alpha, self.metering_stride)

    def tonemap_reinhard(self, images, gamma=1.0, intensity=1.0, light_adapt=1.0, color_adapt=0.0,
                         out_dtype="u8"):
        self.update_metering(images)
        outs = [isp_reinhard(im, self.metrics, gamma, intensity, light_adapt, color_adapt, out_dtype)
                for im in images]
        return [transform(o, self.transform) for o in outs]

    def tonemap_linear(self, images, gamma=1.0, out_dtype="u8"):
        self.update_metering(images)
        outs = [isp_linear(im, self.metrics, gamma, out_dtype) for im in images]
        return [transform(o, self.transform) for o in outs]


# --------------------------------------------------------------------------
# color/yuv_420.py
# --------------------------------------------------------------------------
# yuv_420.py:12-16.  NB the reference applies this matrix to the BGR-swizzled vector (yuv_420.py:24-26), i.e.
# Y = 0.299 B + 0.587 G + 0.114 R: reproduced as is.
YCRCB_T_BGR = np.array([[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5], [0.5, -0.418688, -0.081312]], np.float64)
BGR_T_YCRCB = np.linalg.inv(YCRCB_T_BGR.astype(f32).astype(np.float64)).astype(f32)      # yuv_420.py:18 (Python-scope inverse)


def _mat_vec(m, x):
    """mat3 @ vec3 in f32, row dot products left to right"""
    m = m.astype(f32)
    return np.stack([(m[r, 0] * x[..., 0] + m[r, 1] * x[..., 1]) + m[r, 2] * x[..., 2] for r in range(3)], -1)


def _ref_clamp01(v):
    """tm.clamp(0, 1, v) as the reference writes it (yuv_420.py:57, 61, 88): Taichi's clamp is clamp(x, xmin, xmax),
    so the call evaluates min(max(0, 1), v) = min(1, v) -- no lower clamp."""
    return np.minimum(f32(1.0), v)


def rgb_yuv420(src: np.ndarray, dtype: str | None = None) -> np.ndarray:
    """yuv_420.py:38-64, :104-118: planar Y (H rows) followed by two (H/2, W/2) chroma planes; the second matrix
    row goes to plane 1, the third to plane 0 (yuv_420.py:62-63)."""
    in_name = dtype_name(src)
    dtype = dtype or in_name
    h, w, _ = src.shape
    x = src.astype(f32) / f32(SCALE[in_name])
    yuv = _mat_vec(YCRCB_T_BGR, x[..., ::-1]) + np.array([0, 0.5, 0.5], f32)
    out = np.zeros(((h * 3) // 2, w), NP_DTYPE[dtype])
    out[:h] = cast_to(_ref_clamp01(yuv[..., 0]) * f32(SCALE[dtype]), dtype)
    he, we = h // 2 * 2, w // 2 * 2
    uv = ((yuv[0:he:2, 0:we:2, 1:] + yuv[0:he:2, 1:we:2, 1:]) + yuv[1:he:2, 0:we:2, 1:]) + yuv[1:he:2, 1:we:2, 1:]
    uvq = cast_to(_ref_clamp01(uv / f32(4.0)) * f32(SCALE[dtype]), dtype)
    planes = out[h:].reshape(2, h // 2, w // 2)
    planes[1] = uvq[..., 0]
    planes[0] = uvq[..., 1]
    return out


def yuv420_rgb_unit(yuv: np.ndarray) -> np.ndarray:
    """f32 RGB before the upper clamp and the cast.  Negative components are undefined behaviour in the reference
    for integer outputs (no lower clamp, yuv_420.py:88); the product saturates them to 0."""
    in_name = dtype_name(yuv)
    h = yuv.shape[0] * 2 // 3
    w = yuv.shape[1]
    y = yuv[:h].astype(f32)
    planes = yuv[h:].reshape(2, h // 2, w // 2).astype(f32)
    ii, jj = np.arange(h) // 2, np.arange(w) // 2
    v = np.stack([y, planes[1][ii][:, jj], planes[0][ii][:, jj]], -1) / f32(SCALE[in_name])
    bgr = _mat_vec(BGR_T_YCRCB, v - np.array([0, 0.5, 0.5], f32))
    return bgr[..., ::-1]


def yuv420_rgb(yuv: np.ndarray, dtype: str | None = None) -> np.ndarray:
    """yuv_420.py:66-90, :95-101, :120-131"""
    dtype = dtype or dtype_name(yuv)
    return cast_to(_ref_clamp01(yuv420_rgb_unit(yuv)) * f32(SCALE[dtype]), dtype)


# --------------------------------------------------------------------------
# EXTENSION: bilinear demosaic (no reference counterpart; rule of bayer.py:137-155 with 3x3 kernels)
# --------------------------------------------------------------------------
BILINEAR = {   # site kernel -> 3x3 taps (row-major) of (wR, wG, wB), x4
    0: [(0, 0, 1), (0, 1, 0), (0, 0, 1), (0, 1, 0), (4, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 0), (0, 0, 1)],
    1: [(0, 0, 0), (2, 0, 0), (0, 0, 0), (0, 0, 2), (0, 4, 0), (0, 0, 2), (0, 0, 0), (2, 0, 0), (0, 0, 0)],
    2: [(0, 0, 0), (0, 0, 2), (0, 0, 0), (2, 0, 0), (0, 4, 0), (2, 0, 0), (0, 0, 0), (0, 0, 2), (0, 0, 0)],
    3: [(1, 0, 0), (0, 1, 0), (1, 0, 0), (0, 1, 0), (0, 0, 4), (0, 1, 0), (1, 0, 0), (0, 1, 0), (1, 0, 0)],
}


def bayer_to_rgb_bilinear(bayer: np.ndarray, pattern: str = "RGGB", ccm=None, dtype: str | None = None) -> np.ndarray:
    dtype = dtype or dtype_name(bayer)
    h, w = bayer.shape
    in_scale = f32(SCALE[dtype_name(bayer)])
    pad = np.zeros((h + 2, w + 2), f32)
    pad[1:-1, 1:-1] = bayer.astype(f32)
    valid = np.zeros((h + 2, w + 2), bool)
    valid[1:-1, 1:-1] = True
    out = np.zeros((h, w, 3), f32)
    kidx = KERNEL_PATTERN[pattern]
    for slot in range(4):
        r0, c0 = slot & 1, slot >> 1
        c = np.zeros((h // 2, w // 2, 3), f32)
        t = np.zeros((h // 2, w // 2, 3), f32)
        for i, wt in enumerate(BILINEAR[kidx[slot]]):
            d0, d1 = i // 3 - 1, i % 3 - 1
            v = pad[1 + r0 + d0: 1 + r0 + d0 + h: 2, 1 + c0 + d1: 1 + c0 + d1 + w: 2]
            m = valid[1 + r0 + d0: 1 + r0 + d0 + h: 2, 1 + c0 + d1: 1 + c0 + d1 + w: 2]
            wv = np.asarray(wt, f32)
            c = c + np.where(m[..., None], v[..., None] * wv, f32(0))
            t = t + np.where(m[..., None], wv, f32(0))
        with np.errstate(all="ignore"):
            c = np.where(t > 0, c / (in_scale * t), f32(0)).astype(f32)
        if ccm is not None:
            c = _ccm_apply(c, ccm)
        out[r0::2, c0::2] = np.clip(c, f32(0), f32(1))
    return cast_to(out * f32(SCALE[dtype]), dtype)

"""CPU-only: host logic, the C ABI surface and oracle properties (no GPU compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import isp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from taichi_image_b200 import _lib
    header = open(os.path.join(ROOT, "include", "b200isp.h")).read()
    declared = set(re.findall(r"\b(b200isp_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200isp_fused_params"}
    assert declared, "no declarations found"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/b200isp.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib.b200isp_version() == 100
    assert _lib.WORKSPACE_BYTES >= 4096


def test_fused_params_struct_matches_header():
    from taichi_image_b200 import _lib
    # 7 ints, 9 floats, 4 floats, int, float, 7 ints, 2 floats, 5 ints (36 x 4 bytes, no padding) + 4 pointers + 2 size_t
    assert ctypes.sizeof(_lib.FusedParams) == 4 * (7 + 9 + 4 + 1 + 1 + 7 + 2 + 5) + 6 * ctypes.sizeof(ctypes.c_void_p)
    header = open(os.path.join(ROOT, "include", "b200isp.h")).read()
    body = header[header.index("typedef struct {"):header.index("} b200isp_fused_params;")]
    names = re.findall(r"\b(?:int|float|void\*|size_t)\s+([a-z_0-9, \[\]]+);", body)
    flat = [n.strip().split("[")[0] for group in names for n in group.split(",")]
    assert flat == [f[0] for f in _lib.FusedParams._fields_], flat


def test_error_paths_do_not_need_a_gpu():
    from taichi_image_b200 import _lib
    lib = _lib.lib
    assert lib.b200isp_decode12(None, 3, None, 1, 0, 0, None) == -3       # odd size
    assert b"even" in lib.b200isp_last_error()
    assert lib.b200isp_bayer_to_rgb(None, 0, None, 0, 3, 4, 0, None, None) == -3
    assert lib.b200isp_bayer_to_rgb(None, 0, None, 0, 4, 4, 9, None, None) == -1
    assert lib.b200isp_transform(None, None, 0, 4, 4, 99, None) == -1
    assert lib.b200isp_decode12(None, 0, None, 1, 0, 0, None) == 0        # empty input is a no-op
    p = _lib.FusedParams()
    p.height, p.width = 10, 12                                            # width % 8 != 0
    arr = (ctypes.c_void_p * 1)()
    assert lib.b200isp_process_packed12(arr, arr, 1, p, None, ctypes.c_void_p(8), None) == -3
    assert lib.b200isp_process_packed12(arr, arr, 65, p, None, ctypes.c_void_p(8), None) == -6


def test_no_cpu_fallback():
    import torch
    from taichi_image_b200 import packed, bayer, types
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        packed.encode12(np.zeros(4, np.uint16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bayer.bayer_to_rgb(np.zeros((4, 4), np.uint8))
    with pytest.raises(Exception):
        packed.decode12_kernel(types.u16)(torch.zeros(3, dtype=torch.uint8), torch.zeros(2, dtype=torch.uint16))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "taichi_image_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_dtypes_and_types_tables():
    import torch
    from taichi_image_b200 import dtypes, types, u8, u16, f16, f32, uint8
    assert u8 is uint8 and dtypes.as_dtype(torch.float16) is f16 and dtypes.as_dtype(np.uint16) is u16
    assert dtypes.as_dtype("f32") is f32 and dtypes.as_dtype("float32") is f32 and dtypes.as_dtype(np.dtype("uint8")) is u8
    assert types.scale_factor[u8] == 255 and types.scale_factor[u16] == 65535 and types.scale_factor[dtypes.i16] == 32767
    assert types.ti_type(np.zeros(1, np.float16)) is f16 and types.ti_type(torch.zeros(1, dtype=torch.uint16)) is u16
    e = types.empty_like(np.zeros((2, 3), np.uint8), (4, 5), f32)
    assert e.shape == (4, 5) and e.dtype == np.float32
    with pytest.raises(KeyError):
        dtypes.as_dtype(torch.float64)
    with pytest.raises(ValueError):
        types.ti_type([1, 2, 3])


def test_shard_cameras_partition():
    from taichi_image_b200.distributed import shard_cameras
    for n, w in [(12, 8), (12, 2), (12, 4), (6, 1), (3, 8), (64, 8)]:
        parts = [list(shard_cameras(n, w, r)) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert max(len(shard_cameras(12, 8, r)) for r in range(8)) == 2      # 12 cameras on 8 GPUs cap at 6x


def test_integer_demosaic_equals_floor_division():
    """The float quantisation chain of bayer.py:150-155 equals clamp(floor(S / t), 0, scale) for every
    reachable integer tap sum S and every border normaliser t -- the identity the CUDA integer path uses."""
    for s in (255, 65535, 32767):
        for t in (10, 11, 12, 13, 14, 15, 16, 17, 18, 20):
            S = np.arange(-64, t * s + 65, dtype=np.int64)
            f = np.clip(S.astype(np.float32) / np.float32(np.float32(s) * np.float32(t)), np.float32(0), np.float32(1))
            out = np.trunc(f * np.float32(s)).astype(np.int64)
            assert np.array_equal(out, np.clip(S // t, 0, s)), (s, t)


def test_border_normaliser_never_zero_and_constant_images():
    for p in O.PATTERNS:
        for shape in [(2, 2), (2, 4), (4, 2), (4, 6), (6, 6), (8, 10)]:
            b = np.full(shape, 200, np.uint8)
            assert np.all(O.bayer_to_rgb(b, p) == 200), (p, shape)


def test_malvar_matches_published_coefficients():
    from scipy.ndimage import correlate
    r = np.random.default_rng(0)
    cfa = (r.integers(0, 4096, size=(16, 20)) / 4095).astype(np.float32)
    rgb = O.demosaic_unit(cfa, "RGGB")
    g_at_rb = np.array([[0, 0, -1, 0, 0], [0, 0, 2, 0, 0], [-1, 2, 4, 2, -1], [0, 0, 2, 0, 0], [0, 0, -1, 0, 0]]) / 8
    r_at_g_rrow = np.array([[0, 0, .5, 0, 0], [0, -1, 0, -1, 0], [-1, 4, 5, 4, -1], [0, -1, 0, -1, 0], [0, 0, .5, 0, 0]]) / 8
    r_at_b = np.array([[0, 0, -1.5, 0, 0], [0, 2, 0, 2, 0], [-1.5, 0, 6, 0, -1.5], [0, 2, 0, 2, 0], [0, 0, -1.5, 0, 0]]) / 8
    f = lambda k: np.clip(correlate(cfa.astype(np.float64), k, mode="constant"), 0, 1)
    assert np.abs(f(g_at_rb)[2:-2:2, 2:-2:2] - rgb[2:-2:2, 2:-2:2, 1]).max() < 1e-6
    assert np.abs(f(r_at_g_rrow)[2:-2:2, 3:-2:2] - rgb[2:-2:2, 3:-2:2, 0]).max() < 1e-6
    assert np.abs(f(r_at_g_rrow.T)[3:-2:2, 2:-2:2] - rgb[3:-2:2, 2:-2:2, 0]).max() < 1e-6
    assert np.abs(f(r_at_b)[3:-2:2, 3:-2:2] - rgb[3:-2:2, 3:-2:2, 0]).max() < 1e-6


def test_packed_roundtrip_and_ids_quirk():
    r = np.random.default_rng(1)
    for _ in range(20):                                   # reference test/packed.py:6-15
        x = r.integers(0, 4096, size=int(r.integers(1000)) * 2).astype(np.uint16)
        assert np.array_equal(O.decode12(O.encode12(x)), x)
    x = r.integers(0, 4096, size=64).astype(np.uint16)
    y = O.decode12(O.encode12(x, ids_format=True), ids_format=True)
    exp = x.copy()                                        # the reference's IDS encode swaps the low nibbles
    exp[0::2] = (x[0::2] & 0xFF0) | (x[1::2] & 0xF)
    exp[1::2] = (x[1::2] & 0xFF0) | (x[0::2] & 0xF)
    assert np.array_equal(y, exp)


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    """Every entry point validates its arguments before the first CUDA call and reports through the status code +
    b200isp_last_error (include/b200isp.h conventions): exercised on the CPU with arguments that must be refused."""
    import ctypes as C
    from taichi_image_b200 import _lib
    lib = _lib.lib
    E_ARG, E_DTYPE, E_SHAPE, E_FRAMES = -1, -2, -3, -6
    dummy = C.c_void_p(0x1000)                                  # never dereferenced: validation fails first
    f9 = (C.c_float * 9)()
    assert lib.b200isp_bayer_to_rgb(dummy, 0, dummy, 0, 5, 8, 0, None, None) == E_SHAPE             # odd height
    assert b"even size" in lib.b200isp_last_error()
    assert lib.b200isp_bayer_to_rgb(dummy, 0, dummy, 0, 4, 8, 7, None, None) == E_ARG               # unknown pattern
    assert lib.b200isp_bayer_to_rgb(dummy, 9, dummy, 0, 4, 8, 0, None, None) == E_DTYPE
    assert lib.b200isp_bayer_to_rgb(None, 0, None, 0, 0, 0, 0, None, None) == 0                      # empty image: no-op
    assert lib.b200isp_rgb_to_bayer(dummy, dummy, 0, 4, 4, 9, None) == E_ARG
    assert lib.b200isp_rgb_yuv420(dummy, 0, dummy, 0, 5, 4, f9, None) == E_SHAPE
    assert lib.b200isp_transform(dummy, dummy, 0, 4, 4, 42, None) == E_ARG
    assert lib.b200isp_metering_update(None, 1, 4, 8, 8, 8, 0.0, None, None, None) == E_ARG
    p = _lib.FusedParams()
    p.height, p.width, p.pattern, p.isp_dtype, p.out_dtype, p.tonemap, p.gamma = 8, 16, 0, 4, 0, 0, 1.0
    ptrs = (C.c_void_p * 1)(0x1000)
    assert lib.b200isp_process_packed12(ptrs, ptrs, 0, p, dummy, dummy, None) == E_FRAMES           # no frames
    assert lib.b200isp_process_packed12(ptrs, ptrs, 65, p, dummy, dummy, None) == E_FRAMES          # > B200ISP_MAX_FRAMES
    p.width = 12
    assert lib.b200isp_process_packed12(ptrs, ptrs, 1, p, dummy, dummy, None) == E_SHAPE            # width % 8 != 0
    p.width, p.isp_dtype = 16, 0
    assert lib.b200isp_process_packed12(ptrs, ptrs, 1, p, dummy, dummy, None) == E_DTYPE            # ISP dtype must be f16 / f32
    p.isp_dtype, p.gamma = 4, 0.0
    ptrs16 = (C.c_void_p * 1)(0x1000)
    assert lib.b200isp_process_packed12(ptrs16, ptrs16, 1, p, dummy, dummy, None) == E_ARG          # gamma must be positive
    assert lib.b200isp_mailbox_bytes(0) == 0 and lib.b200isp_mailbox_bytes(8) > 0
    assert lib.b200isp_mailbox_exchange(dummy, 3, ptrs, 1, 0, dummy, None) == E_ARG                  # kind must be 1 | 2
    assert lib.b200isp_sample_histogram(dummy, 10, 0, dummy, None) == E_ARG
    with pytest.raises(_lib.B200ISPError, match="bad shape"):
        _lib.check(E_SHAPE, "demo")



def test_transform_routing_of_the_fused_call(monkeypatch):
    """host logic of ISP._fused_flip (no GPU): which ISP transforms the fused call applies itself -- flips in the sweep's
    store on every path, transposing transforms (interpolate.py:36-56) in the normalise pass of the one-sweep Reinhard -> u8
    forms only; everything else returns 0 = the transform kernel runs behind the call"""
    import torch
    from taichi_image_b200 import camera_isp, bayer
    from taichi_image_b200.dtypes import u8, u16, f16
    from taichi_image_b200.interpolate import ImageTransform as T
    for var in ("B200ISP_FUSED_TRANSPOSE", "B200ISP_TURN_IN_PASS", "B200ISP_NO_FUSED_TRANSFORM", "B200ISP_CAM16_RECOMPUTE",
                "B200ISP_CAM32_ONE_SWEEP", "B200ISP_REINHARD_EXACT"):
        monkeypatch.delenv(var, raising=False)
    cpu = torch.device("cpu")
    codes = {T.flip_horiz: 1, T.flip_vert: 2, T.rotate_180: 3, T.transpose: 4, T.rotate_270: 5, T.rotate_90: 6, T.transverse: 7}
    for cam in (camera_isp.Camera16, camera_isp.Camera32):
        for t, code in codes.items():
            isp = cam(bayer.BayerPattern.RGGB, transform=t, device=cpu)
            assert isp._fused_flip(3000, "reinhard", u8, 0.9, 0.0) == code
            turning = code & 4
            assert isp._fused_flip(3000, "linear", u16, 1.0, 0.0) == (0 if turning else code)        # the sweep writes the output itself
            assert isp._fused_flip(3000, "reinhard", f16, 0.9, 0.0) == (0 if turning else code)
            assert isp._fused_flip(3004, "reinhard", u8, 0.9, 0.0) == (0 if turning else code)       # height % 8 != 0
            assert isp._fused_flip(3000, "reinhard", u8, 0.9, 0.0, ids_format=True) == (0 if turning else code)
            assert isp._fused_flip(3000, "reinhard", u8, 0.9, 0.0, yuv420=True) == 0
        assert cam(bayer.BayerPattern.RGGB, device=cpu)._fused_flip(3000, "reinhard", u8, 0.9, 0.0) == 0
        assert cam(bayer.BayerPattern.RGGB, transform=T.rotate_90, resize_width=1920, device=cpu)._fused_flip(3000, "reinhard", u8, 0.9, 0.0) == 0
    # Camera32: only where the one-sweep u16 map applies (color_adapt == 0, 0.3 <= gamma <= 1, not reinhard_exact)
    isp = camera_isp.Camera32(bayer.BayerPattern.RGGB, transform=T.rotate_90, device=cpu)
    assert isp._fused_flip(3000, "reinhard", u8, 1.2, 0.0) == 0 and isp._fused_flip(3000, "reinhard", u8, 0.9, 0.5) == 0
    assert camera_isp.Camera32(bayer.BayerPattern.RGGB, transform=T.rotate_90, reinhard_exact=True, device=cpu)._fused_flip(3000, "reinhard", u8, 0.9, 0.0) == 0
    assert camera_isp.Camera16(bayer.BayerPattern.RGGB, transform=T.rotate_90, device=cpu)._fused_flip(3000, "reinhard", u8, 1.2, 0.5) == 6
    monkeypatch.setenv("B200ISP_TURN_IN_PASS", "0")
    assert isp._fused_flip(3000, "reinhard", u8, 0.9, 0.0) == 0
    monkeypatch.setenv("B200ISP_FUSED_TRANSPOSE", "1")                                               # the sweep's own transposing store (opt-in)
    assert isp._fused_flip(3000, "linear", u16, 1.0, 0.0) == 6

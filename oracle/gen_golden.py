#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference sources (/root/reference) on the
pure-Python Taichi emulation in oracle/taichi_shim.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the reference tree does not exist on the GPU box):

    python oracle/gen_golden.py

Every .npz holds the inputs and the reference outputs of one module; tests/test_oracle_golden.py checks
the numpy oracle against them on CPU, tests/test_gpu_golden.py checks the CUDA path against them.
Inputs are tiny because the emulation runs the kernels as Python loops.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "taichi_shim"))     # fake `taichi`, stub `turtle`
sys.path.insert(0, REFERENCE)

import numpy as np          # noqa: E402
import torch                # noqa: E402
import taichi as ti         # noqa: E402  (the shim)
from taichi_image import packed, bayer, tonemap, interpolate, camera_isp, color   # noqa: E402  (the reference)

OUT = os.path.join(ROOT, "tests", "golden")
TI = {"u8": ti.u8, "u16": ti.u16, "i16": ti.i16, "f16": ti.f16, "f32": ti.f32}
NP = {"u8": np.uint8, "u16": np.uint16, "i16": np.int16, "f16": np.float16, "f32": np.float32}
PATTERNS = ("RGGB", "GRBG", "GBRG", "BGGR")


def plane(r, shape, name):
    if name == "u8":
        return r.integers(0, 256, size=shape, dtype=np.uint8)
    if name == "u16":
        return r.integers(0, 65536, size=shape).astype(np.uint16)
    if name == "i16":
        return r.integers(0, 32768, size=shape).astype(np.int16)
    return r.random(shape, dtype=np.float32).astype(NP[name])


def gen_packed():
    r = np.random.default_rng(100)
    d = {}
    x = r.integers(0, 4096, size=(5, 24)).astype(np.uint16)
    d["values12"] = x
    for ids in (0, 1):
        e = packed.encode12(x, ids_format=bool(ids))
        d[f"encode_ids{ids}"] = e
        d[f"decode_ids{ids}"] = packed.decode12(e, ids_format=bool(ids))
    enc = r.integers(0, 256, size=(4, 36), dtype=np.uint8)
    d["encoded_random"] = enc
    for name in ("u8", "u16", "i16", "f16", "f32"):
        for ids in (0, 1):
            for scaled in (0, 1):
                # numpy u16 output with the default torch-free path; all dtypes exist for numpy in the reference
                d[f"decode12_{name}_ids{ids}_scaled{scaled}"] = packed.decode12(enc, dtype=TI[name], scaled=bool(scaled), ids_format=bool(ids))
        v = plane(r, (3, 16), name)
        d[f"values_{name}"] = v
        for ids in (0, 1):
            d[f"encode12_scaled_{name}_ids{ids}"] = packed.encode12(v, scaled=True, ids_format=bool(ids))
    enc16 = r.integers(0, 256, size=(3, 20), dtype=np.uint8)
    d["encoded16"] = enc16
    for name in ("u16", "f16", "f32"):
        for scaled in (0, 1):
            k = packed.decode16_kernel(TI[name], scaled=bool(scaled))     # decode16() itself always raises
            out = np.empty(enc16.size // 2, NP[name])
            k(enc16.reshape(-1), out)
            d[f"decode16_{name}_scaled{scaled}"] = out.reshape(3, 10)
    np.savez_compressed(os.path.join(OUT, "packed.npz"), **d)


def gen_bayer():
    r = np.random.default_rng(101)
    d = {}
    ccm = (camera_isp.Camera32(bayer.BayerPattern.RGGB, correct_colors=True, device=torch.device("cpu")).color_correct_matrix)
    d["ccm"] = ccm
    for name in ("u8", "u16", "f16", "f32"):
        rgb = plane(r, (8, 12, 3), name)
        cfa = plane(r, (10, 12), name)
        d[f"rgb_{name}"], d[f"cfa_{name}"] = rgb, cfa
        for p in PATTERNS:
            pat = bayer.BayerPattern[p]
            d[f"mosaic_{name}_{p}"] = bayer.rgb_to_bayer(rgb, pat)
            d[f"demosaic_{name}_{p}"] = bayer.bayer_to_rgb(cfa, pat)
            d[f"demosaic_ccm_{name}_{p}"] = bayer.bayer_to_rgb(cfa, pat, correct_colors=ccm)
    small = plane(r, (2, 2), "u8")
    d["cfa_2x2"] = small
    d["demosaic_2x2"] = bayer.bayer_to_rgb(small)
    d["demosaic_u8_to_f32"] = bayer.bayer_to_rgb(d["cfa_u8"], bayer.BayerPattern.GBRG, dtype=ti.f32)
    d["demosaic_u16_to_u8"] = bayer.bayer_to_rgb(d["cfa_u16"], bayer.BayerPattern.GRBG, dtype=ti.u8)
    np.savez_compressed(os.path.join(OUT, "bayer.npz"), **d)


def gen_tonemap():
    r = np.random.default_rng(102)
    d = {}
    img = (0.05 + 0.9 * r.random((9, 11, 3), dtype=np.float32)).astype(np.float32)
    img[2, 3] = (0.0, 0.02, 0.01)
    d["img_f32"] = img
    d["img_u8"] = (img * 255).astype(np.uint8)
    d["img_f16"] = img.astype(np.float16)
    for src in ("f32", "u8", "f16"):
        for out in ("u8", "u16", "f16", "f32"):
            for gi, gamma in enumerate((1.0, 0.6)):
                d[f"linear_{src}_{out}_g{gi}"] = tonemap.tonemap_linear(d[f"img_{src}"], gamma, TI[out])
        for out in ("u8", "u16", "f32"):
            d[f"reinhard_{src}_{out}_default"] = tonemap.tonemap_reinhard(d[f"img_{src}"], dtype=TI[out])
            d[f"reinhard_{src}_{out}_params"] = tonemap.tonemap_reinhard(d[f"img_{src}"], 0.6, 3.0, 0.9, 0.2, TI[out])
    np.savez_compressed(os.path.join(OUT, "tonemap.npz"), **d)


def gen_interpolate():
    r = np.random.default_rng(103)
    d = {}
    for name in ("u8", "f16", "f32"):
        img = plane(r, (10, 14, 3), name)
        d[f"img_{name}"] = img
        for si, s in enumerate((0.8, 0.469, 1.5)):
            d[f"scale_{name}_{si}"] = interpolate.scale_bilinear(img, s)
        d[f"width_{name}"] = interpolate.resize_width(img, 9)
        for t in interpolate.ImageTransform:
            if t == interpolate.ImageTransform.transverse:
                continue       # out of bounds for non-square images in the reference
            d[f"transform_{name}_{t.value}"] = interpolate.transform(img, t)
    sq = plane(r, (7, 7, 3), "u8")
    d["img_square"] = sq
    d["transform_square_transverse"] = interpolate.transform(sq, interpolate.ImageTransform.transverse)
    np.savez_compressed(os.path.join(OUT, "interpolate.npz"), **d)


def gen_color():
    """color/yuv_420.py: planar YUV 4:2:0 encode / decode"""
    r = np.random.default_rng(105)
    d = {}
    for name in ("u8", "u16", "f16", "f32"):
        # smooth-ish content keeps the chroma inside [0, 1]: the reference's "clamp" has no lower bound (yuv_420.py:57)
        img = plane(r, (6, 10, 3), name)
        d[f"rgb_{name}"] = img
        enc = color.rgb_yuv420_image(img)
        d[f"yuv_{name}"] = enc
        d[f"rgb_back_{name}"] = color.yuv420_rgb_image(enc)
    d["yuv_u8_to_f32"] = color.rgb_yuv420_image(d["rgb_u8"], ti.f32)
    d["yuv_f32_to_u8"] = color.rgb_yuv420_image(d["rgb_f32"], ti.u8)
    np.savez_compressed(os.path.join(OUT, "color.npz"), **d)


def smooth(r, h, w):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    base = 0.5 + 0.35 * np.sin(x / w * 5.1 + 0.3) * np.cos(y / h * 3.7)
    img = np.stack([base * 0.9, base, base * 0.7], -1) + 0.05 + 0.1 * (r.random((h, w, 3), dtype=np.float32) - 0.5)
    return np.clip(img, 0, 1).astype(np.float32)


def gen_camera_isp():
    """Two time steps x two cameras per configuration; RGGB only (the reference ISP ignores its
    bayer_pattern argument)."""
    r = np.random.default_rng(104)
    d = {}
    h, w = 16, 24
    frames = [[packed.encode12(bayer.rgb_to_bayer(smooth(r, h, w)), scaled=True) for _ in range(2)] for _ in range(2)]
    for s, fs in enumerate(frames):
        for c, f in enumerate(fs):
            d[f"frame_s{s}_c{c}"] = f
    configs = {
        "plain": dict(),
        "ccm": dict(correct_colors=True),
        "resize": dict(resize_width=16),
        "stride3": dict(metering_stride=3, moving_alpha=0.3),
    }
    tms = {"default": dict(), "script": dict(gamma=0.9, intensity=3.0, light_adapt=0.9, color_adapt=0.0),
           "coloradapt": dict(gamma=0.6, intensity=2.0, light_adapt=0.7, color_adapt=0.3)}
    for cam_name, cam in (("f16", camera_isp.Camera16), ("f32", camera_isp.Camera32)):
        for cfg_name, cfg in configs.items():
            for tm_name, tm in list(tms.items()) + [("linear", dict(gamma=0.8)), ("linear1", dict(gamma=1.0))]:
                isp = cam(bayer.BayerPattern.RGGB, device=torch.device("cpu"), **cfg)
                for s, fs in enumerate(frames):
                    images = [isp.load_packed12(torch.from_numpy(f.copy())) for f in fs]
                    key = f"{cam_name}_{cfg_name}_{tm_name}_s{s}"
                    if tm_name == "default":
                        for c, im in enumerate(images):
                            d[f"{cam_name}_{cfg_name}_rgb_s{s}_c{c}"] = im.numpy().copy()
                    if tm_name.startswith("linear"):
                        outs = isp.tonemap_linear(images, **tm)
                    else:
                        outs = isp.tonemap_reinhard(images, **tm)
                    d[key + "_metrics"] = isp.metrics.numpy().copy()
                    for c, o in enumerate(outs):
                        d[key + f"_c{c}"] = o.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "camera_isp.npz"), **d)


def gen_camera_isp_wide():
    """One wide case (20 x 776: four 256-pixel strips, the last one partial; with rows_per_task = 6 three row
    chunks): the interior-strip kind (K_CORE), the partial last strip and multi-chunk tasks of the CUDA sweep are
    pinned to the reference source too, not only to the oracle.  Two time steps x two cameras, script Reinhard
    settings -> u8 and linear gamma 1 -> u8, Camera16 and Camera32.  Takes ~10 minutes on the emulation."""
    r = np.random.default_rng(106)
    d = {}
    h, w = 20, 776
    frames = [[packed.encode12(bayer.rgb_to_bayer(smooth(r, h, w)), scaled=True) for _ in range(2)] for _ in range(2)]
    for s, fs in enumerate(frames):
        for c, f in enumerate(fs):
            d[f"frame_s{s}_c{c}"] = f
    tms = {"script": dict(gamma=0.9, intensity=3.0, light_adapt=0.9, color_adapt=0.0), "linear1": dict(gamma=1.0)}
    for cam_name, cam in (("f16", camera_isp.Camera16), ("f32", camera_isp.Camera32)):
        isps = {k: cam(bayer.BayerPattern.RGGB, device=torch.device("cpu")) for k in tms}
        for s, fs in enumerate(frames):
            images = [isps["script"].load_packed12(torch.from_numpy(f.copy())) for f in fs]
            for c, im in enumerate(images):
                d[f"{cam_name}_rgb_s{s}_c{c}"] = im.numpy().copy()
            for tm_name, tm in tms.items():
                ims = [im.clone() for im in images]           # tonemap_reinhard overwrites its inputs
                outs = isps[tm_name].tonemap_linear(ims, **tm) if tm_name.startswith("linear") else isps[tm_name].tonemap_reinhard(ims, **tm)
                key = f"{cam_name}_{tm_name}_s{s}"
                d[key + "_metrics"] = isps[tm_name].metrics.numpy().copy()
                for c, o in enumerate(outs):
                    d[key + f"_c{c}"] = o.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "camera_isp_wide.npz"), **d)


def main():
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for fn in (gen_packed, gen_bayer, gen_tonemap, gen_interpolate, gen_camera_isp, gen_color, gen_camera_isp_wide):
        if only and fn.__name__[4:] not in only:
            continue
        fn()
        print("wrote", fn.__name__[4:])
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("Golden vectors generated by `python oracle/gen_golden.py`: the unmodified reference sources\n"
                "(uc-vision/taichi_image 0.3.2, /root/reference) executed on the pure-Python Taichi emulation in\n"
                "`oracle/taichi_shim/`.  Keys are `<op>_<dtype>_<variant>`; inputs are stored next to the outputs.\n")


if __name__ == "__main__":
    main()

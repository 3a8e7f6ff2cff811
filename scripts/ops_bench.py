#!/usr/bin/env python
"""Device throughput of the stand-alone operators (the rows of SURVEY 8a outside the fused sweep) at BASELINE sizes:
algorithmic GB/s (each input byte read once, each output byte written once) against the measured HBM peak.
CUDA events around 20 repetitions replayed as one CUDA graph (device time, no host launch gaps) after 3 warm-ups,
inputs larger than L2 where the op allows it."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import taichi_image_b200 as tib                      # noqa: E402
from taichi_image_b200 import bayer, packed, interpolate, color, tonemap   # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timeit(fn, reps=20):
    """device time per call: the ``reps`` calls are captured into one CUDA graph (an eager Python call costs ~20 us
    on the host -- more than most of these kernels run); falls back to eager timing if the op cannot be captured"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a.record()
        g.replay()
        b.record()
    except Exception:
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, ms, nbytes, npx):
    gbs = nbytes / ms / 1e6
    note = "   [working set < 126 MB L2: L2-resident between repetitions, not an HBM figure]" if nbytes < 120e6 else ""
    print(f"{name:<58} {ms:8.3f} ms  {npx / ms / 1e6:8.1f} Gpx/s  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f} % of measured HBM peak{note}")


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    H8, W8 = 4320, 7680                      # "8K" (config 4)
    H, W = 3000, 4096                        # config 1 / 3 / 5 frame
    npx8, npx = H8 * W8, H * W
    for dt, name, s in ((torch.uint8, "u8", 1), (torch.uint16, "u16", 2), (torch.float32, "f32", 4)):
        if dt == torch.float32:
            cfa = torch.rand((H8, W8), device=dev)
        else:
            cfa = torch.randint(0, 256 if s == 1 else 65536, (H8, W8), device=dev, dtype=torch.int32).to(dt)
        out = torch.empty((H8, W8, 3), dtype=dt, device=dev)
        k = bayer.bayer_to_rgb_kernel(bayer.BayerPattern.RGGB, None, tib.types.ti_type(cfa), tib.types.ti_type(cfa))
        report(f"bayer_to_rgb {name} 7680x4320 (Malvar)", timeit(lambda: k(cfa, out)), npx8 * s * 4, npx8)
        report(f"rgb_to_bayer {name} 7680x4320", timeit(lambda: bayer.rgb_to_bayer(out)), npx8 * s * 4, npx8)
        report(f"bayer_to_rgb {name} 7680x4320 (bilinear, extension)",
               timeit(lambda: bayer.bayer_to_rgb(cfa, method="bilinear")), npx8 * s * 4, npx8)
    raw = torch.randint(0, 256, (H, W * 3 // 2), device=dev, dtype=torch.int32).to(torch.uint8)
    for dt, s in ((tib.u16, 2), (tib.f16, 2), (tib.f32, 4)):
        report(f"decode12 -> {dt.name} 4096x3000 (scaled)", timeit(lambda: packed.decode12(raw, dtype=dt, scaled=True)), npx * (1.5 + s), npx)
    vals = packed.decode12(raw, dtype=tib.u16)
    report("encode12 u16 4096x3000", timeit(lambda: packed.encode12(vals)), npx * 3.5, npx)
    rgb16 = torch.rand((H, W, 3), device=dev).half()
    rgb8 = (torch.rand((H, W, 3), device=dev) * 255).to(torch.uint8)
    ho, wo = round(H * 1920 / W), 1920
    report("resize_bilinear f16 4096x3000 -> 1920x1406", timeit(lambda: interpolate.resize_bilinear(rgb16, (wo, ho), 1920 / W)),
           (npx + ho * wo) * 6, npx)
    report("transform rotate_90 u8 4096x3000", timeit(lambda: interpolate.transform(rgb8, interpolate.ImageTransform.rotate_90)), npx * 6, npx)
    report("transform flip_horiz u8 4096x3000", timeit(lambda: interpolate.transform(rgb8, interpolate.ImageTransform.flip_horiz)), npx * 6, npx)
    report("rgb_yuv420_image u8 4096x3000", timeit(lambda: color.rgb_yuv420_image(rgb8)), npx * 4.5, npx)
    yuv = color.rgb_yuv420_image(rgb8)
    report("yuv420_rgb_image u8 4096x3000", timeit(lambda: color.yuv420_rgb_image(yuv)), npx * 4.5, npx)
    rgb32 = torch.rand((H, W, 3), device=dev)
    report("tonemap_linear f32 -> u8 4096x3000 (stand-alone)", timeit(lambda: tonemap.tonemap_linear(rgb32, 1.0, tib.u8)), npx * (12 * 2 + 3), npx)
    report("tonemap_reinhard f32 -> u8 4096x3000 (stand-alone, 4 reads + 1 write)", timeit(lambda: tonemap.tonemap_reinhard(rgb32, 0.9, 3.0, 0.9, 0.0, tib.u8)),
           npx * (12 + 3), npx)


if __name__ == "__main__":
    main()

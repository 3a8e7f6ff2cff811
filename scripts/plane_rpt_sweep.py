#!/usr/bin/env python
"""rows_per_task sweep of the stand-alone bayer_to_rgb sweep at 8K (B200ISP_PLANE_RPT knob of plane_sweep.cuh)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taichi_image_b200 as tib
from taichi_image_b200 import bayer
from scripts.ops_bench import timeit, PEAK

H8, W8 = 4320, 7680
dev = torch.device("cuda", 0)
for dt, name, s in ((torch.uint8, "u8", 1), (torch.uint16, "u16", 2), (torch.float32, "f32", 4)):
    cfa = torch.rand((H8, W8), device=dev) if dt == torch.float32 else torch.randint(0, 256 if s == 1 else 65536, (H8, W8), device=dev, dtype=torch.int32).to(dt)
    out = torch.empty((H8, W8, 3), dtype=dt, device=dev)
    k = bayer.bayer_to_rgb_kernel(bayer.BayerPattern.RGGB, None, tib.types.ti_type(cfa), tib.types.ti_type(cfa))
    res = []
    for rpt in (0, 12, 16, 20, 24, 28, 32, 40, 48):
        if rpt:
            os.environ["B200ISP_PLANE_RPT"] = str(rpt)
        else:
            os.environ.pop("B200ISP_PLANE_RPT", None)
        ms = timeit(lambda: k(cfa, out))
        res.append(f"{rpt}:{ms * 1e3:.1f}us/{100 * H8 * W8 * 4 * s / ms / 1e6 / PEAK:.0f}%")
    print(name, " ".join(res), flush=True)

#!/usr/bin/env python
"""a few stand-alone bayer_to_rgb launches at 8K for ncu: python scripts/plane_once.py u8|u16|f32"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taichi_image_b200 as tib
from taichi_image_b200 import bayer
name = sys.argv[1] if len(sys.argv) > 1 else "u8"
dt = {"u8": torch.uint8, "u16": torch.uint16, "f32": torch.float32}[name]
H8, W8 = 4320, 7680
cfa = torch.rand((H8, W8), device="cuda") if dt == torch.float32 else torch.randint(0, 256 if name == "u8" else 65536, (H8, W8), device="cuda", dtype=torch.int32).to(dt)
out = torch.empty((H8, W8, 3), dtype=dt, device="cuda")
k = bayer.bayer_to_rgb_kernel(bayer.BayerPattern.RGGB, None, tib.types.ti_type(cfa), tib.types.ti_type(cfa))
for _ in range(6):
    k(cfa, out)
torch.cuda.synchronize()

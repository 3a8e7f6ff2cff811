#!/usr/bin/env python
"""stand-alone tonemap_reinhard, a few calls (ncu launch-list target): in dtype from argv[1] (f32 / u8 / u16 / f16)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import taichi_image_b200 as tib
from taichi_image_b200 import tonemap
dt = sys.argv[1] if len(sys.argv) > 1 else "f32"
H, W = 3000, 4096
dev = torch.device("cuda", 0)
x = torch.rand((H, W, 3), device=dev)
if dt == "u8": x = (x * 255).to(torch.uint8)
elif dt == "u16": x = (x * 65535).to(torch.int32).to(torch.uint16)
elif dt == "f16": x = x.half()
for _ in range(4):
    y = tonemap.tonemap_reinhard(x, 0.9, 3.0, 0.9, 0.0, tib.u8)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    y = tonemap.tonemap_reinhard(x, 0.9, 3.0, 0.9, 0.0, tib.u8)
b.record(); torch.cuda.synchronize()
print(f"tonemap_reinhard {dt} -> u8 {W}x{H}: {a.elapsed_time(b) / 20:.4f} ms")

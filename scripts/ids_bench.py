#!/usr/bin/env python
"""cfg2-shaped frames: the fused path on the standard layout vs the IDS layout decoded inside the row loader (eager calls;
the same bytes are simply interpreted as IDS -- only the decode cost matters here)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import taichi_image_b200 as tib
from bench import synth_frames

n, h, w = 6, 3648, 5472
dev = torch.device("cuda", 0)
frames = synth_frames(n, h, w, 1234, dev)
outs = [torch.empty((h, w, 3), dtype=torch.uint16, device=dev) for _ in range(n)]
for ids in (False, True):
    isp = tib.camera_isp.Camera32(tib.bayer.BayerPattern.RGGB, device=dev)
    for _ in range(5):
        isp.process_packed12(frames, tonemap="linear", dtype="u16", out=outs, ids_format=ids)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30):
        isp.process_packed12(frames, tonemap="linear", dtype="u16", out=outs, ids_format=ids)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print(f"cfg2 frames, ids_format={ids}: {ms:.4f} ms/step  {n * h * w / ms / 1e6:.1f} Gpixel/s (eager, metering in the call)")

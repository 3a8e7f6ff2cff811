#!/usr/bin/env python
"""The ISP's transforms: applied by the sweep's store (default) vs the transform kernel behind the sweep (what r01 did,
forced here with B200ISP_NO_FUSED_TRANSFORM=1).  Eager calls with the metering in the call, cfg1-shaped (RGB8 Reinhard, 6 x
4096x3000) and cfg2-shaped (RGB16 linear, 6 x 5472x3648) frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import taichi_image_b200 as tib
from taichi_image_b200.interpolate import ImageTransform
from bench import synth_frames

dev = torch.device("cuda", 0)
for name, (n, h, w), tm, dt in (("cfg1", (6, 3000, 4096), "reinhard", "u8"), ("cfg2", (6, 3648, 5472), "linear", "u16")):
    frames = synth_frames(n, h, w, 1234, dev)
    for tname in ("none", "rotate_180", "rotate_90", "transpose"):
        for fused in ((True,) if tname == "none" else (True, False)):
            os.environ["B200ISP_NO_FUSED_TRANSFORM"] = "0" if fused else "1"
            isp = tib.camera_isp.Camera32(tib.bayer.BayerPattern.RGGB, device=dev, transform=ImageTransform[tname])
            for _ in range(5):
                isp.process_packed12(frames, tonemap=tm, dtype=dt, gamma=0.9 if tm == "reinhard" else 1.0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(30):
                isp.process_packed12(frames, tonemap=tm, dtype=dt, gamma=0.9 if tm == "reinhard" else 1.0)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 30
            print(f"{name} {tm}->{dt} {tname:10s} {'in the store' if fused else 'kernel behind':13s}: {ms:.4f} ms/step  {n * h * w / ms / 1e6:.1f} Gpixel/s", flush=True)

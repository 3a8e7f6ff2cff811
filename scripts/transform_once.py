#!/usr/bin/env python
"""A few eager calls of one (workload, transform) pair -- the target of an ncu capture (scripts/gpu_ncu_transform.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import taichi_image_b200 as tib
from taichi_image_b200.interpolate import ImageTransform
from bench import synth_frames
name, tname = sys.argv[1], sys.argv[2]
n, h, w, tm, dt = {"cfg1": (6, 3000, 4096, "reinhard", "u8"), "cfg2": (6, 3648, 5472, "linear", "u16")}[name]
dev = torch.device("cuda", 0)
frames = synth_frames(n, h, w, 1234, dev)
isp = tib.camera_isp.Camera32(tib.bayer.BayerPattern.RGGB, device=dev, transform=ImageTransform[tname])
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 4):
    isp.process_packed12(frames, tonemap=tm, dtype=dt, gamma=0.9 if tm == "reinhard" else 1.0)
torch.cuda.synchronize()

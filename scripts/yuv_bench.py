#!/usr/bin/env python
"""Camera16 + Reinhard + u8 on the reference's bench shape (6 x 4096x3000): RGB8 output, RGB8 followed by
rgb_yuv420_image, and YUV 4:2:0 written directly by the normalise pass (process_packed12(yuv420=True))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import taichi_image_b200 as tib
from bench import synth_frames

n, h, w = 6, 3000, 4096
dev = torch.device("cuda", 0)
frames = [torch.from_numpy(f).to(dev) for f in synth_frames(n, h, w)]


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


kw = dict(tonemap="reinhard", gamma=0.6)
isp = tib.camera_isp.Camera16(tib.bayer.BayerPattern.RGGB, device=dev)
rgb_out = [torch.empty((h, w, 3), dtype=torch.uint8, device=dev) for _ in range(n)]
yuv_out = [torch.empty((h * 3 // 2, w), dtype=torch.uint8, device=dev) for _ in range(n)]
t_rgb = timed(lambda: isp.process_packed12(frames, out=rgb_out, **kw))
t_two = timed(lambda: [tib.color.rgb_yuv420_image(o) for o in isp.process_packed12(frames, out=rgb_out, **kw)])
t_yuv = timed(lambda: isp.process_packed12(frames, out=yuv_out, yuv420=True, **kw))
px = n * h * w
for name, t in (("RGB8", t_rgb), ("RGB8 + rgb_yuv420_image", t_two), ("YUV 4:2:0 fused", t_yuv)):
    print(f"Camera16 Reinhard 6 x 4096x3000 -> {name:<26} {t:.4f} ms/step  {px / t / 1e6:7.1f} Gpixel/s")

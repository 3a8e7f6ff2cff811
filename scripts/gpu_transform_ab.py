#!/usr/bin/env python
"""Stand-alone interpolate.transform on the GPU: 6 distinct 4096x3000 images per repetition (442 / 885 MB of traffic for u8 /
u16: larger than L2), CUDA events around 20 repetitions.  B200ISP_TRANSFORM_TURN=0 selects the byte-granule tile kernel
(transform_words_kernel) for the transposing transforms, the default is the whole-word pixel tile (transform_turn_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_image_b200 import interpolate                      # noqa: E402
from taichi_image_b200.interpolate import ImageTransform       # noqa: E402

peak = 6551.4
for dt, esz in ((torch.uint8, 1), (torch.uint16, 2)):
    imgs = [torch.randint(0, 255, (3000, 4096, 3), dtype=torch.uint8, device="cuda").to(dt) for _ in range(6)]
    for tname in ("rotate_90", "rotate_270", "transpose", "transverse", "flip_horiz"):
        t = ImageTransform[tname]
        for _ in range(3):
            outs = [interpolate.transform(i, t) for i in imgs]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(20):
            outs = [interpolate.transform(i, t) for i in imgs]
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20 / 6
        gbs = 2 * 3000 * 4096 * 3 * esz / ms / 1e6
        print(f"turn={os.environ.get('B200ISP_TRANSFORM_TURN', '1')} {str(dt):12s} {tname:10s} {ms * 1e3:7.1f} us per image  {gbs:7.0f} GB/s  {100 * gbs / peak:5.1f} % of the measured HBM peak", flush=True)

#!/usr/bin/env python
"""A/B on the GPU: Reinhard -> RGB8 with the rig script's rotate_90 (scripts/tonemap_scan.py default), 6 x 4096x3000:
the transform kernel behind the call (B200ISP_TURN_IN_PASS=0) against the normalise pass turning the image itself
(csrc/fused_isp.cuh reinhard_out_transposed_kernel), and the untransformed call for scale.  Eager calls, metering off
(metrics from one first call), CUDA events around 40 calls after 5 warm-up calls.  Prints one line per variant."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taichi_image_b200 as tib                                   # noqa: E402
from taichi_image_b200.interpolate import ImageTransform          # noqa: E402
from bench import synth_frames                                    # noqa: E402


def run(dt, tname, in_pass, tm, n=6, h=3000, w=4096, iters=40):
    os.environ["B200ISP_TURN_IN_PASS"] = "1" if in_pass else "0"
    dev = torch.device("cuda", 0)
    cam = tib.camera_isp.Camera16 if dt == "f16" else tib.camera_isp.Camera32
    isp = cam(tib.bayer.BayerPattern.RGGB, device=dev, transform=ImageTransform[tname])
    frames = [synth_frames(1, h, w, 1234 + 17 * c, dev)[0] for c in range(n)]
    shape = (w, h, 3) if tname in ("rotate_90", "rotate_270", "transpose", "transverse") else (h, w, 3)
    outs = [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(n)]
    isp.process_packed12(frames, tonemap="reinhard", out=outs, **tm)
    for _ in range(5):
        isp.process_packed12(frames, tonemap="reinhard", out=outs, update_metering=False, **tm)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        isp.process_packed12(frames, tonemap="reinhard", out=outs, update_metering=False, **tm)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    chk = int(sum(int(o[::97, ::89].sum()) for o in outs))
    print(f"{dt} {tname:10s} in_pass={int(in_pass)} {ms:.4f} ms per call  {n * h * w / ms / 1e6:.1f} Gpixel/s  checksum {chk}", flush=True)
    return ms


if __name__ == "__main__":
    t0 = time.time()
    for dt, tm in (("f32", dict(gamma=0.9, intensity=3.0, light_adapt=0.9)), ("f16", dict(gamma=0.6))):
        run(dt, "none", True, tm)
        run(dt, "rotate_90", False, tm)
        run(dt, "rotate_90", True, tm)
        run(dt, "transpose", True, tm)
    print(f"done in {time.time() - t0:.1f} s")

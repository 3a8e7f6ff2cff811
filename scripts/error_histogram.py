#!/usr/bin/env python
"""Measured error histogram |CUDA - oracle| per output path ("measure and commit the error histogram per path").
Runs on the GPU box; prints one line per path: max |diff| in LSB and the fraction of values in each |diff| bucket.
The tolerances asserted in tests/ are the maxima printed here (profiles/r02_error_histogram.txt)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import isp_oracle as O                     # noqa: E402
from tests.util import rng, packed_frame, to_cuda, to_np, smooth_rgb   # noqa: E402
from tests.test_gpu_camera_isp import make_isp         # noqa: E402

EDGES = [0, 1, 2, 4, 8, 16, 32, 64, 1 << 30]


def hist(got, exp):
    d = np.concatenate([np.abs(g.astype(np.int64) - e.astype(np.int64)).ravel() for g, e in zip(got, exp)])
    h = [float(np.count_nonzero((d >= a) & (d < b))) / d.size for a, b in zip(EDGES[:-1], EDGES[1:])]
    return int(d.max()), h


def line(name, got, exp):
    mx, h = hist(got, exp)
    print(f"{name:58s} max {mx:4d}  " + " ".join(f"[{a}{'' if b == a + 1 else '..' + str(b - 1) if b < 1 << 30 else '+'}]={v:.2e}" for a, b, v in zip(EDGES[:-1], EDGES[1:], h)),
          flush=True)


def main():
    h, w, n = 256, 776, 2
    for dt in ("f32", "f16"):
        for ccm in (False, True):
            r = rng(7)
            fr = [packed_frame(r, h, w) for _ in range(n)]
            cu = [to_cuda(f) for f in fr]
            for gamma in (1.0, 0.7, 2.2):
                for out in ("u8", "u16"):
                    isp, ref = make_isp(dt, correct_colors=ccm), O.ISP(dt, correct_colors=ccm)
                    got = [to_np(g) for g in isp.process_packed12(cu, tonemap="linear", gamma=gamma, dtype=out)]
                    exp = ref.tonemap_linear([ref.load_packed12(f) for f in fr], gamma=gamma, out_dtype=out)
                    line(f"fused linear   {dt} ccm={int(ccm)} gamma={gamma} -> {out}", got, exp)
            for tm in (dict(), dict(gamma=0.6), dict(gamma=0.9, intensity=3.0, light_adapt=0.9), dict(gamma=0.8, intensity=2.0, light_adapt=0.7, color_adapt=0.3),
                       dict(gamma=2.2)):
                for out in ("u8", "u16"):
                    isp, ref = make_isp(dt, correct_colors=ccm), O.ISP(dt, correct_colors=ccm)
                    got = [to_np(g) for g in isp.process_packed12(cu, tonemap="reinhard", dtype=out, **tm)]
                    exp = ref.tonemap_reinhard([ref.load_packed12(f) for f in fr], out_dtype=out, **tm)
                    line(f"fused reinhard {dt} ccm={int(ccm)} {tm} -> {out}", got, exp)
    # stand-alone tonemap module (tonemap.py:26-46, :134-168)
    from taichi_image_b200 import tonemap
    r = rng(8)
    img = smooth_rgb(r, 256, 384)
    for out in ("u8", "u16"):
        for kw in (dict(), dict(gamma=0.6, intensity=3.0, light_adapt=0.9, color_adapt=0.2)):
            got = to_np(tonemap.tonemap_reinhard(to_cuda(img), dtype=getattr(__import__("taichi_image_b200"), out), **kw))
            exp = O.tonemap_reinhard(img, dtype=out, **kw)
            line(f"standalone tonemap_reinhard f32 {kw} -> {out}", [got], [exp])
        for gamma in (1.0, 0.6):
            got = to_np(tonemap.tonemap_linear(to_cuda(img), gamma=gamma, dtype=getattr(__import__("taichi_image_b200"), out)))
            exp = O.tonemap_linear(img, gamma=gamma, dtype=out)
            line(f"standalone tonemap_linear f32 gamma={gamma} -> {out}", [got], [exp])


if __name__ == "__main__":
    assert torch.cuda.is_available()
    main()

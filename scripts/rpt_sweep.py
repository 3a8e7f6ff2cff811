#!/usr/bin/env python
"""rows_per_task sweep of the fused sweep (kernel alone, CUDA events): validates the wave model of make_geom2."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taichi_image_b200 as tib
from tests.test_gpu_fullsize import synth_packed

CASES = [("cfg1", 1, 3000, 4096, "f32", "reinhard", "u8", dict()),
         ("cfg3", 6, 3000, 4096, "f32", "reinhard", "u8", dict(gamma=0.9, intensity=3.0, light_adapt=0.9)),
         ("cfg2", 6, 3648, 5472, "f32", "linear", "u16", dict())]
RPTS = [0, 12, 16, 20, 22, 24, 26, 28, 32, 36, 42, 48, 54, 64]
for name, n, h, w, dt, tm, out, kw in CASES:
    cam = tib.camera_isp.Camera16 if dt == "f16" else tib.camera_isp.Camera32
    isp = cam(tib.bayer.BayerPattern.RGGB)
    dev, _ = synth_packed(n, h, w, seed=3)
    outs = isp.process_packed12(dev, tonemap=tm, dtype=out, **kw)
    res = []
    for rpt in RPTS:
        ts = []
        for rep in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            isp._run_fused(dev, tm, tib.as_dtype(out), outs, kw, update_metering=False, rows_per_task=rpt)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        res.append((rpt, ts[len(ts) // 2]))
    print(name, " ".join(f"{r}:{t:.1f}us" for r, t in res), flush=True)

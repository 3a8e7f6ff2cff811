import sys, torch, numpy as np
sys.path.insert(0, ".")
from tests.test_gpu_camera_isp import frames, make_isp
from tests.util import rng, to_cuda
from taichi_image_b200.interpolate import ImageTransform, transform
r = rng(99)
for (h, w) in [(48, 1032), (64, 264)]:
    cu = [to_cuda(f) for f in frames(r, 2, h, w)]
    for tname in ["rotate_90", "transpose"]:
        t = ImageTransform[tname]
        for rpt in (0, 8, 16):
            plain, turned = make_isp("f32"), make_isp("f32", transform=t)
            kw = dict(gamma=0.9, intensity=2.0)
            exp = plain.process_packed12(cu, tonemap="reinhard", dtype="u8", rows_per_task=rpt, **kw)
            got = turned.process_packed12(cu, tonemap="reinhard", dtype="u8", rows_per_task=rpt, **kw)
            inv = {"rotate_90": ImageTransform.rotate_270, "transpose": ImageTransform.transpose}[tname]
            for g, e in zip(got, exp):
                gb = transform(g, inv)
                d = (gb.int() - e.int()).abs()
                idx = torch.nonzero(d.sum(-1))
                print(h, w, tname, rpt, "ndiff", len(idx), "max", int(d.max()), idx[:12].tolist())

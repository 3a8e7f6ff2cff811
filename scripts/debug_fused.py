import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import isp_oracle as O
from tests.util import rng, packed_frame, to_cuda, to_np
from taichi_image_b200 import camera_isp, bayer

for dt, cam in (("f16", camera_isp.Camera16), ("f32", camera_isp.Camera32)):
    for ccm in (False, True):
        r = rng(36)
        isp, ref = cam(bayer.BayerPattern.RGGB, correct_colors=ccm), O.ISP(dt, correct_colors=ccm)
        for step in range(2):
            fr = [packed_frame(r, 36, 72) for _ in range(2)]
            got = isp.process_packed12([to_cuda(f) for f in fr], tonemap="linear", gamma=1.0, dtype="u16")
            ims = [ref.load_packed12(f) for f in fr]
            exp = ref.tonemap_linear(ims, gamma=1.0, out_dtype="u16")
            print(dt, "ccm", ccm, "step", step, "metrics gpu", to_np(isp.metrics)[:2], "ref", ref.metrics[:2],
                  "eq", np.array_equal(to_np(isp.metrics)[:2], ref.metrics[:2]))
            # eager rgb for comparison
            rgb_gpu = to_np(isp.load_packed12(to_cuda(fr[0])))
            print("   rgb eager-vs-oracle mismatches", np.count_nonzero(rgb_gpu != ims[0]), "of", ims[0].size,
                  "max", np.abs(rgb_gpu.astype(np.float64) - ims[0].astype(np.float64)).max())
            for g, e in zip(got, exp):
                d = np.abs(to_np(g).astype(np.int64) - e.astype(np.int64))
                print("   out diff: max", d.max(), "frac", np.count_nonzero(d) / d.size, "hist", np.bincount(np.minimum(d.ravel(), 40))[:8],
                      "sat frac", np.mean(e == 65535))

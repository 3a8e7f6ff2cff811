"""CPU: the C/OpenMP restatement (bench.py's CPU baseline) against the numpy oracle."""
import numpy as np
import pytest

from oracle import isp_oracle as O
from oracle import c_oracle
from tests.util import rng, packed_frame


@pytest.mark.parametrize("cam", ["f16", "f32"])
@pytest.mark.parametrize("pattern", O.PATTERNS)
@pytest.mark.parametrize("mode", ["reinhard", "linear"])
def test_c_oracle_matches_numpy_oracle(cam, pattern, mode):
    r = rng(50)
    ccm = (O.DEFAULT_CC * O.DEFAULT_WB) if pattern == "GRBG" else None
    ref = O.ISP(cam, pattern, correct_colors=ccm is not None, metering_stride=4)
    metrics = None
    for step in range(2):
        frames = [packed_frame(r, 24, 40, pattern) for _ in range(2)]
        ims = [ref.load_packed12(f) for f in frames]
        out_dtype = "u16" if mode == "linear" else "u8"
        if mode == "reinhard":
            exp = ref.tonemap_reinhard(ims, gamma=0.8, intensity=2.0, light_adapt=0.9, color_adapt=0.1)
        else:
            exp = ref.tonemap_linear(ims, gamma=1.0, out_dtype="u16")
        got, metrics = c_oracle.process(frames, pattern, cam == "f16", out_dtype, mode, ccm, 0.8 if mode == "reinhard" else 1.0,
                                        2.0, 0.9, 0.1, stride=4, alpha=0.0 if step == 0 else 0.9, metrics=metrics, nthreads=2)
        assert np.allclose(metrics, ref.metrics, rtol=2e-6, atol=2e-6)
        for g, e in zip(got, exp):
            assert np.abs(g.astype(np.int64) - e.astype(np.int64)).max() <= 1

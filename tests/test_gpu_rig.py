"""GPU: rig ingestion / output stage (taichi_image_b200/rig.py; reference scripts/tonemap_scan.py:63-100, :151-179):
raw files -> pinned buffers -> fused ISP writing the camera grid in place == reference-shaped calls + concat_image_grid."""
import numpy as np
import pytest
import torch

from tests.test_gpu_camera_isp import frames, make_isp
from tests.util import rng

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt,tm", [("f32", "reinhard"), ("f32", "linear"), ("f16", "reinhard")])
def test_reader_and_grid_output(cuda, tmp_path, dt, tm):
    from taichi_image_b200 import rig
    r = rng(95)
    n_cam, steps, h, w = 4, 4, 48, 64
    data = [[frames(r, 1, h, w)[0] for _ in range(n_cam)] for _ in range(steps)]
    for c in range(n_cam):
        (tmp_path / f"cam{c}").mkdir()
        for s in range(steps):
            data[s][c].tofile(tmp_path / f"cam{c}" / f"frame_{s:03d}.raw")
    folders, names = rig.find_folder_images(tmp_path)
    assert len(folders) == n_cam and names == [f"frame_{s:03d}.raw" for s in range(steps)]
    reader = rig.RawFrameReader(folders, names, h, w, depth=2)
    direct, gridded = make_isp(dt, moving_alpha=0.2), make_isp(dt, moving_alpha=0.2)
    grid = rig.GridOutput(n_cam, 2, h, w)
    assert tuple(grid.image.shape) == (2 * h, 2 * w, 3) and not grid.tiles[1].is_contiguous()
    for s, (name, host) in enumerate(reader):
        assert name == names[s] and all(f.is_pinned() for f in host)
        for c in range(n_cam):
            assert np.array_equal(host[c].numpy(), data[s][c])
        dev = [f.to("cuda", non_blocking=True) for f in host]
        exp = rig.concat_image_grid(direct.process_packed12(dev, tonemap=tm, gamma=0.9), rows=2)
        out = gridded.process_packed12(dev, tonemap=tm, gamma=0.9, out=grid.tiles)
        assert out[0].data_ptr() == grid.tiles[0].data_ptr()
        torch.cuda.synchronize()
        assert torch.equal(grid.image, exp), f"step {s}: grid written in place differs from concat_image_grid"
        assert torch.equal(gridded.metrics, direct.metrics)

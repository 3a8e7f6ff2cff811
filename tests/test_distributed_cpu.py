"""Host-side logic of the multi-GPU path on CPU: camera sharding and the shared-exposure exchange protocol
(taichi_image_b200/distributed.py) with world_size-2 gloo process groups.  The CUDA kernels behind the three
local steps are replaced by the oracle's restatement of the same steps (oracle.metering_phase1/2/finalize);
the kernels themselves are checked against those in tests/test_gpu_distributed.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import isp_oracle as O
from taichi_image_b200.distributed import (SharedExposure, exchange, max_cameras_per_rank, shard_cameras,
                                           shared_metering)


def test_shard_cameras_partition():
    for n in (1, 6, 12, 16, 64):
        for world in (1, 2, 4, 8):
            parts = [list(shard_cameras(n, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(n))                       # contiguous, disjoint, complete
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == max_cameras_per_rank(n, world)
    assert [len(shard_cameras(12, 8, r)) for r in range(8)] == [2, 2, 2, 2, 1, 1, 1, 1]


class OracleBackend:
    """distributed.CudaMeteringBackend with the oracle's arithmetic (CPU tensors)"""

    def __init__(self, moving_alpha=0.1, stride=8):
        self.metrics, self.moving_alpha, self.stride = None, moving_alpha, stride

    def begin(self):
        if self.metrics is None:
            self.metrics = torch.zeros(9)
            return 0.0
        return 1.0 - self.moving_alpha

    def phase1(self, source):
        return torch.from_numpy(O.metering_phase1(source, self.stride))

    def phase2(self, source, g1, alpha):
        return torch.from_numpy(O.metering_phase2(source, g1.numpy(), alpha, self.metrics.numpy(), self.stride))

    def finalize(self, g1, g2, alpha):
        self.metrics.copy_(torch.from_numpy(O.metering_finalize(g1.numpy(), g2.numpy(), alpha, self.metrics.numpy())))


def rank_images(rank, step, n=2, h=40, w=56):
    r = np.random.default_rng(100 * step + rank)
    return [np.clip(r.random((h, w, 3), dtype=np.float32) * (0.5 + 0.3 * rank) + 0.05 * step, 0, 1) for _ in range(n + rank)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = exchange(torch.tensor([float(rank), 10.0 + rank]))
        assert g.tolist() == [[float(r), 10.0 + r] for r in range(world)]
        backend = OracleBackend()
        history = []
        for step in range(3):
            shared_metering(backend, rank_images(rank, step))
            history.append(backend.metrics.clone())
        # all ranks must hold bit-identical metrics without a broadcast
        mine = torch.stack(history)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        for other in gathered:
            assert torch.equal(other, mine)
        if rank == 0:
            out.put(mine.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_shared_exposure_gloo_world2_matches_joint_metering():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, PORT, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0, "a rank failed (see its traceback above)"
    got = out.get(timeout=10)
    # reference semantics: ONE metering call over the images of all cameras (camera_isp.py:168-175)
    metrics = np.zeros(9, np.float32)
    for step in range(3):
        images = sum((rank_images(r, step) for r in range(world)), [])
        metrics = O.metering_update(images, metrics, 0.0 if step == 0 else 0.9, 8)
        np.testing.assert_allclose(got[step], metrics, rtol=2e-6, atol=1e-7)


PORT = _free_port()


def test_shared_exposure_wrapper_forwards_and_skips_local_metering():
    """single process (no process group): SharedExposure must hand the ISP a metering function that runs the
    exchange protocol through the backend exactly once per call (fused path), and tone-map with
    update_metering=False on the eager path"""
    calls = []

    class FakeISP:
        device = torch.device("cpu")
        metrics = None
        answer = 42

        def _fused_ok(self, f, ids):
            return True
        _resizes = False

        def process_packed12(self, frames, **kw):        # like ISP.process_packed12: meters through meter_fn, then sweeps
            kw["meter_fn"](frames, None, None, True)      # alpha None: the backend decides (first call -> 0)
            calls.append(("process", kw.get("tonemap")))
            return ["out"]

        def tonemap_linear(self, images, gamma, **kw):
            calls.append(("linear", kw.get("update_metering")))
            return images

    class Backend(OracleBackend):
        def phase1(self, source):
            calls.append(("phase1", len(source)))
            return torch.tensor([0.0, 1.0])

        def phase2(self, source, g1, alpha):
            calls.append(("phase2", tuple(g1.shape), alpha))
            return torch.tensor([-1.0, 0.0, -2.0, 1.0, 1.0, 1.0, 1.0, 4.0])

    isp = SharedExposure(FakeISP(), backend=Backend())
    assert isp.answer == 42                                    # attribute forwarding
    assert isp.process_packed12([torch.zeros(4, 12, dtype=torch.uint8)], tonemap="linear") == ["out"]
    assert calls == [("phase1", 1), ("phase2", (1, 2), 0.0), ("process", "linear")]
    calls.clear()
    isp.tonemap_linear([torch.zeros(8, 8, 3)], 1.0)
    assert [c[0] for c in calls] == ["phase1", "phase2", "linear"] and calls[-1][1] is False
    assert calls[1][2] == pytest.approx(0.9)                   # second call: EMA weight of the previous metrics

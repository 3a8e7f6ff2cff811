"""GPU parity of the shared-exposure kernels (include/b200isp.h "multi-GPU shared exposure") on ONE GPU: the
ranks of a 2- / 3-rank rig are emulated as successive calls on disjoint camera shards whose records are
concatenated the way the NCCL all-gather would; the result must equal the joint single-call metering (what the
reference computes, camera_isp.py:168-175) and the oracle's split restatement."""
import numpy as np
import pytest
import torch

from oracle import isp_oracle as O
from taichi_image_b200.distributed import SharedExposure, shard_cameras
from tests.test_gpu_camera_isp import frames, make_isp
from tests.util import rng, to_cuda, to_np, assert_close_int

pytestmark = pytest.mark.gpu


class EmulatedRanks:
    """distributed.CudaMeteringBackend for `world` camera shards living on one GPU (one ISP per rank)"""

    def __init__(self, isps):
        self.isps = isps

    def step(self, shards):
        alphas = [isp._metrics_and_alpha() for isp in self.isps]
        g1 = torch.stack([isp.meter_phase1(s) for isp, s in zip(self.isps, shards)]).contiguous()
        g2 = torch.stack([isp.meter_phase2(s, g1, a) for isp, s, a in zip(self.isps, shards, alphas)]).contiguous()
        for isp, a in zip(self.isps, alphas):
            isp.meter_finalize(g1, g2, a)


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("source", ["packed", "images"])
def test_split_metering_equals_joint(cuda, dt, world, source):
    r = rng(70 + world)
    n_cam = 5
    ranks = EmulatedRanks([make_isp(dt, moving_alpha=0.1) for _ in range(world)])
    joint, ref = make_isp(dt, moving_alpha=0.1), O.ISP(dt, moving_alpha=0.1)
    for step in range(3):
        fr = frames(r, n_cam, 40, 56)
        cu = [to_cuda(f) for f in fr]
        ims_ref = [ref.load_packed12(f) for f in fr]
        ref.update_metering(ims_ref)
        if source == "packed":
            src = cu
            joint.process_packed12(cu, tonemap="linear")
        else:
            src = [joint.load_packed12(f) for f in cu]
            joint.update_metering(src)
        ranks.step([[src[i] for i in shard_cameras(n_cam, world, k)] for k in range(world)])
        m_joint = to_np(joint.metrics)
        for isp in ranks.isps:
            m = to_np(isp.metrics)
            np.testing.assert_array_equal(m, to_np(ranks.isps[0].metrics))       # bit-identical on every rank
            np.testing.assert_allclose(m, m_joint, rtol=5e-6, atol=1e-6)         # == the joint call (sum order differs)
            np.testing.assert_allclose(m, ref.metrics, rtol=2e-3 if dt == "f16" else 2e-5, atol=1e-5)
    # oracle's split restatement on the last step
    g1 = np.stack([O.metering_phase1([ims_ref[i] for i in shard_cameras(n_cam, world, k)]) for k in range(world)])
    assert g1.shape == (world, 2)


def test_shared_exposure_world1_equals_plain(cuda):
    """no process group: SharedExposure(isp) must reproduce isp.process_packed12 bit for bit"""
    r = rng(77)
    a, b = make_isp("f32"), SharedExposure(make_isp("f32"))
    for step in range(2):
        cu = [to_cuda(f) for f in frames(r, 3, 32, 64)]
        ya = a.process_packed12(cu, tonemap="reinhard", gamma=0.9, intensity=2.0)
        yb = b.process_packed12(cu, tonemap="reinhard", gamma=0.9, intensity=2.0)
        np.testing.assert_allclose(to_np(a.metrics), to_np(b.metrics), rtol=1e-6, atol=1e-7)
        for x, y in zip(ya, yb):
            assert_close_int(to_np(x), to_np(y), 1, "shared vs plain")


def test_peer_exchange_world1_and_emulated_ranks(cuda):
    """csrc/exchange.cu on one GPU: (a) PeerExchange without a process group is a local copy, repeatable (sequence
    numbers / parity); (b) two mailboxes owned by this process emulate two ranks: both post, then both wait"""
    import ctypes as C
    from taichi_image_b200 import _lib
    from taichi_image_b200.distributed import PeerExchange
    px = PeerExchange(torch.device("cuda", 0))
    for step in range(5):
        r1 = torch.tensor([0.1 * step, 1.0 + step], device="cuda")
        r2 = torch.arange(8, dtype=torch.float32, device="cuda") + step
        assert torch.equal(px(r1, 1), r1.reshape(1, 2)) and torch.equal(px(r2, 2), r2.reshape(1, 8))
    px.check()
    # (b) two emulated ranks
    world, boxes = 2, []
    for _ in range(world):
        p, h = C.c_void_p(), (C.c_ubyte * 64)()
        _lib.check(_lib.lib.b200isp_mailbox_create(world, C.byref(p), h), "create")
        boxes.append(p)
    peers = (C.c_void_p * world)(*[b.value for b in boxes])
    st = _lib.stream_ptr(torch.device("cuda", 0))
    gathered = [torch.zeros((world, 8), device="cuda") for _ in range(world)]
    for step in range(4):
        recs = [torch.arange(8, dtype=torch.float32, device="cuda") * (r + 1) + 10 * step for r in range(world)]
        for r in range(world):
            _lib.check(_lib.lib.b200isp_mailbox_post(recs[r].data_ptr(), 2, peers, world, r, st), "post")
        for r in range(world):
            _lib.check(_lib.lib.b200isp_mailbox_wait(boxes[r], 2, world, gathered[r].data_ptr(), st), "wait")
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(gathered[r], torch.stack(recs)), f"step {step} rank {r}"
            assert _lib.lib.b200isp_mailbox_error(boxes[r], world, st) == 0
    for b in boxes:
        _lib.check(_lib.lib.b200isp_mailbox_close(b, 1), "close")


class _FakePeer:
    """the part of distributed.PeerExchange that meter_packed12_shared uses: mailbox table, world, rank"""

    def __init__(self, peers, world, rank):
        self._peers, self.world, self.rank = peers, world, rank


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("world", [2, 3])
def test_in_kernel_exchange_emulated_ranks(cuda, dt, world):
    """b200isp_meter_packed12_shared (the record exchanges run INSIDE the two metering kernels): `world` ranks emulated on
    one GPU -- one mailbox, one ISP and one stream per rank, so that a rank's spinning last block really waits for the
    other ranks' kernels.  All ranks must end with bit-identical metrics, identical to the split phase1 / phase2 /
    finalize chain, and equal to the joint single-call metering of all cameras (reduction order aside)."""
    import ctypes as C
    from taichi_image_b200 import _lib
    r = rng(90 + world)
    n_cam = 5
    boxes = []
    for _ in range(world):
        p, h = C.c_void_p(), (C.c_ubyte * 64)()
        _lib.check(_lib.lib.b200isp_mailbox_create(world, C.byref(p), h), "create")
        boxes.append(p)
    table = (C.c_void_p * world)(*[b.value for b in boxes])
    peers = [_FakePeer(table, world, k) for k in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    isps = [make_isp(dt, moving_alpha=0.1) for _ in range(world)]
    chain = EmulatedRanks([make_isp(dt, moving_alpha=0.1) for _ in range(world)])
    joint = make_isp(dt, moving_alpha=0.1)
    for step in range(4):
        cu = [to_cuda(f) for f in frames(r, n_cam, 40, 56)]
        shards = [[cu[i] for i in shard_cameras(n_cam, world, k)] for k in range(world)]
        torch.cuda.synchronize()
        for k in range(world):
            alpha = isps[k]._metrics_and_alpha()
            with torch.cuda.stream(streams[k]):
                isps[k].meter_packed12_shared(shards[k], alpha, peers[k])
        torch.cuda.synchronize()
        chain.step(shards)
        joint.process_packed12(cu, tonemap="linear")
        for k in range(world):
            assert _lib.lib.b200isp_mailbox_error(boxes[k], world, _lib.stream_ptr(torch.device("cuda", 0))) == 0
            np.testing.assert_array_equal(to_np(isps[k].metrics), to_np(isps[0].metrics))
            np.testing.assert_array_equal(to_np(isps[k].metrics), to_np(chain.isps[0].metrics))
            np.testing.assert_allclose(to_np(isps[k].metrics), to_np(joint.metrics), rtol=5e-6, atol=1e-6)
    for b in boxes:
        _lib.check(_lib.lib.b200isp_mailbox_close(b, 1), "close")

"""Multi-GPU camera sharding: one process per GPU, one camera stream (or a contiguous
group of streams) per rank, no pixel traffic between GPUs.

The only coupling between cameras in the reference is the *joint* metering of all cameras of a time
step (camera_isp.py:168-175, :376-385): one min/max reduction, a moving-average blend of the bounds,
a second reduction w.r.t. the blended bounds, and the moving-average update of the 9-float metrics.
``SharedExposure`` reproduces that across ranks by splitting the two reductions at their exchange
points (include/b200isp.h, "multi-GPU shared exposure"):

    rec1 = phase1(local frames)                      {min, max}                        2 floats
    g1   = all_gather(rec1)                                                            NCCL, 8 B / rank
    rec2 = phase2(local frames, g1, alpha, metrics)  {lmin, lmax, 5 sums, n}            8 floats
    g2   = all_gather(rec2)                                                            NCCL, 32 B / rank
    finalize(g1, g2, alpha)                          metrics = lerp(alpha, joint stats, metrics)

Every rank folds the gathered records in rank order, so all ranks hold bit-identical ``metrics``
without a broadcast, and a rank with fewer cameras (12 cameras on 8 GPUs) simply contributes a smaller
``n``.  The messages are latency-only; the collectives are enqueued by torch.distributed behind the
compute stream and nothing synchronises with the host.  Without shared exposure the ranks are
independent replicas and no collective runs at all.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist


def shard_cameras(n_cameras: int, world_size: int, rank: int) -> range:
    """Contiguous balanced partition of camera indices: the first ``n % world`` ranks get one extra."""
    base, extra = divmod(n_cameras, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def max_cameras_per_rank(n_cameras: int, world_size: int) -> int:
    """Cameras on the most loaded rank = what bounds the step time (12 cameras on 8 GPUs -> 2)."""
    return -(-n_cameras // world_size)


class CudaMeteringBackend:
    """The three local steps of the exchange on the CUDA kernels of an ``ISP`` (camera_isp.py)."""

    def __init__(self, isp):
        self.isp = isp

    def begin(self) -> float:
        """allocates ``metrics`` on the first call; returns the weight of the previous metrics
        (0 on the first call, 1 - moving_alpha afterwards; camera_isp.py:376-385)"""
        return self.isp._metrics_and_alpha()

    def phase1(self, source) -> torch.Tensor:
        return self.isp.meter_phase1(source)

    def phase2(self, source, gathered1: torch.Tensor, alpha: float) -> torch.Tensor:
        return self.isp.meter_phase2(source, gathered1, alpha)

    def finalize(self, gathered1: torch.Tensor, gathered2: torch.Tensor, alpha: float, out=None) -> None:
        self.isp.meter_finalize(gathered1, gathered2, alpha, out)


def exchange(record: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather one small per-rank record -> (world, len(record)), rank order"""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return record.reshape(1, -1).clone()
    out = torch.empty(world * record.numel(), dtype=record.dtype, device=record.device)
    dist.all_gather_into_tensor(out, record.contiguous().view(-1), group=group)
    return out.view(world, record.numel())


class PeerExchangeUnavailable(RuntimeError):
    """raised on EVERY rank when at least one rank could not map its peers' mailboxes (no P2P / IPC between the GPUs)"""


class PeerExchange:
    """All-gather of the two shared-exposure records through peer-mapped mailboxes over NVLink (csrc/exchange.cu):
    two 1-warp kernels per gather on the current stream, no host call on the data path, CUDA-graph capturable.
    One process per GPU of ONE node; the CUDA IPC handles travel once through ``torch.distributed``.  Without a
    process group (single rank) the mailbox is local and the exchange degenerates to a copy."""

    def __init__(self, device, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self.device, self.group = _lib, torch.device(device), group
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        assert self.world <= 32, "PeerExchange supports up to 32 ranks"
        handle = (C.c_ubyte * 64)()
        self._own = C.c_void_p()
        error = None
        with torch.cuda.device(self.device):
            try:
                _lib.check(_lib.lib.b200isp_mailbox_create(self.world, C.byref(self._own), handle), "mailbox_create")
            except Exception as e:                # keep going: the collectives below must run on every rank
                error = e
            handles = [bytes(handle)]
            if multi:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle), group=group)
            self._peers = (C.c_void_p * self.world)()
            for r, h in enumerate(handles):
                if error is not None:
                    break
                if r == self.rank:
                    self._peers[r] = self._own.value
                else:
                    p = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(h)
                    try:
                        _lib.check(_lib.lib.b200isp_mailbox_open(buf, C.byref(p)), f"mailbox_open(rank {r})")
                    except Exception as e:
                        error = e
                    self._peers[r] = p.value
            self.gathered = {1: torch.empty((self.world, 2), dtype=torch.float32, device=self.device),
                             2: torch.empty((self.world, 8), dtype=torch.float32, device=self.device)}
            torch.cuda.synchronize(self.device)
            if multi:                             # agree: every mailbox is open everywhere before anyone posts
                ok = torch.tensor([0 if error is not None else 1], device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if int(ok.item()) == 0:
                    raise PeerExchangeUnavailable(f"rank {self.rank}: peer mailboxes unavailable ({error})")
            elif error is not None:
                raise error

    def __call__(self, record: torch.Tensor, kind: int) -> torch.Tensor:
        """kind 1: record of 2 floats, kind 2: 8 floats -> (world, 2 | 8) in rank order (a reused buffer: valid until
        the next exchange of the same kind on this stream)"""
        assert record.is_cuda and record.dtype == torch.float32 and record.numel() == (2 if kind == 1 else 8)
        out = self.gathered[kind]
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib.b200isp_mailbox_exchange(
                record.data_ptr(), kind, self._peers, self.world, self.rank, out.data_ptr(),
                self._lib.stream_ptr(self.device)), "mailbox_exchange")
        return out

    def check(self) -> None:
        """raises if a bounded wait expired (a peer never posted); synchronises the current stream"""
        with torch.cuda.device(self.device):
            st = self._lib.lib.b200isp_mailbox_error(self._own, self.world, self._lib.stream_ptr(self.device))
        if st == 1:
            raise RuntimeError(f"rank {self.rank}: shared-exposure exchange timed out waiting for a peer")
        self._lib.check(st, "mailbox_error")


def shared_metering(backend, source, group=None, alpha: Optional[float] = None, out=None, peer: Optional[PeerExchange] = None) -> None:
    """One joint metering update over the frames of ALL ranks (each rank passes its own ``source``).
    ``alpha`` / ``out``: given by the look-ahead pipeline (weight of the previous metrics, second metrics buffer).
    ``peer``: exchange through NVLink mailboxes instead of ``torch.distributed`` all-gathers."""
    if alpha is None:
        alpha = backend.begin()
    source = list(source)
    if peer is not None and isinstance(backend, CudaMeteringBackend) and source[0].dtype == torch.uint8 and source[0].ndim == 2:
        # packed12 frames + NVLink mailboxes: the two exchanges run inside the metering kernels (2 launches, csrc/metering.cuh)
        backend.isp.meter_packed12_shared(source, alpha, peer, out)
        return
    ex = (lambda rec, kind: peer(rec, kind)) if peer is not None else (lambda rec, kind: exchange(rec, group))
    g1 = ex(backend.phase1(source), 1)
    g2 = ex(backend.phase2(source, g1, alpha), 2)
    if out is None:
        backend.finalize(g1, g2, alpha)
    else:
        backend.finalize(g1, g2, alpha, out)


class SharedExposure:
    """Wraps a ``Camera16`` / ``Camera32`` so that its tone-mapping calls meter jointly with the other
    ranks of ``group`` (rig-wide shared exposure).  Same call signatures as the wrapped ISP; everything
    else (``set``, ``load_*``, attributes) is forwarded."""

    def __init__(self, isp, group=None, backend=None, exchange: str = "auto"):
        """exchange: "peer" = NVLink mailboxes (PeerExchange; CUDA ranks of one node), "nccl" = two
        ``torch.distributed`` all-gathers per update (any backend), "auto" = peer for the CUDA backend"""
        self.isp = isp
        self.group = group
        self.backend = backend if backend is not None else CudaMeteringBackend(isp)
        if exchange == "auto":
            exchange = "peer" if backend is None else "nccl"
        assert exchange in ("peer", "nccl")
        self.peer = None
        self.check_every = 256                    # process_packed12 calls between two PeerExchange.check() (a stream sync)
        if exchange == "peer":
            try:
                self.peer = PeerExchange(isp.device, group)
            except PeerExchangeUnavailable:       # raised on every rank alike: all fall back to the all-gather exchange
                self.peer = None

    def __getattr__(self, name):
        return getattr(self.isp, name)

    @property
    def metrics(self):
        return self.isp.metrics

    def update_metering(self, images: Sequence[torch.Tensor]) -> None:
        shared_metering(self.backend, images, self.group, peer=self.peer)

    def tonemap_reinhard(self, images, gamma: float = 1.0, intensity: float = 1.0, light_adapt: float = 1.0,
                         color_adapt: float = 0.0, **kw):
        self.update_metering(images)
        return self.isp.tonemap_reinhard(images, gamma, intensity, light_adapt, color_adapt, update_metering=False, **kw)

    def tonemap_linear(self, images, gamma: float = 1.0, **kw):
        self.update_metering(images)
        return self.isp.tonemap_linear(images, gamma, update_metering=False, **kw)

    def process_packed12(self, frames, tonemap: str = "reinhard", ids_format: bool = False, **kw):
        """fused path: the joint statistics come straight from the packed frames of every rank"""
        # the ISP's look-ahead pipeline drives the joint metering: the exchange of batch k+1 then runs on the side stream
        # under the sweep of batch k.  Frames the fused sweep cannot take (resize, ragged width) go through the ISP's
        # staged kernels, which call the same ``meter_fn`` on the demosaiced (resized) images.
        meter = lambda fs, alpha, out, cooperative: shared_metering(self.backend, fs, self.group, alpha, out, self.peer)
        self._steps = getattr(self, "_steps", 0) + 1
        if self.peer is not None and self._steps % self.check_every == 0:
            self.peer.check()                     # a bounded mailbox wait expired on this rank: raise instead of drifting
        return self.isp.process_packed12(frames, tonemap=tonemap, ids_format=ids_format, meter_fn=meter, **kw)

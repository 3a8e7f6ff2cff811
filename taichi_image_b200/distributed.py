"""Multi-GPU camera sharding: one process per GPU, one camera stream (or a contiguous
group of streams) per rank, no bulk pixel traffic between GPUs.  The only exchange is the optional
rig-wide shared-exposure metering, see ``SharedExposure`` (added with the multi-GPU milestone)."""
from __future__ import annotations


def shard_cameras(n_cameras: int, world_size: int, rank: int) -> range:
    """Contiguous balanced partition of camera indices: the first ``n % world`` ranks get one extra."""
    base, extra = divmod(n_cameras, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))

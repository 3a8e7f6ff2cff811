"""Rig ingestion and output stage around the fused ISP (reference: scripts/tonemap_scan.py:63-100,
:151-179).

The reference's rig driver reads one raw file per camera and time step on a thread pool (``load_raw_bytes``,
``load_images_iter``), views the bytes as ``(H, 1.5 W)`` packed12, runs the ISP, concatenates the cameras' outputs into a
grid image (``concat_image_grid``) and copies that to the host.  The same three pieces, B200-shaped:

* ``RawFrameReader``: a thread pool reads the files of time step k + ``depth`` - 1 straight into PINNED host buffers
  (``readinto``, no intermediate ``bytes`` object, no pageable staging copy) while step k is processed; the pinned
  tensors go to ``pipeline.RigPipeline.submit`` or ``.to(device, non_blocking=True)`` as they are.
* ``GridOutput``: the grid image is allocated once; every camera's output is a row-pitched ``(H, W, 3)`` view of its
  tile, passed as ``out=`` to ``process_packed12`` -- the fused sweep writes the tiles in place, the concatenation
  costs nothing (the reference's ``torch.concat`` re-reads and re-writes every output byte twice).
* ``concat_image_grid``: the reference function itself, for outputs that already exist.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Iterator, List, Sequence, Tuple

import numpy as np
import torch

from .dtypes import as_dtype, u8


def load_raw_bytes(filepath, out: torch.Tensor) -> torch.Tensor:
    """scripts/tonemap_scan.py:63-67 without the copies: the file's bytes land in ``out`` (a pinned uint8 tensor, any shape
    with as many elements as the file has bytes)"""
    buf = out.view(-1).numpy()
    with open(filepath, "rb", buffering=0) as f:
        got = f.readinto(memoryview(buf))
        while got < buf.size:
            more = f.readinto(memoryview(buf)[got:])
            if not more:
                break
            got += more
    assert got == buf.size, f"{filepath}: expected {buf.size} bytes, read {got}"
    return out


class RawFrameReader:
    """``load_images_iter`` (scripts/tonemap_scan.py:69-88): iterate over time steps, each a list of packed12 frames (one per
    camera folder) as pinned ``(height, 1.5 * width)`` uint8 tensors.  ``depth`` buffer sets rotate: the files of the
    following steps are read by the pool while the caller works on the current one; a yielded list stays valid until
    ``depth - 1`` further steps have been yielded."""

    def __init__(self, folders: Sequence, names: Sequence[str], height: int, width: int, depth: int = 3, workers: int = 0):
        assert depth >= 2 and width % 2 == 0
        self.folders, self.names = [Path(f) for f in folders], list(names)
        self.shape = (height, width * 3 // 2)
        self.slots = [[torch.empty(self.shape, dtype=torch.uint8, pin_memory=torch.cuda.is_available()) for _ in self.folders]
                      for _ in range(depth)]
        self.workers = workers or min(32, max(4, len(self.folders)))

    def __len__(self):
        return len(self.names)

    def __iter__(self) -> Iterator[Tuple[str, List[torch.Tensor]]]:
        depth = len(self.slots)
        with ThreadPoolExecutor(max_workers=self.workers) as pool:
            def submit(i):
                slot = self.slots[i % depth]
                return [pool.submit(load_raw_bytes, folder / self.names[i], buf) for folder, buf in zip(self.folders, slot)]
            pending = [submit(i) for i in range(min(depth - 1, len(self.names)))]
            for i, name in enumerate(self.names):
                frames = [f.result() for f in pending.pop(0)]
                nxt = i + depth - 1
                if nxt < len(self.names):
                    pending.append(submit(nxt))          # reuses the slot yielded depth - 1 steps ago
                yield name, frames


def grid_shape(n_images: int, rows: int) -> Tuple[int, int]:
    """(grid rows, grid columns) of concat_image_grid (scripts/tonemap_scan.py:91-93)"""
    n_cols = (n_images + rows - 1) // rows
    return (n_images + n_cols - 1) // n_cols, n_cols


def concat_image_grid(images: Sequence[torch.Tensor], rows: int) -> torch.Tensor:
    """scripts/tonemap_scan.py:91-100"""
    n_cols = (len(images) + rows - 1) // rows
    return torch.concat([torch.concat(list(images[i:i + n_cols]), dim=1) for i in range(0, len(images), n_cols)], dim=0)


class GridOutput:
    """The camera grid of ``concat_image_grid`` allocated once; ``tiles`` are the per-camera ``(H, W, 3)`` views to pass as
    ``out=`` to ``process_packed12``.  All grid rows must be full (``n_images`` a multiple of the column count), like the
    reference needs for its ``torch.concat``."""

    def __init__(self, n_images: int, rows: int, height: int, width: int, dtype=u8, device="cuda"):
        g_rows, n_cols = grid_shape(n_images, rows)
        assert g_rows * n_cols == n_images, "the grid must be rectangular"
        dt = as_dtype(dtype)
        assert (width * 3 * dt.itemsize) % 16 == 0, "tile rows must start 16-byte aligned"
        self.image = torch.empty((g_rows * height, n_cols * width, 3), dtype=dt.torch, device=device)
        self.tiles = [self.image[(i // n_cols) * height:(i // n_cols + 1) * height, (i % n_cols) * width:(i % n_cols + 1) * width]
                      for i in range(n_images)]


def find_folder_images(folder, suffix=".raw") -> Tuple[List[Path], List[str]]:
    """camera folders below ``folder`` and the file names present in all of them (scripts/tonemap_scan.py:39-60)"""
    folders = sorted(p for p in Path(folder).iterdir() if p.is_dir())
    common = None
    for f in folders:
        names = {p.name for p in f.iterdir() if p.suffix == suffix}
        common = names if common is None else common & names
    return folders, sorted(common or [])

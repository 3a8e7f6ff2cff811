"""CUDA-graph replay of the camera-stream step ("capture the per-frame sequence").

A rig processes the SAME device buffers every time step (the capture / H2D stage refills them), so one step --
sweep of batch k on the main stream, metering (and, with ``distributed.SharedExposure``, the two exposure
all-gathers) of batch k+1 on a high-priority side stream -- is a fixed launch sequence.  ``GraphedStream`` captures
it once per metrics-buffer parity and replays it: one ``cudaGraphLaunch`` per step instead of ~10 Python / ctypes /
NCCL enqueues, which is what keeps a 2..8-GPU rig from being host-bound (each step is only ~0.2 ms of GPU time).

Double-buffered ingest (``next_frames`` given): the two buffer sets alternate -- step k sweeps set k % 2 (batch k) and
meters set (k + 1) % 2, which the caller has filled with batch k + 1 before calling ``step()``.  The semantics are then
exactly those of ``ISP.process_packed12(batch_k, lookahead=batch_k+1)`` called in a loop: batch k is tone-mapped with
the metrics that already include batch k (camera_isp.py:376-413), the moving average advances once per step.
``isp.metrics`` always refers to the metrics of the batch the NEXT ``step()`` will tone-map.

Single-buffered ingest (``next_frames=None``): sweep and look-ahead metering read the SAME buffers, so batch k + 1 is
tone-mapped with metrics whose last update saw batch k (twice) instead of batch k + 1 -- the exposure LAGS the scene by
one step.  That is exact only for a static scene; it is kept for callers that cannot afford a second input set.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .dtypes import as_dtype, u8


class GraphedStream:
    def __init__(self, isp, frames: Sequence[torch.Tensor], outs: Sequence[torch.Tensor], tonemap: str = "reinhard",
                 dtype=u8, next_frames: Optional[Sequence[torch.Tensor]] = None, rows_per_task: int = 0,
                 next_outs: Optional[Sequence[torch.Tensor]] = None, ids_format: bool = False, **tonemap_args):
        """isp: Camera16 / Camera32 or a distributed.SharedExposure around one.  frames / outs: the device buffers of
        the even steps (batch 0 must already be in ``frames``).  next_frames / next_outs: the buffers of the odd steps
        (double-buffered ingest / egress); ``next_outs`` defaults to ``outs``, ``next_frames=None`` selects the
        single-buffered mode with its one-step exposure lag (module docstring)."""
        self.wrapper = isp if hasattr(isp, "backend") else None
        base = isp.isp if self.wrapper is not None else isp
        self.isp = base
        self.double_buffered = next_frames is not None
        self.ids_format = bool(ids_format)        # packed layout of the stream's frames (decoded in the sweep's row loader)
        self.F = [list(frames), list(next_frames) if next_frames is not None else list(frames)]
        self.O = [list(outs), list(next_outs) if next_outs is not None else list(outs)]
        assert all(base._fused_ok(f, self.ids_format) for f in self.F[0] + self.F[1]), \
            "GraphedStream needs frames the fused sweep accepts (width % 8 == 0; IDS layout: Malvar, no resize)"
        base._ids_layout = self.ids_format        # read by every parameter block built during warm-up and capture
        self.tonemap, self.out_dtype, self.tm = tonemap, as_dtype(dtype), dict(tonemap_args)
        self.rows_per_task = rows_per_task
        # the ISP's transform is applied where the fused call applies it itself (flips in the sweep's store, transposing
        # transforms in the normalise pass of the one-sweep Reinhard -> u8 forms); ``outs`` then have the transformed shape
        self.flip = base._fused_flip(self.F[0][0].shape[0], tonemap, self.out_dtype, self.tm.get("gamma", 1.0),
                                     self.tm.get("color_adapt", 0.0), self.ids_format)
        from .interpolate import ImageTransform
        assert self.flip or base.transform == ImageTransform.none, \
            "GraphedStream: this ISP transform needs the separate transform kernel here -- use process_packed12"
        dev = base.device
        with torch.cuda.device(dev):
            self.main = torch.cuda.Stream(dev)
            self.side = torch.cuda.Stream(dev, priority=-1)     # metering CTAs go ahead of the sweep's undispatched CTAs
            self.M = [torch.zeros(9, dtype=torch.float32, device=dev) for _ in range(2)]
            scratch = [torch.zeros(9, dtype=torch.float32, device=dev) for _ in range(2)]
            ahead = 1.0 - float(base.moving_alpha)
            self.main.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.main):
                # metrics of the first batch (eager): first-call semantics of camera_isp.py:376-385
                if base.metrics is None:
                    base.metrics = self.M[0]
                    self._meter(self.F[0], 0.0, None, cooperative=True)
                else:
                    self.M[0].copy_(base.metrics)
                    base.metrics = self.M[0]
                    self._meter(self.F[0], ahead, None, cooperative=True)
                # eager warm-up of the step body on scratch metrics: allocates the per-stream workspaces, the sample
                # cache and the exchange buffers outside the capture
                scratch[0].copy_(self.M[0])
                for p in (0, 1):
                    base.metrics = scratch[p]
                    self._body(p, scratch[1 - p])
            torch.cuda.synchronize(dev)
            self.graphs = []
            for p in (0, 1):
                base.metrics = self.M[p]
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.main):
                    self._body(p, self.M[1 - p])
                self.graphs.append(g)
            base.metrics = self.M[0]
        self.parity = 0

    # -- one step: sweep(k) on the current stream, metering(k+1) forked onto the side stream, joined at the end
    def _meter(self, frames, alpha, out, cooperative):
        if self.wrapper is not None:
            from .distributed import shared_metering
            shared_metering(self.wrapper.backend, frames, self.wrapper.group, alpha, out, self.wrapper.peer)
        else:
            self.isp.meter_packed12(frames, alpha, out, cooperative)

    def _body(self, p, metrics_next):
        isp = self.isp
        cur = torch.cuda.current_stream(isp.device)
        self.side.wait_stream(cur)
        isp._run_fused(self.F[p], self.tonemap, self.out_dtype, self.O[p], self.tm, update_metering=False,
                       rows_per_task=self.rows_per_task, flip=self.flip)
        with torch.cuda.stream(self.side):
            self._meter(self.F[1 - p], 1.0 - float(isp.moving_alpha), metrics_next, cooperative=False)
        cur.wait_stream(self.side)

    @property
    def frames(self):
        """the input buffers the NEXT ``step()`` sweeps"""
        return self.F[self.parity]

    @property
    def next_frames(self):
        """the input buffers the next ``step()`` meters = where batch k + 1 must be before that step is enqueued"""
        return self.F[1 - self.parity]

    def step(self):
        """Tone-map the batch in ``frames`` into its output buffers (returned) and meter the batch in ``next_frames``
        (asynchronous, on the current stream)."""
        outs = self.O[self.parity]
        self.graphs[self.parity].replay()
        self.parity ^= 1
        self.isp.metrics = self.M[self.parity]
        return outs

"""CUDA-graph replay of the camera-stream step ("capture the per-frame sequence").

A rig processes the SAME device buffers every time step (the capture / H2D stage refills them), so one step --
sweep of batch k on the main stream, metering (and, with ``distributed.SharedExposure``, the two exposure
all-gathers) of batch k+1 on a high-priority side stream -- is a fixed launch sequence.  ``GraphedStream`` captures
it once per metrics-buffer parity and replays it: one ``cudaGraphLaunch`` per step instead of ~10 Python / ctypes /
NCCL enqueues, which is what keeps a 2..8-GPU rig from being host-bound (each step is only ~0.2 ms of GPU time).

Semantics are exactly those of ``ISP.process_packed12(frames, lookahead=next_frames)`` called in a loop: batch k is
tone-mapped with the metrics that already include batch k (camera_isp.py:376-413), the moving average advances once
per step.  ``isp.metrics`` always refers to the metrics of the batch the NEXT ``step()`` will tone-map.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .dtypes import as_dtype, u8


class GraphedStream:
    def __init__(self, isp, frames: Sequence[torch.Tensor], outs: Sequence[torch.Tensor], tonemap: str = "reinhard",
                 dtype=u8, next_frames: Optional[Sequence[torch.Tensor]] = None, rows_per_task: int = 0,
                 profile: bool = False, **tonemap_args):
        """isp: Camera16 / Camera32 or a distributed.SharedExposure around one.  frames / outs: the fixed device
        buffers of the stream.  next_frames: buffers holding batch k+1 while batch k is processed (double-buffered
        ingest); default: the same buffers (single-buffered ingest refilled between steps)."""
        self.wrapper = isp if hasattr(isp, "backend") else None
        base = isp.isp if self.wrapper is not None else isp
        self.isp = base
        self.frames, self.outs = list(frames), list(outs)
        self.next_frames = list(next_frames) if next_frames is not None else self.frames
        assert all(base._fused_ok(f, False) for f in self.frames + self.next_frames) and not base._resizes, \
            "GraphedStream needs frames the fused sweep accepts (standard layout, width % 8 == 0, no resize)"
        self.tonemap, self.out_dtype, self.tm = tonemap, as_dtype(dtype), dict(tonemap_args)
        self.rows_per_task = rows_per_task
        dev = base.device
        with torch.cuda.device(dev):
            self.main = torch.cuda.Stream(dev)
            self.side = torch.cuda.Stream(dev, priority=-1)     # metering CTAs go ahead of the sweep's undispatched CTAs
            self.M = [torch.zeros(9, dtype=torch.float32, device=dev) for _ in range(2)]
            scratch = [torch.zeros(9, dtype=torch.float32, device=dev) for _ in range(2)]
            self.events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if profile else None
                           for _ in range(2)]
            for ev in self.events:
                if ev is not None:
                    ev[0].record(); ev[1].record()       # torch creates the cudaEvent lazily
            ahead = 1.0 - float(base.moving_alpha)
            self.main.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.main):
                # metrics of the first batch (eager): first-call semantics of camera_isp.py:376-385
                if base.metrics is None:
                    base.metrics = self.M[0]
                    self._meter(self.frames, 0.0, None, cooperative=True)
                else:
                    self.M[0].copy_(base.metrics)
                    base.metrics = self.M[0]
                    self._meter(self.frames, ahead, None, cooperative=True)
                # eager warm-up of the step body on scratch metrics: allocates the per-stream workspaces, the sample
                # cache and NCCL's buffers outside the capture
                scratch[0].copy_(self.M[0])
                for p in (0, 1):
                    base.metrics = scratch[p]
                    self._body(scratch[1 - p], None)
            torch.cuda.synchronize(dev)
            self.graphs = []
            for p in (0, 1):
                base.metrics = self.M[p]
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.main):
                    self._body(self.M[1 - p], self.events[p])
                self.graphs.append(g)
            base.metrics = self.M[0]
        self.parity = 0
        self.launches_per_step = 3 if self.wrapper is None else 5     # sweep + metering kernels (+ fold, finalize)

    # -- one step: sweep(k) on the current stream, metering(k+1) forked onto the side stream, joined at the end
    def _meter(self, frames, alpha, out, cooperative):
        if self.wrapper is not None:
            from .distributed import shared_metering
            shared_metering(self.wrapper.backend, frames, self.wrapper.group, alpha, out, self.wrapper.peer)
        else:
            self.isp.meter_packed12(frames, alpha, out, cooperative)

    def _body(self, metrics_next, events):
        isp = self.isp
        cur = torch.cuda.current_stream(isp.device)
        self.side.wait_stream(cur)
        isp._run_fused(self.frames, self.tonemap, self.out_dtype, self.outs, self.tm, update_metering=False,
                       rows_per_task=self.rows_per_task, profile_events=events)
        with torch.cuda.stream(self.side):
            self._meter(self.next_frames, 1.0 - float(isp.moving_alpha), metrics_next, cooperative=False)
        cur.wait_stream(self.side)

    def step(self):
        """Tone-map the batch in ``frames`` into ``outs`` and meter the batch in ``next_frames`` (asynchronous, on the
        current stream)."""
        self.graphs[self.parity].replay()
        self.parity ^= 1
        self.isp.metrics = self.M[self.parity]
        return self.outs

    def kernel_ms(self):
        """device time of the sweep kernel in the last two steps (needs profile=True and a synchronize)"""
        return [a.elapsed_time(b) for a, b in (e for e in self.events if e is not None)]

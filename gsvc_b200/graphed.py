"""CUDA-graph replay of a rasterizer step on static tensors.

The eager call has one unavoidable host wait per forward — `num_rendered` must come back as a Python
int (ortho_gaussian_renderer/renderer.py:90) — so the host can never run more than one step ahead of
the device, and with the blend kernels at ~0.1 ms the Python/autograd launch path between "count
published" and "backward enqueued" shows up as idle SM time.  A training loop whose tensors keep
their addresses (parameters updated in place by the optimizer, a fixed seed-gradient / loss buffer)
can instead capture forward + backward once and replay it: one graph launch per step, no host wait,
the 8 kernels back to back with their programmatic-dependent-launch edges preserved.

    step = GraphedStep(rasterizer, params, dL)        # params: dict of static CUDA tensors (GRAD_LAYOUT keys)
    color, radii, grads = step()                       # replay; outputs are static tensors, overwritten each call
    ...
    if not step.capacity_ok(): step.recapture()        # after a synchronisation, e.g. once per N steps

The binning capacity is fixed at capture time from an eager warm-up (+50 % + 64 Ki instances);
`capacity_ok()` compares it with the instance count the device published for the last replay.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import rasterizer as R
from .rasterizer import RasterizerError
from .sharding import GRAD_LAYOUT, GRAD_WIDTH, packed_backward
from .views import ViewBatch, rasterize_views


class GraphedStep:
    def __init__(self, rast, params: Dict[str, torch.Tensor], dL: Optional[torch.Tensor],
                 packed: Optional[torch.Tensor] = None, warmup: int = 2, exchange=None):
        """`rast`: a GaussianRasterizer (one view, dL [3,H,W]) or a views.ViewBatch (all its views in one chain,
        dL [n_out,3,H,W], gradients summed over the views).  `dL` None: forward only.  `packed`: optional [P,14]
        buffer the backward writes (sharding.GRAD_LAYOUT); allocated here if omitted.  `exchange`: a
        sharding.SwitchAllReduce over P*14 floats — the backward then writes into ITS buffer and carries the sum over the
        ranks in its own launches (ViewBatch steps only; every rank builds and replays the same step)."""
        self.rast, self.params, self.dL = rast, params, dL
        self.batch = rast if isinstance(rast, ViewBatch) else None
        m = params["means3D"]
        if not m.is_cuda:
            raise RasterizerError("GraphedStep needs CUDA tensors: gsvc_b200 has no CPU fallback")
        self.device, self.P = m.device, int(m.shape[0])
        self.backward = dL is not None
        self.exchange = exchange
        if exchange is not None:
            if self.batch is None or not self.backward:
                raise RasterizerError("a step carries the exchange in its batched-view backward: pass a ViewBatch and dL")
            if packed is not None:
                raise RasterizerError("with `exchange` the packed buffer is the exchange's own")
            packed = exchange.buffer().view(self.P, GRAD_WIDTH)
            self._exchange_generation = exchange.generation
        if self.backward and packed is None:
            packed = torch.empty((self.P, GRAD_WIDTH), dtype=torch.float32, device=self.device)
        self.packed = packed
        self.warmup = max(int(warmup), 1)
        self.graph = None
        self.color = self.radii = None
        # the graph's own pinned count word: its scan kernel publishes num_rendered here on every replay, whatever
        # other graphs or eager calls of this thread do meanwhile
        self.slot = R.CountSlot()
        self.capacity = None
        self.recapture()

    def _call(self, p, means2D):
        if self.batch is not None:
            return rasterize_views(self.batch, means3D=p["means3D"], opacities=p["opacities"], means2D=None,
                                   colors_precomp=p["colors_precomp"], scales=p["scales"], rotations=p["rotations"])
        return self.rast(means3D=p["means3D"], means2D=means2D, shs=None, colors_precomp=p["colors_precomp"],
                         opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)

    def _run(self):
        p = self.params
        if not self.backward:
            with torch.no_grad():
                return self._call(p, p["means3D"])
        leaves = {k: p[k].detach().requires_grad_(True) for k, _ in GRAD_LAYOUT}
        means2D = None if self.batch is not None else torch.zeros_like(leaves["means3D"], requires_grad=True)
        color, radii, n = self._call(leaves, means2D)
        with packed_backward(self.packed, exchange=self.exchange):
            torch.autograd.grad(color, [leaves[k] for k, _ in GRAD_LAYOUT], grad_outputs=self.dL)
        return color, radii, n

    def recapture(self) -> None:
        """(Re)build the graph: eager warm-up on a side stream (sets the capacity hint), then capture."""
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                _, _, n = self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        R.overflow_events(self.device)      # an eager warm-up call that outgrew its hint re-ran by itself: not ours
        self.num_rendered_at_capture = n
        self.graph = torch.cuda.CUDAGraph()
        with R.own_count_slot(self.slot), torch.cuda.graph(self.graph):
            self.color, self.radii, _ = self._run()
        rs = self.batch.settings[0] if self.batch is not None else self.rast.raster_settings
        key = (self.device.index, self.P, int(rs.image_height), int(rs.image_width))
        self.capacity = R._captured_caps.get(key + ((self.batch.n_views,) if self.batch is not None else ()))

    def __call__(self):
        if self.exchange is not None and self.exchange.generation != self._exchange_generation:
            raise RasterizerError("the exchange's symmetric buffer was reallocated after this step was captured "
                                  "(SwitchAllReduce.sum_ outgrew it): build a new GraphedStep")
        self.graph.replay()
        return self.color, self.radii, self.packed

    def num_rendered(self) -> int:
        """Instance count of this graph's last replay (call after a synchronisation)."""
        return self.slot.value()

    def capacity_ok(self) -> bool:
        """After a synchronisation: no replay since the last check ran out of the instance capacity fixed at
        capture time (sticky device-side counter), and the last replay's published count fits too."""
        if R.overflow_events(self.device) > 0:
            return False
        return self.capacity is None or self.num_rendered() <= self.capacity


class FrameStreamer:
    """Throughput rendering of INDEPENDENT frames (video decode / evaluation: the Gaussians are fixed, only the
    frame's view matrices change — utils/report_utils.py:297-319 renders them one at a time with a synchronise
    around each).  `n_streams` forward-only graphs, each with private view-matrix tensors, are replayed round-robin
    on their own CUDA streams, so the latency-bound binning kernels of frame i+1 run under the issue-bound blend of
    frame i (measured: 170 -> 147 us per 1080p frame with 4 streams, config 2).

        streamer = FrameStreamer(front0, back0, params, n_streams=4)       # settings of any frame of the cube
        for i, (V_front, V_back) in enumerate(frames):                      # logical 4x4 view matrices (tensors)
            image = streamer.render(V_front, V_back)                        # [3,H,W], valid after streamer.wait(image)
        streamer.synchronize()

    `render` returns the stream's static output tensor: it is overwritten n_streams frames later, so consume it
    (on its stream, or after `wait`) before that."""

    def __init__(self, front, back, params: Dict[str, torch.Tensor], n_streams: int = 4, toast: bool = True):
        m = params["means3D"]
        if not m.is_cuda:
            raise RasterizerError("FrameStreamer needs CUDA tensors: gsvc_b200 has no CPU fallback")
        self.device, self.n = m.device, int(n_streams)
        self.streams, self.steps, self.mats, self.events = [], [], [], []
        cur = torch.cuda.current_stream(self.device)
        for _ in range(self.n):
            s = torch.cuda.Stream(self.device)
            vf = front.viewmatrix.to(self.device, torch.float32).contiguous().clone()
            vb = back.viewmatrix.to(self.device, torch.float32).contiguous().clone()
            f, b = front._replace(viewmatrix=vf), back._replace(viewmatrix=vb)
            batch = ViewBatch.toast(f, b) if toast else ViewBatch([f, b])
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                step = GraphedStep(batch, params, None)
            self.streams.append(s)
            self.steps.append(step)
            self.mats.append((vf, vb))
            self.events.append(None)
        self.i = 0

    def render(self, V_front: torch.Tensor, V_back: torch.Tensor) -> torch.Tensor:
        k = self.i % self.n
        self.i += 1
        s = self.streams[k]
        s.wait_stream(torch.cuda.current_stream(self.device))      # V_front / V_back may have just been produced
        with torch.cuda.stream(s):
            self.mats[k][0].copy_(V_front, non_blocking=True)
            self.mats[k][1].copy_(V_back, non_blocking=True)
            color, _, _ = self.steps[k]()
            self.events[k] = s.record_event()
        self._last = k
        return color[0] if color.shape[0] == 1 else color

    def wait(self, image: torch.Tensor = None) -> None:
        """Make the current stream wait for the most recently submitted frame."""
        torch.cuda.current_stream(self.device).wait_event(self.events[self._last])

    def synchronize(self) -> None:
        for s in self.streams:
            s.synchronize()

    def capacity_ok(self) -> bool:
        """After `synchronize()`: every frame since the last check fitted the instance capacity of the graphs."""
        return R.overflow_events(self.device) == 0

    def recapture(self) -> None:
        for s, step in zip(self.streams, self.steps):
            with torch.cuda.stream(s):
                step.recapture()

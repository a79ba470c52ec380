"""Frame-sharded rendering of a TSW window across the GPUs of one box.

The reference is single-GPU (utils/general_utils.py:153) but already evaluates independent views
per step (/root/reference/pipeline/train.py:353-387: front/back view of two frames); views only
couple through gradient accumulation into the shared Gaussian parameters.  So: replicate the
Gaussians, give rank r the frames {f : f mod G == r} of the window, render both views of each
locally, and sum the [P,14] parameter-gradient buffer with ONE NCCL all-reduce per step.
There is no other exchange on this path (forward-only decode needs none at all).
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, List, Sequence

import torch
import torch.distributed as dist

# layout of the packed per-Gaussian gradient buffer: 14 fp32 per Gaussian (56 B)
GRAD_LAYOUT = (("means3D", 3), ("colors_precomp", 3), ("opacities", 1), ("scales", 3), ("rotations", 4))
GRAD_WIDTH = sum(w for _, w in GRAD_LAYOUT)


def frames_for_rank(frames: Sequence[int], rank: int, world: int) -> List[int]:
    """Round-robin frame assignment (BASELINE config 3)."""
    return [f for i, f in enumerate(frames) if i % world == rank]


def pack_grads(grads: Dict[str, torch.Tensor], out: torch.Tensor = None) -> torch.Tensor:
    P = grads["means3D"].shape[0]
    if out is None:
        out = torch.empty((P, GRAD_WIDTH), dtype=torch.float32, device=grads["means3D"].device)
    c = 0
    for name, w in GRAD_LAYOUT:
        out[:, c:c + w] = grads[name].reshape(P, w)
        c += w
    return out


@contextlib.contextmanager
def packed_backward(buf: torch.Tensor):
    """While active, the rasterizer backward writes the GRAD_LAYOUT gradients directly into `buf` ([P,14] fp32,
    contiguous, on the rasterizer's device) and returns views of it, so the all-reduce needs no pack pass."""
    from . import rasterizer
    prev = rasterizer._packed_target.buf
    rasterizer._packed_target.buf = buf
    try:
        yield buf
    finally:
        rasterizer._packed_target.buf = prev


def unpack_grads(buf: torch.Tensor) -> Dict[str, torch.Tensor]:
    out, c = {}, 0
    for name, w in GRAD_LAYOUT:
        out[name] = buf[:, c:c + w]
        c += w
    return out


def allreduce_grads(buf: torch.Tensor, average: bool = False) -> torch.Tensor:
    """Sum (optionally average) the packed gradient buffer over all ranks: fp32, one collective."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        if average:
            buf /= dist.get_world_size()
    return buf


ViewFn = Callable[[int, bool], Dict[str, torch.Tensor]]


def render_window_grads(frames: Sequence[int], view_grads: ViewFn, rank: int = 0, world: int = 1,
                        both_views: bool = True) -> torch.Tensor:
    """Accumulate this rank's frames (front and back view each), then all-reduce.

    `view_grads(frame_id, back)` renders one view forward+backward and returns the parameter
    gradients keyed like GRAD_LAYOUT.  Returns the summed [P,14] buffer, identical on every rank.
    """
    total = None
    for f in frames_for_rank(frames, rank, world):
        for back in ((False, True) if both_views else (False,)):
            g = pack_grads(view_grads(f, back))
            total = g if total is None else total.add_(g)
    if total is None:
        raise ValueError("rank has no frames; give every rank at least one frame of the window")
    return allreduce_grads(total)

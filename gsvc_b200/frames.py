"""Frame-cube geometry and the synthetic workload generator of SURVEY.md §8d.

Mirrors, without PyGLM (not installed), what the reference computes in
/root/reference/frame_cube/frame.py:18-43 (make_view_matrix) and :98-101,156-190 (cube geometry,
get_z_frame): a video is a cube (x,y = pixels, z = time) and every frame is an orthographic camera
at z looking along -z (front view) or +z (back view, x-mirrored).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import math
import torch


def make_view_matrix(x: float = 0.0, y: float = 0.0, z: float = 0.0, plane: str = "xy"):
    """Closed form of glm.lookAt for the reference's 'xy' plane cameras (frame.py:18-43).

    Returns (view_matrix, view_matrix_s, cam_pos) exactly as the reference stores them:
    `np.array(glm.mat4)` is column-major, i.e. the TRANSPOSE of the mathematical V, so that the
    caller's `view_matrix.permute(1, 0)` (renderer.py:77) is the logical V with
    p_view = V[:3,:3] @ p + V[:3,3].
    """
    if plane != "xy":
        raise ValueError("only the 'xy' plane is used by the reference pipelines (frame.py:156-190)")
    # front: f=(0,0,-1), s=(1,0,0), u=(0,1,0)  →  rows s,u,-f ; t = (-s·eye, -u·eye, f·eye)
    V = torch.tensor([[1.0, 0.0, 0.0, -x],
                      [0.0, 1.0, 0.0, -y],
                      [0.0, 0.0, 1.0, -z],
                      [0.0, 0.0, 0.0, 1.0]], dtype=torch.float32)
    # back: f=(0,0,1), s=(-1,0,0), u=(0,1,0)
    Vs = torch.tensor([[-1.0, 0.0, 0.0, x],
                       [0.0, 1.0, 0.0, -y],
                       [0.0, 0.0, -1.0, z],
                       [0.0, 0.0, 0.0, 1.0]], dtype=torch.float32)
    cam_pos = torch.tensor([x, y, z], dtype=torch.float32)
    return V.t().contiguous(), Vs.t().contiguous(), cam_pos


@dataclass
class Frame:
    """Field-for-field the reference's Frame dataclass (frame.py:46-59)."""
    image_id: int
    plane: str
    image: Optional[torch.Tensor]
    x_min: float
    y_min: float
    z: float
    image_width: int
    image_height: int
    view_matrix: torch.Tensor
    view_matrix_s: torch.Tensor
    scale: float
    cam_pos: torch.Tensor


@dataclass
class CubeGeometry:
    """frame.py:98-101: scale = max(H,W,F)/2; x_min = -W/2/scale; y_min = -H/2/scale."""
    width: int
    height: int
    frames: int

    @property
    def scale(self) -> float:
        return max(self.height, self.width, self.frames) / 2

    @property
    def x_min(self) -> float:
        return -self.width / 2 / self.scale

    @property
    def y_min(self) -> float:
        return -self.height / 2 / self.scale

    def z_of(self, image_id: int) -> float:
        return (image_id - self.frames / 2) / self.scale  # frame.py:158

    def frame(self, image_id: int) -> Frame:
        z = self.z_of(image_id)
        vm, vms, cam = make_view_matrix(z=z, plane="xy")
        return Frame(image_id=image_id, plane="xy", image=None, x_min=self.x_min, y_min=self.y_min, z=z,
                     image_width=self.width, image_height=self.height, view_matrix=vm, view_matrix_s=vms,
                     scale=self.scale, cam_pos=cam)


# BASELINE.json configs: (P, W, H, F)
CONFIGS = {
    1: dict(P=20_000, W=256, H=256, F=256),
    2: dict(P=200_000, W=1920, H=1080, F=600),
    3: dict(P=500_000, W=1920, H=1080, F=600, window=8),
    4: dict(P=1_000_000, W=1920, H=1080, F=600),
    5: dict(P=2_000_000, W=3840, H=2160, F=600),
}


def synthetic_gaussians(P: int, geom: CubeGeometry, frame_lo: int, frame_hi: Optional[int] = None,
                        threshold: float = 0.05, seed: int = 1, device="cpu"):
    """SURVEY.md §8d generator (CPU generator for reproducibility, then moved to `device`).

    x,y uniform over 1.1x the image extent; z uniform in [z(frame_lo) - 1.5 thr, z(frame_hi) + 1.5 thr];
    per-axis sigma_px ~ LogNormal(ln 2, 0.6) clipped to [0.3, 30] px; unit quaternions; opacity ~ U(0.05,1);
    colour ~ U(0,1)^3.
    """
    g = torch.Generator().manual_seed(seed)
    if frame_hi is None:
        frame_hi = frame_lo
    z0, z1 = geom.z_of(frame_lo) - 1.5 * threshold, geom.z_of(frame_hi) + 1.5 * threshold
    u = torch.rand(P, 3, generator=g)
    xs = (u[:, 0] * 2 - 1) * (-1.1 * geom.x_min)
    ys = (u[:, 1] * 2 - 1) * (-1.1 * geom.y_min)
    zs = z0 + u[:, 2] * (z1 - z0)
    means3D = torch.stack([xs, ys, zs], dim=-1)
    sigma_px = torch.exp(math.log(2.0) + 0.6 * torch.randn(P, 3, generator=g)).clamp(0.3, 30.0)
    scales = sigma_px / geom.scale
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=-1, keepdim=True)
    opacities = 0.05 + 0.95 * torch.rand(P, 1, generator=g)
    colors = torch.rand(P, 3, generator=g)
    out = dict(means3D=means3D, scales=scales, rotations=rotations, opacities=opacities, colors_precomp=colors)
    return {k: v.to(device=device, dtype=torch.float32).contiguous() for k, v in out.items()}


def z_interval_table(z_sorted: torch.Tensor, interval: float = 0.01):
    """The stream codec's slab table for anchors sorted by z (utils/encodings.py:827-862 `reorder_and_split`):
    z intervals of width `interval` starting at the rounded minimum, each with its [start, end) index range.
    Returns (z_lo: float, interval: float, starts: LongTensor[n_intervals + 1]) — interval k covers
    z in [z_lo + k*interval, z_lo + (k+1)*interval) = indices [starts[k], starts[k+1])."""
    z = z_sorted.detach().double().cpu()
    if z.numel() and not bool((z[1:] >= z[:-1]).all()):
        raise ValueError("anchors are not sorted by z")
    z_min = float(z.min()) if z.numel() else 0.0
    z_max = float(z.max()) if z.numel() else 0.0
    z_lo = math.floor(z_min / interval) * interval
    n = max(1, int(math.ceil((z_max - z_lo) / interval + 1e-9)) + 1)
    edges = z_lo + interval * torch.arange(n + 1, dtype=torch.float64)
    starts = torch.searchsorted(z, edges, right=False)
    return z_lo, interval, starts


def slab_index_range(table, z_frame: float, threshold: float):
    """Index range [lo, hi) of the z-sorted anchors whose interval can intersect the TSW slab |z - z_frame| <= threshold
    (front and back view share it): whole intervals, one interval of margin on both sides for rounding."""
    z_lo, interval, starts = table
    n = starts.numel() - 1
    k0 = int(math.floor((z_frame - threshold - z_lo) / interval)) - 1
    k1 = int(math.floor((z_frame + threshold - z_lo) / interval)) + 2
    k0, k1 = min(max(k0, 0), n), min(max(k1, 0), n)
    return int(starts[k0]), int(starts[k1])

"""gsvc_b200 — B200-native (sm_100a) orthographic TSW Gaussian rasterizer for GSVC's render() hot path.

  rasterizer   GaussianRasterizationSettings / GaussianRasterizer: the reference-shaped drop-in (renderer.py:63-98)
  views        ViewBatch / rasterize_views / render_toast: a frame's front + back view (or a window) in one chain
  graphed      GraphedStep: CUDA-graph replay of forward + backward on static tensors
  hostpipe     HostStepPipeline: pinned host parameters in, pinned host gradients out, copies overlapped
  sharding     frame-sharded rendering, packed [P,14] gradients, NCCL all-reduce
  frames       cube geometry and the synthetic workload generator
"""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, RasterizerError  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "RasterizerError"]

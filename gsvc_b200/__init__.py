"""gsvc_b200 — B200-native (sm_100a) orthographic TSW Gaussian rasterizer for GSVC's render() hot path."""
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, RasterizerError  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "RasterizerError"]

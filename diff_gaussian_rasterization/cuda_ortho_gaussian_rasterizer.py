"""Drop-in module for /root/reference/ortho_gaussian_renderer/renderer.py:6 and preprocess.py:21."""
from gsvc_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                  rasterize_gaussians)

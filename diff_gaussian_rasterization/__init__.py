"""Import-path shim: the reference imports
`from diff_gaussian_rasterization.cuda_ortho_gaussian_rasterizer import GaussianRasterizationSettings, GaussianRasterizer`
(/root/reference/ortho_gaussian_renderer/renderer.py:6, preprocess.py:21).  The implementation is gsvc_b200."""
from gsvc_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

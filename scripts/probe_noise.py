"""Probe: run-to-run difference of the gradients (order of the float atomics) on a small scene, max over repeats."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.scenes import make_scene, product_settings
from gsvc_b200.rasterizer import GaussianRasterizer
from gsvc_b200.sharding import GRAD_LAYOUT
dev = torch.device("cuda:0")
for P, W, H, F, seed in ((9000, 160, 96, 160, 31), (10000, 176, 112, 176, 37), (12000, 192, 128, 192, 41)):
    scene = make_scene(P=P, W=W, H=H, F=F, seed=seed)
    rast = GaussianRasterizer(raster_settings=product_settings(scene, dev))
    g = {k: v.to(dev) for k, v in scene["gaussians"].items()}
    dL = torch.randn((3, H, W), generator=torch.Generator().manual_seed(2)).to(dev)
    def grads():
        p = {k: g[k].clone().requires_grad_(True) for k, _ in GRAD_LAYOUT}
        m2d = torch.zeros_like(p["means3D"], requires_grad=True)
        c, _, _ = rast(means3D=p["means3D"], means2D=m2d, shs=None, colors_precomp=p["colors_precomp"], opacities=p["opacities"],
                       scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)
        return torch.cat([x.reshape(P, -1) for x in torch.autograd.grad(c, [p[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)], 1)
    ref = grads()
    worst = 0.0
    for _ in range(300):
        worst = max(worst, float((grads() - ref).abs().max() / ref.abs().max()))
    print(f"P={P} {W}x{H}: worst run-to-run rel diff over 300 repeats {worst:.2e}")

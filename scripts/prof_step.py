"""Profiling target: a few eager steps of the bench workload (config 2, one frame = front + back view in one
batched chain, forward + backward), so that `ncu -k regex:gsvc -s <warm-up launches> -c 7` sees one launch of each
kernel of the chain.  Usage: python scripts/prof_step.py [views=2] [steps=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for
from gsvc_b200.rasterizer import GaussianRasterizer
from gsvc_b200.sharding import GRAD_LAYOUT
from gsvc_b200.views import ViewBatch, rasterize_views

views = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(1, dev)
H, W = cfg["H"], cfg["W"]
front, back = settings_for(geom, f0, dev), settings_for(geom, f0, dev, back=True)
dL = torch.randn((1, 3, H, W), device=dev)
for _ in range(steps):
    p = {k: g[k].detach().requires_grad_(True) for k, _ in GRAD_LAYOUT}
    if views == 2:
        img, radii, n = rasterize_views(ViewBatch.toast(front, back), means3D=p["means3D"], opacities=p["opacities"],
                                        colors_precomp=p["colors_precomp"], scales=p["scales"], rotations=p["rotations"])
    else:
        m2d = torch.zeros_like(p["means3D"], requires_grad=True)
        img, radii, n = GaussianRasterizer(raster_settings=front)(
            means3D=p["means3D"], means2D=m2d, shs=None, colors_precomp=p["colors_precomp"], opacities=p["opacities"],
            scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)
        img = img[None]
    torch.autograd.grad(img, [p[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)
torch.cuda.synchronize()
print("num_rendered", n)

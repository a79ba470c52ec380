"""Ad-hoc probe: Python/host overhead of the rasterizer call path (cProfile over train steps)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for
from gsvc_b200.rasterizer import GaussianRasterizer
from gsvc_b200.sharding import GRAD_LAYOUT

dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(1, dev)
rast = GaussianRasterizer(raster_settings=settings_for(geom, f0, dev))
params = {k: v.clone().requires_grad_(True) for k, v in g.items()}
dL = torch.randn((3, cfg["H"], cfg["W"]), device=dev)


def step():
    means2D = torch.zeros_like(params["means3D"], requires_grad=True)
    color, radii, n = rast(means3D=params["means3D"], means2D=means2D, shs=None, colors_precomp=params["colors_precomp"],
                           opacities=params["opacities"], scales=params["scales"], rotations=params["rotations"],
                           cov3D_precomp=None)
    return torch.autograd.grad(color, [params[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step()
torch.cuda.synchronize()
print("wall per step (ms):", (time.perf_counter() - t0) / 300 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

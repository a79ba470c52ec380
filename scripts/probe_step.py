"""Ad-hoc probe: where does a train step's time go (host vs device)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for
from gsvc_b200 import _lib
from gsvc_b200.rasterizer import GaussianRasterizer
from gsvc_b200.sharding import GRAD_LAYOUT

dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(1, dev)
rast = GaussianRasterizer(raster_settings=settings_for(geom, f0, dev))
params = {k: v.clone().requires_grad_(True) for k, v in g.items()}
dL = torch.randn((3, cfg["H"], cfg["W"]), device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

def step(timing):
    t0 = time.perf_counter()
    means2D = torch.zeros_like(params["means3D"], requires_grad=True)
    color, radii, n = rast(means3D=params["means3D"], means2D=means2D, shs=None, colors_precomp=params["colors_precomp"],
                           opacities=params["opacities"], scales=params["scales"], rotations=params["rotations"], cov3D_precomp=None)
    t1 = time.perf_counter()
    grads = torch.autograd.grad(color, [params[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3

for timing in (False, True, False):
    _lib.stage_timing(timing)
    for i in range(8):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a, b, c = step(timing)
        e1.record()
        torch.cuda.synchronize()
        print(f"timing={timing} i={i} fwd_host={a:.3f} ms bwd_host={b:.3f} ms tail_sync={c:.3f} ms events={e0.elapsed_time(e1):.3f} ms", _lib.stage_times() if timing else "")

"""Ad-hoc probe: visible_filter over 1M anchors (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for, THRESHOLD
from gsvc_b200.frames import synthetic_gaussians
from gsvc_b200.rasterizer import GaussianRasterizer
dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(1, dev)
rast = GaussianRasterizer(raster_settings=settings_for(geom, f0, dev))
ga = synthetic_gaussians(1_000_000, geom, f0, f0, threshold=THRESHOLD, seed=4, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for _ in range(5):
    flush.zero_()
    r = rast.visible_filter(means3D=ga["means3D"], scales=ga["scales"], rotations=ga["rotations"], cov3D_precomp=None)
torch.cuda.synchronize()
print(int((r > 0).sum()))

"""Probe: visible_filter / visible_filter_compact stage time over 1M anchors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for, THRESHOLD
from gsvc_b200 import _lib
from gsvc_b200.frames import synthetic_gaussians
from gsvc_b200.rasterizer import GaussianRasterizer
dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(1, dev)
rast = GaussianRasterizer(raster_settings=settings_for(geom, f0, dev))
ga = synthetic_gaussians(1_000_000, geom, f0, f0, threshold=THRESHOLD, seed=4, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for name, fn in (("visible_filter", lambda: rast.visible_filter(means3D=ga["means3D"], scales=ga["scales"], rotations=ga["rotations"], cov3D_precomp=None)),
                 ("visible_filter_compact", lambda: rast.visible_filter_compact(means3D=ga["means3D"], scales=ga["scales"], rotations=ga["rotations"]))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    _lib.stage_timing(True)
    for _ in range(30):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    ms = _lib.stage_times()["visible_filter"]
    _lib.stage_timing(False)
    print(f"{name}: {ms*1e3:.1f} us  {44e6/(ms*1e-3)/1e9:.0f} GB/s algorithmic")

"""Probe: visible_filter / visible_filter_compact stage time over 1M anchors, three anchor sets:
  slab    anchors generated around the frame's slab only (57 % survive) — the bench's roofline scene
  cube    anchors over the whole cube depth in random order (18 % survive) — prefilter_voxel during training
  sorted  the same, z-sorted (the stream codec's layout; no index_range given)"""
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
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
sets = {"slab": synthetic_gaussians(1_000_000, geom, f0, f0, threshold=THRESHOLD, seed=4, device=dev),
        "cube": synthetic_gaussians(1_000_000, geom, 0, cfg["F"] - 1, threshold=THRESHOLD, seed=5, device=dev)}
order = torch.argsort(sets["cube"]["means3D"][:, 2], stable=True)
sets["sorted"] = {k: v[order].contiguous() for k, v in sets["cube"].items()}
for sname, ga in sets.items():
    for name, fn in (("visible_filter", lambda: rast.visible_filter(means3D=ga["means3D"], scales=ga["scales"], rotations=ga["rotations"], cov3D_precomp=None)),
                     ("visible_filter_compact", lambda: rast.visible_filter_compact(means3D=ga["means3D"], scales=ga["scales"], rotations=ga["rotations"])[1])):
        for _ in range(3):
            r = fn()
        torch.cuda.synchronize()
        _lib.stage_timing(True)
        for _ in range(30):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
        ms = _lib.stage_times()["visible_filter"]
        _lib.stage_timing(False)
        print(f"{sname:6s} {name:24s}: {ms*1e3:5.1f} us  ({100.0 * float((r > 0).float().mean()):.0f} % visible)")

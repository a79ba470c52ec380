"""Probe: stage times of BASELINE configs 2, 4, 5 (single view; config 4 forward only as in BASELINE.json)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gsvc_b200 import _lib
from gsvc_b200.frames import CONFIGS, CubeGeometry, synthetic_gaussians
from gsvc_b200.rasterizer import GaussianRasterizer
from gsvc_b200.sharding import GRAD_LAYOUT
from bench import settings_for, THRESHOLD

dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for cid in [int(a) for a in sys.argv[1:]] or [2, 4, 5]:
    cfg = CONFIGS[cid]
    geom = CubeGeometry(cfg["W"], cfg["H"], cfg["F"])
    f0 = cfg["F"] // 2
    g = synthetic_gaussians(cfg["P"], geom, f0, f0, threshold=THRESHOLD, seed=cid, device=dev)
    rast = GaussianRasterizer(raster_settings=settings_for(geom, f0, dev))
    dL = torch.randn((3, cfg["H"], cfg["W"]), device=dev)
    bwd = cid != 4

    def step():
        p = {k: g[k].detach().requires_grad_(bwd) for k, _ in GRAD_LAYOUT}
        m2d = torch.zeros_like(p["means3D"], requires_grad=bwd)
        color, radii, n = rast(means3D=p["means3D"], means2D=m2d, shs=None, colors_precomp=p["colors_precomp"],
                               opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)
        if bwd:
            torch.autograd.grad(color, [p[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)
        return n, radii

    for _ in range(4):
        n, radii = step()
    torch.cuda.synchronize()
    _lib.stage_timing(True)
    for _ in range(20):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    st = _lib.stage_times()
    _lib.stage_timing(False)
    T = ((cfg["W"] + 15) // 16) * ((cfg["H"] + 15) // 16)
    print(f"config {cid}: P={cfg['P']} {cfg['W']}x{cfg['H']} V={int((radii > 0).sum())} R={n} R/T={n / T:.0f}",
          {k: round(v * 1e3, 1) for k, v in st.items()}, "sum_us", round(sum(st.values()) * 1e3, 1))

"""Probe: per-stage device times (CUDA events around every kernel) of the bench step, eager launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for
from gsvc_b200 import _lib
from gsvc_b200.sharding import GRAD_LAYOUT
from gsvc_b200.views import ViewBatch, rasterize_views

views = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(1, dev)
H, W = cfg["H"], cfg["W"]
sets = [settings_for(geom, f0, dev), settings_for(geom, f0, dev, back=True)][:views]
batch = ViewBatch.toast(*sets) if views == 2 else ViewBatch(sets)
dL = torch.randn((batch.n_out, 3, H, W), device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def step():
    p = {k: g[k].detach().requires_grad_(True) for k, _ in GRAD_LAYOUT}
    img, radii, n = rasterize_views(batch, means3D=p["means3D"], opacities=p["opacities"],
                                    colors_precomp=p["colors_precomp"], scales=p["scales"], rotations=p["rotations"])
    torch.autograd.grad(img, [p[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)
    return n


for _ in range(5):
    n = step()
torch.cuda.synchronize()
_lib.stage_timing(True)
for _ in range(50):
    flush.zero_()
    step()
torch.cuda.synchronize()
st = _lib.stage_times()
print(f"views={views} R={n}", {k: round(v * 1e3, 1) for k, v in st.items()}, "sum_us", round(sum(st.values()) * 1e3, 1))

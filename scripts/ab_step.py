"""Quick A/B timing of the bench step (config 2, front + back view in one chain, forward + backward): CUDA-graph
replay, median / p10 of 300 steps with a 256 MiB L2 flush between steps, plus the per-stage times of an eager pass.
Usage: python scripts/ab_step.py [label]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for
from gsvc_b200 import _lib
from gsvc_b200.graphed import GraphedStep
from gsvc_b200.sharding import GRAD_LAYOUT, packed_backward
from gsvc_b200.views import ViewBatch, rasterize_views

dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(1, dev)
H, W = cfg["H"], cfg["W"]
toast = ViewBatch.toast(settings_for(geom, f0, dev), settings_for(geom, f0, dev, back=True))
dL = torch.randn((1, 3, H, W), generator=torch.Generator().manual_seed(100)).to(dev)
params = {k: v.clone() for k, v in g.items()}
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
step = GraphedStep(toast, params, dL)
fwd = GraphedStep(toast, params, None)


def run(fn, n=300):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return t[len(t) // 2], t[len(t) // 10]


m, p10 = run(step)
mf, pf = run(fwd)
buf = torch.empty((cfg["P"], 14), device=dev)


def eager():
    p = {k: params[k].detach().requires_grad_(True) for k, _ in GRAD_LAYOUT}
    img, _, _ = rasterize_views(toast, means3D=p["means3D"], opacities=p["opacities"], colors_precomp=p["colors_precomp"],
                                scales=p["scales"], rotations=p["rotations"])
    with packed_backward(buf):
        torch.autograd.grad(img, [p[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)


_lib.stage_timing(True)
for _ in range(5):
    eager()
torch.cuda.synchronize(); _lib.stage_times()
for _ in range(100):
    flush.zero_(); eager()
torch.cuda.synchronize()
st = _lib.stage_times()
_lib.stage_timing(False)
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'run'}: step median {m * 1e3:.1f} us (p10 {p10 * 1e3:.1f}) | fwd median {mf * 1e3:.1f} us | "
      + " ".join(f"{k}={v * 1e3:.1f}" for k, v in st.items()))

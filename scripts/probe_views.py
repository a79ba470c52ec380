"""Probe: per-view cost of the batched-view chain (config-2 scene) against the single-view step, graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for
from gsvc_b200.graphed import GraphedStep
from gsvc_b200.rasterizer import GaussianRasterizer
from gsvc_b200.views import ViewBatch

dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timed(step, n=60):
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    return ts[len(ts) // 2]


for frames in (1, 2, 4, 8):
    cfg, geom, f0, g = build_scene(frames, dev)
    H, W = cfg["H"], cfg["W"]
    params = {k: v.clone() for k, v in g.items()}
    single = GaussianRasterizer(raster_settings=settings_for(geom, f0, dev))
    dL1 = torch.randn((3, H, W), device=dev)
    t1 = timed(GraphedStep(single, params, dL1))
    t1f = timed(GraphedStep(single, params, None))
    pairs = [(settings_for(geom, f0 + i, dev), settings_for(geom, f0 + i, dev, back=True)) for i in range(frames)]
    toast = ViewBatch.toasts(pairs)
    dLt = torch.randn((frames, 3, H, W), device=dev)
    st = GraphedStep(toast, params, dLt)
    tt = timed(st)
    ttf = timed(GraphedStep(toast, params, None))
    torch.cuda.synchronize()
    V = 2 * frames
    print(f"window of {frames} frame(s), z-range widened accordingly: single view fwd+bwd {t1*1e3:.1f} us (fwd {t1f*1e3:.1f}); "
          f"toast batch of {V} views fwd+bwd {tt*1e3:.1f} us = {tt/V*1e3:.1f} us/view (fwd {ttf/V*1e3:.1f} us/view); "
          f"R_total={st.num_rendered()} cap_ok={st.capacity_ok()}")

"""Probe: forward-only frames replayed alternately on two streams (frame i+1's binning under frame i's blend)
against the same frames on one stream."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for
from gsvc_b200.graphed import GraphedStep
from gsvc_b200.views import ViewBatch

dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(1, dev)
params = {k: v.clone() for k, v in g.items()}
toast = ViewBatch.toast(settings_for(geom, f0, dev), settings_for(geom, f0, dev, back=True))
N = 400
for nstreams, prio in ((1, False), (3, False), (4, False), (6, False), (8, False), (12, False)):
    streams = [torch.cuda.Stream(dev, priority=(-1 if (prio and i % 2) else 0)) for i in range(nstreams)]
    steps = []
    for s in streams:
        with torch.cuda.stream(s):
            steps.append(GraphedStep(toast, params, None))
    torch.cuda.synchronize()
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(N):
            with torch.cuda.stream(streams[i % nstreams]):
                steps[i % nstreams]()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{nstreams} stream(s) priority={prio}: {dt / N * 1e6:.1f} us per frame = {N / dt:.0f} frames/s ({2 * N / dt:.0f} views/s)")

"""Probe: two frames' forward+backward (a reference iteration) as (a) one 4-view chain, (b) two toast chains one after
the other, (c) two toast chains on two streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_scene, settings_for
from gsvc_b200.graphed import GraphedStep
from gsvc_b200.views import ViewBatch

dev = torch.device("cuda:0")
cfg, geom, f0, g = build_scene(2, dev)
H, W = cfg["H"], cfg["W"]
params = {k: v.clone() for k, v in g.items()}
pairs = [(settings_for(geom, f0 + i, dev), settings_for(geom, f0 + i, dev, back=True)) for i in range(2)]
N = 200


def run(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e6


four = GraphedStep(ViewBatch.toasts(pairs), params, torch.randn((2, 3, H, W), device=dev))
print(f"(a) one 4-view chain: {run(four):.1f} us per iteration")
dL = torch.randn((1, 3, H, W), device=dev)
seq = [GraphedStep(ViewBatch.toast(*p), params, dL) for p in pairs]
print(f"(b) two toast chains, one stream: {run(lambda: (seq[0](), seq[1]())):.1f} us per iteration")
streams = [torch.cuda.Stream(dev) for _ in range(2)]
par = []
for s, p in zip(streams, pairs):
    with torch.cuda.stream(s):
        par.append(GraphedStep(ViewBatch.toast(*p), params, dL))
torch.cuda.synchronize()


def two_streams():
    cur = torch.cuda.current_stream(dev)
    for s, st in zip(streams, par):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            st()
    for s in streams:
        cur.wait_stream(s)


print(f"(c) two toast chains, two streams: {run(two_streams):.1f} us per iteration")

"""torchrun probe (N >= 2): e2e step time of HostStepPipeline with the two exchange paths (peer copies over symmetric
memory vs NCCL all-gather / reduce-scatter), same scene as bench.py's e2e leg.  Prints ms per step for each."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import build_scene, settings_for
from gsvc_b200.hostpipe import HostStepPipeline
from gsvc_b200.sharding import GRAD_LAYOUT
from gsvc_b200.views import ViewBatch

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
dist.init_process_group("nccl", device_id=dev)
cfg, geom, f0, g = build_scene(world, dev)
P, H, W = cfg["P"], cfg["H"], cfg["W"]
toast = ViewBatch.toast(settings_for(geom, f0 + rank, dev), settings_for(geom, f0 + rank, dev, back=True))
dL = torch.randn((1, 3, H, W), generator=torch.Generator().manual_seed(100 + rank)).to(dev)
host = torch.empty(14 * P, dtype=torch.float32).pin_memory()
off = 0
for k, w in GRAD_LAYOUT:
    host[off:off + w * P].copy_(g[k].reshape(-1).cpu()); off += w * P
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for mode in (sys.argv[2:] or ["peer", "nccl", "peer", "nccl"]):
    pipe = HostStepPipeline(P, dev, slots=2, sharded=True, peer_copies=(mode == "peer"))

    def steps(n):
        pipe.prefetch(host)
        for i in range(n):
            if i + 1 < n:
                pipe.prefetch(host)
            pipe.step(toast, dL)
    steps(8)
    torch.cuda.synchronize(dev); dist.barrier(device_ids=[dev.index])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(pipe.s_h2d)
    steps(STEPS)
    t_host = time.perf_counter() - t0
    e1.record(pipe.s_d2h)
    torch.cuda.synchronize(dev); dist.barrier(device_ids=[dev.index])
    ms = torch.tensor([e0.elapsed_time(e1) / STEPS], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{mode} (peer path active: {pipe.peer is not None}): {float(ms):.4f} ms per step, host loop {t_host / STEPS * 1e3:.4f} ms per step", flush=True)
    del pipe
dist.destroy_process_group()

"""The batched-view backward that CARRIES the all-reduce (gsvc_rast_backward_views_exchange) under torchrun, 2/4/8 GPUs:
every rank rasterizes its own frame (front + back view) of a shared Gaussian set; the packed [P,14] gradients summed over
the ranks by the backward's own launches are compared with the plain backward followed by an NCCL all-reduce, the ranks
are checked to be bit-identical, the step is replayed from a CUDA graph, and both variants are timed.
Prints one JSON line on rank 0.  Usage: torchrun ... scripts/check_fused_exchange.py [config=2|3] [mode=auto|multicast|peer]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import CONFIGS, THRESHOLD, build_scene, settings_for
from gsvc_b200.frames import synthetic_gaussians
from gsvc_b200.graphed import GraphedStep
from gsvc_b200.sharding import GRAD_LAYOUT, SwitchAllReduce, packed_backward
from gsvc_b200.views import ViewBatch, rasterize_views

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
which = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = sys.argv[2] if len(sys.argv) > 2 else "auto"
cfg, geom, f0, g = build_scene(world, dev)
H, W = cfg["H"], cfg["W"]
if which == 3:                                   # BASELINE config 3: 500k Gaussians, 8-frame window dealt to the ranks
    c3 = CONFIGS[3]
    frames = list(range(f0, f0 + c3["window"]))
    g = synthetic_gaussians(c3["P"], geom, frames[0], frames[-1], threshold=THRESHOLD, seed=3, device=dev)
    mine = [f for i, f in enumerate(frames) if i % world == rank]
else:
    mine = [f0 + rank]
P = int(g["means3D"].shape[0])
batch = ViewBatch.toasts([(settings_for(geom, f, dev), settings_for(geom, f, dev, back=True)) for f in mine])
dL = torch.randn((len(mine), 3, H, W), generator=torch.Generator().manual_seed(100 + rank)).to(dev)
ar = SwitchAllReduce(P * 14, dev, mode=mode)
fused_buf = ar.buffer().view(P, 14)
plain_buf = torch.empty((P, 14), device=dev)


def backward_into(buf, exchange=None):
    p = {k: g[k].detach().requires_grad_(True) for k, _ in GRAD_LAYOUT}
    img, _, _ = rasterize_views(batch, means3D=p["means3D"], opacities=p["opacities"], colors_precomp=p["colors_precomp"],
                                scales=p["scales"], rotations=p["rotations"])
    with packed_backward(buf, exchange=exchange):
        torch.autograd.grad(img, [p[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)


report = {"world": world, "config": which, "P": P, "views_per_rank": 2 * len(mine), "mode": ar.mode}
worst, same = 0.0, True
for rep in range(3):
    backward_into(plain_buf)
    dist.all_reduce(plain_buf)
    backward_into(fused_buf, exchange=ar)
    torch.cuda.synchronize()
    scale = plain_buf.abs().amax(dim=0).clamp_min(1e-30)
    worst = max(worst, ((fused_buf - plain_buf).abs() / scale).max().item())
    gathered = [torch.empty_like(fused_buf) for _ in range(world)]
    dist.all_gather(gathered, fused_buf.clone())
    same = same and all(torch.equal(gathered[0], x) for x in gathered)
report["max_err_vs_nccl_rel_to_column_max"] = worst
report["ranks_bit_identical"] = same
report["fused_launches"] = ar.fused_launches

# CUDA-graph replay of the whole step with the exchange inside
step = GraphedStep(batch, g, dL, exchange=ar)
for _ in range(3):
    step()
torch.cuda.synchronize()
report["graph_replay_err"] = ((fused_buf - plain_buf).abs() / scale).max().item()
plain_step = GraphedStep(batch, g, dL, packed=plain_buf)


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


sep = SwitchAllReduce(P * 14, dev, mode=mode)
sep_step = GraphedStep(batch, g, dL, packed=sep.buffer().view(P, 14))
report["ms_compute_only"] = timed(lambda: plain_step())
report["ms_then_nccl"] = timed(lambda: (plain_step(), dist.all_reduce(plain_buf)))
report["ms_then_switch_allreduce"] = timed(lambda: (sep_step(), sep.run()))
report["ms_fused"] = timed(lambda: step())
ok = worst <= 2e-5 and same and report["graph_replay_err"] <= 2e-5 and ar.fused_launches >= 3
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
report["ok"] = bool(flag.item())
if rank == 0:
    print(json.dumps(report))
dist.destroy_process_group()
sys.exit(0 if report["ok"] else 1)

"""torchrun check (N >= 2 GPUs): the row-sharded HostStepPipeline returns, on every rank, exactly its rows of the
gradient summed over the ranks' frames — against each rank computing all frames directly.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded_pipeline.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gsvc_b200.frames import CubeGeometry, synthetic_gaussians
from gsvc_b200.hostpipe import HostStepPipeline
from gsvc_b200.sharding import GRAD_LAYOUT
from gsvc_b200.views import ViewBatch, rasterize_views
from bench import settings_for

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
P, W, H, F = 8000 * world, 320, 192, 320
geom = CubeGeometry(W, H, F)
f0 = F // 2
g = synthetic_gaussians(P, geom, f0, f0 + world - 1, seed=5)
toasts = [ViewBatch.toast(settings_for(geom, f0 + r, dev), settings_for(geom, f0 + r, dev, back=True)) for r in range(world)]
dL = torch.randn((1, 3, H, W), generator=torch.Generator().manual_seed(3)).to(dev)   # same seed gradient on all ranks
host = torch.empty(14 * P, dtype=torch.float32).pin_memory()
off = 0
for k, w in GRAD_LAYOUT:
    host[off:off + w * P].copy_(g[k].reshape(-1)); off += w * P
pipe = HostStepPipeline(P, dev, slots=2, sharded=True)
assert pipe.world == world
outs = []
pipe.prefetch(host)
for i in range(5):
    if i < 4:
        pipe.prefetch(host)
    slot = pipe.step(toasts[rank], dL)
    outs.append(pipe.grads(slot)[pipe.r0:pipe.r1].clone())
# reference: this rank renders every rank's frame itself
p = {k: g[k].to(dev).requires_grad_(True) for k, _ in GRAD_LAYOUT}
total = None
for r in range(world):
    img, _, _ = rasterize_views(toasts[r], means3D=p["means3D"], opacities=p["opacities"], colors_precomp=p["colors_precomp"],
                                scales=p["scales"], rotations=p["rotations"])
    gr = torch.autograd.grad(img, [p[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)
    packed = torch.cat([x.reshape(P, -1) for x in gr], dim=1)
    total = packed if total is None else total + packed
ref = total[pipe.r0:pipe.r1].cpu()
for i, o in enumerate(outs):
    err = (o - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 2e-5, (rank, i, err)
assert pipe.graphs[0] is not None and pipe.capacity_ok(toasts[rank])
dist.barrier()
if rank == 0:
    how = "peer copies over symmetric memory (no collective kernels)" if pipe.peer is not None else \
        f"NCCL all-gather / reduce-scatter (peer mapping unavailable: {getattr(pipe, 'peer_error', 'disabled')})"
    print(f"sharded pipeline ok on {world} ranks: rows per rank {pipe.rows}, max rel err {err:.2e}; exchange: {how}")
dist.destroy_process_group()

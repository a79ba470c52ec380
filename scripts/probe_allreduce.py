"""All-reduce of the window's [P,14] fp32 gradient buffer (28 MB at 500k Gaussians), exposed (nothing overlaps it):
NCCL vs the symmetric-memory variants torch ships vs this repo's own kernels.  torchrun --nproc-per-node N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
numel = int(sys.argv[1]) if len(sys.argv) > 1 else 500000 * 14
t = symm.empty(numel, dtype=torch.float32, device=dev)
h = symm.rendezvous(t, dist.group.WORLD)
gname = dist.group.WORLD.group_name
if rank == 0:
    print("multicast_ptr", hex(getattr(h, "multicast_ptr", 0)), "world", world, "MB", numel * 4 / 1e6, flush=True)
plain = torch.empty(numel, dtype=torch.float32, device=dev)


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
    return ms.item() * 1e3


cases = {"nccl": lambda: dist.all_reduce(plain)}
for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
    op = getattr(torch.ops.symm_mem, name, None)
    if op is not None:
        cases["torch." + name] = (lambda op=op: op(t, "sum", gname))
try:
    from gsvc_b200.sharding import PeerAllReduce
    par = PeerAllReduce(numel, dev, slots=1)
    cases["PeerAllReduce"] = lambda: par.start(0).wait()
except Exception as e:
    if rank == 0:
        print("PeerAllReduce unavailable", e)
want = (torch.arange(numel, device=dev, dtype=torch.float32) % 1000) * world + sum(range(world))
for mode in ("multicast", "peer"):
    try:
        from gsvc_b200.sharding import SwitchAllReduce
        sar = SwitchAllReduce(numel, dev, mode=mode)
        sar.buffer().copy_(torch.arange(numel, device=dev, dtype=torch.float32) % 1000 + rank)
        sar.run()
        torch.cuda.synchronize()
        err = (sar.buffer() - want).abs().max().item()
        if rank == 0:
            print(f"SwitchAllReduce[{mode}] max err", err, "default n_ctas", sar.n_ctas, flush=True)
        for nc in (16, 32, 64, 128, 296):
            def f(nc=nc, sar=sar):
                sar.n_ctas = nc
                sar.run()
            cases[f"SwitchAllReduce[{mode}] n_ctas={nc}"] = f
    except Exception as e:
        if rank == 0:
            print(f"SwitchAllReduce[{mode}] unavailable:", type(e).__name__, e, flush=True)
for name, fn in cases.items():
    try:
        us = timed(fn)
        if rank == 0:
            print(f"{name:28s} {us:8.1f} us  algbw {numel * 4 / us / 1e3:7.1f} GB/s", flush=True)
    except Exception as e:
        if rank == 0:
            print(f"{name:28s} failed: {type(e).__name__}: {str(e)[:200]}", flush=True)
dist.destroy_process_group()

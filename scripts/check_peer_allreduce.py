"""torchrun check (N >= 2 GPUs): sharding.PeerAllReduce (copy-engine pulls over symmetric memory, no collective kernel)
equals NCCL's all_reduce, bit for bit on every rank (both sum the ranks' blocks in rank order per element... NCCL's order
is its own, so the comparison is to 1e-6 relative), over several rounds on both slots."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gsvc_b200.sharding import PeerAllReduce

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
P = 200_000
ar = PeerAllReduce(P * 14, dev, slots=2)
worst = 0.0
for it in range(8):
    slot = it & 1
    x = torch.randn(P * 14, generator=torch.Generator().manual_seed(100 * it + rank)).to(dev)
    ref = x.clone()
    dist.all_reduce(ref)
    ar.buffer(slot).copy_(x)
    ar.start(slot).wait()
    torch.cuda.synchronize()
    got = ar.buffer(slot)
    worst = max(worst, float((got - ref).abs().max() / ref.abs().max()))
    gathered = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, got.double().sum().view(1))
    assert all(torch.equal(gathered[0], g) for g in gathered), "ranks hold different sums"
assert worst <= 1e-6, worst
if rank == 0:
    print(f"PeerAllReduce ok on {world} ranks: max rel diff to NCCL {worst:.1e}, identical on every rank")
dist.destroy_process_group()

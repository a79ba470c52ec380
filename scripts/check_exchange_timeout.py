"""A rank that never arrives: rank 1 skips the collective call.  Rank 0's kernel must give up after its bounded wait
(~20 s), run to its end and report the timeout — not spin for good.  torchrun --nproc-per-node 2."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gsvc_b200.sharding import SwitchAllReduce

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ar = SwitchAllReduce(4 * 1000, dev, mode="peer")
ar.buffer().fill_(1.0)
ar.run(); torch.cuda.synchronize()                 # a normal exchange first: everybody arrives
ok = bool((ar.buffer() == world).all().item()) and ar.timeouts() == 0
dist.barrier()
t0 = time.time()
if rank == 0:
    ar.run()                                        # rank 1 never makes this call
    torch.cuda.synchronize()
    waited = time.time() - t0
    n = ar.timeouts(reset=True)
    print(json.dumps({"first_exchange_ok": ok, "gave_up_after_s": round(waited, 1), "timeouts_counted": n,
                      "ok": ok and n > 0 and 5.0 < waited < 60.0}))
dist.barrier()
dist.destroy_process_group()

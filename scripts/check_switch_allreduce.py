"""Correctness of gsvc_b200.sharding.SwitchAllReduce under torchrun (2, 4 or 8 GPUs of one node): both paths
(multicast through the NVSwitch, peer loads/stores) against a float64 sum of the ranks' inputs, ranks bit-identical,
ragged slice boundaries, repeated launches and a CUDA-graph replay.  Prints one JSON line on rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gsvc_b200.sharding import SwitchAllReduce

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
report = {"world": world, "cases": []}
ok = True
for mode in ("multicast", "peer"):
    for numel in (4, 4 * 1237, 14 * 50000, 16 * 200000):          # ragged: slices that do not divide evenly
        try:
            ar = SwitchAllReduce(numel, dev, mode=mode)
        except Exception as e:
            report["cases"].append({"mode": mode, "numel": numel, "skipped": f"{type(e).__name__}: {e}"[:160]})
            continue
        worst, same = 0.0, True
        for rep in range(3):
            gen = [torch.Generator().manual_seed(1000 * rep + q) for q in range(world)]
            xs = [torch.randn(numel, generator=g) * (10.0 ** (q % 3)) for q, g in enumerate(gen)]
            want = torch.stack([x.double() for x in xs]).sum(0)
            ar.buffer().copy_(xs[rank].to(dev))
            torch.cuda.synchronize(); dist.barrier()
            ar.run()
            torch.cuda.synchronize()
            got = ar.buffer().cpu()
            scale = torch.stack([x.abs() for x in xs]).sum(0).double().clamp_min(1e-30)
            worst = max(worst, ((got.double() - want).abs() / scale).max().item())
            gathered = [torch.empty_like(ar.buffer()) for _ in range(world)]
            dist.all_gather(gathered, ar.buffer().clone())
            same = same and all(torch.equal(gathered[0], g) for g in gathered)
        # CUDA-graph replay (the pad and the state words are back to zero after every launch)
        graph_ok = None
        if numel >= 14 * 50000:
            x = torch.full((numel,), float(rank + 1), device=dev)
            s = torch.cuda.Stream(dev)
            with torch.cuda.stream(s):
                ar.buffer().copy_(x); ar.run()                   # warm-up on the capture stream
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    ar.buffer().copy_(x)
                    ar.run()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            graph_ok = bool((ar.buffer() == world * (world + 1) / 2).all().item())
        case_ok = worst <= 2e-7 * world and same and graph_ok is not False
        ok = ok and case_ok
        report["cases"].append({"mode": mode, "numel": numel, "rel_err_vs_f64": worst, "ranks_bit_identical": same,
                                "graph_replay_ok": graph_ok, "ok": case_ok})
        del ar
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
report["ok"] = bool(flag.item())
if rank == 0:
    print(json.dumps(report))
dist.destroy_process_group()
sys.exit(0 if report["ok"] else 1)

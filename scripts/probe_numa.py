"""Host-to-device / device-to-host copy bandwidth from pinned buffers allocated on each NUMA node (the allocating
thread's CPU affinity decides where cudaHostAlloc's pages land)."""
import glob, os, re, subprocess, sys
import torch

def cpus_of(node):
    s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out

nodes = sorted(int(re.search(r"node(\d+)$", p).group(1)) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes", nodes, "affinity now", len(os.sched_getaffinity(0)), "cpus")
dev = torch.device("cuda:0")
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
    bid = q.lower().replace("00000000:", "0000:")
    print("gpu0 pci", q, "numa_node", open(f"/sys/bus/pci/devices/{bid}/numa_node").read().strip())
except Exception as e:
    print("pci lookup failed", e)
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
n = 64 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev)
all_cpus = os.sched_getaffinity(0)
for node in nodes + [None]:
    if node is not None:
        c = [x for x in cpus_of(node) if x in all_cpus]
        if not c:
            print("node", node, "no allowed cpus"); continue
        os.sched_setaffinity(0, c)
    else:
        os.sched_setaffinity(0, all_cpus)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        print(f"node {node} {name}: {20 * n / a.elapsed_time(b) / 1e6:.1f} GB/s")
    del h

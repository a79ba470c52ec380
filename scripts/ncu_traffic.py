"""Extract per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, per launch) and a few
other per-launch figures from an `ncu --set full` report into a small JSON that bench.py reads for
roofline.traffic.  Usage: python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/traffic.json"""
import csv, json, os, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsvc_b200.build import _digest

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
STAGE = {"sort_heavy_kernel": "sort_heavy", "preprocess_kernel<0>": "preprocess", "preprocess_kernel<1>": "visible_filter", "tile_scan_kernel": "tile_scan",
         "scatter_kernel": "scatter", "sort_tiles_kernel": "sort_tiles", "render_forward_kernel": "render_forward",
         "render_backward_kernel": "render_backward", "preprocess_backward_kernel": "preprocess_backward"}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
out = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    stage = next((v for k, v in STAGE.items() if k in name), None)
    if stage is None:
        continue
    def val(m):
        return float(r[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1)
    rec = out.setdefault(stage, {"launches": 0, "dram_bytes": 0.0, "duration_us": 0.0, "inst_executed": 0.0, "issue_active_pct": 0.0})
    rec["launches"] += 1
    rec["dram_bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    rec["duration_us"] += float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
    rec["inst_executed"] += float(r[ix["smsp__inst_executed.sum"]].replace(",", ""))
    rec["issue_active_pct"] += float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]].replace(",", ""))
for rec in out.values():
    n = rec.pop("launches")
    for k in list(rec):
        rec[k] = rec[k] / n
    rec["launches_averaged"] = n
# the digest of the kernel sources the capture was made from: bench.py refuses a profile whose digest is not the one
# of the library it is timing (a stale profile would silently quote the traffic of older kernels)
json.dump({"source": sys.argv[1], "csrc_digest": _digest(),
           "note": "per-launch means from one ncu --set full capture (cold cache, serialised)", "kernels": out},
          open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))

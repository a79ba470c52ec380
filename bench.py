#!/usr/bin/env python
"""bench.py — headline benchmark of the rasterizer hot path (BASELINE.json metric, config 2).

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: pure-PyTorch oracle on the host cores

A step = one pass of the hot path over one view of the synthetic UVG-shaped workload
(1920x1080, 200k Gaussians, BASELINE.json configs[1]): rasterizer forward + backward through the
reference-shaped GaussianRasterizer autograd surface.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

from gsvc_b200.frames import CONFIGS, CubeGeometry, synthetic_gaussians

WORKLOAD = "UVG-shaped synthetic 1920x1080 frame, 200k Gaussians, forward+backward (BASELINE.json configs[1])"
METRIC = "train_iters_per_s"
UNIT = "iters/s (1 iter = rasterizer fwd+bwd of one 1080p view, 200k Gaussians)"
THRESHOLD = 0.05  # /root/reference/cfgs/cfg_20240919.yaml:13


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(P, V, R, N, T):
    """Compulsory HBM bytes per launch of each kernel (SURVEY.md §8d stages S1..S5
    restated for this design's buffers)."""
    return {
        "preprocess": 56 * P + 4 * P + 56 * V,           # inputs, radii, 3 float4 + rect per visible Gaussian
        "tile_scan": 4 * T + 20 * T,                      # counts in; offsets, cursors, ranges out
        "scatter": 24 * V + 8 * R,                        # rect + depth in; (depth_key|id) out
        "sort_tiles": 8 * R + 8 * R,                      # composites in; point_list + depth_keys out
        "render_forward": 8 * T + 40 * R + 20 * N,        # ranges, id + 36 B features per instance, colour+T+n_contrib
        "render_backward": 8 * T + 40 * R + 20 * N + 36 * V,  # + dL/dC, final_T, n_contrib in; 9 floats per visible out
        "preprocess_backward": 4 * P + 40 * V + 36 * V + 68 * P,
        "visible_filter": 44 * P,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi's start-up stalls the driver for milliseconds: let it settle before anything is timed
            t0 = time.time()
            while not self.samples and time.time() - t0 < 10.0:
                time.sleep(0.05)
            self.samples.clear()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_scene(n_frames: int, device):
    cfg = CONFIGS[2]
    geom = CubeGeometry(cfg["W"], cfg["H"], cfg["F"])
    f0 = cfg["F"] // 2
    g = synthetic_gaussians(cfg["P"], geom, f0, f0 + n_frames - 1, threshold=THRESHOLD, seed=2, device=device)
    return cfg, geom, f0, g


def settings_for(geom, frame_id, device, back=False):
    """GaussianRasterizationSettings exactly as renderer.py:63-83 builds them."""
    from gsvc_b200.rasterizer import GaussianRasterizationSettings
    fr = geom.frame(frame_id)
    vm = fr.view_matrix_s if back else fr.view_matrix
    return GaussianRasterizationSettings(
        image_height=int(fr.image_height), image_width=int(fr.image_width), x_min=fr.x_min, y_min=fr.y_min,
        scale=fr.scale, threshold=THRESHOLD, bg=torch.zeros(3, dtype=torch.float32, device=device),
        scale_modifier=1.0, viewmatrix=vm.permute(1, 0).to(device), sh_degree=0, campos=fr.cam_pos,
        prefiltered=False, debug=False)


# --------------------------------------------------------------------------------------------------
# CPU arm: the pure-PyTorch oracle (there is no reference CPU implementation: renderer.py:37 is CUDA-only,
# and the reference's CUDA rasterizer source is not in /root/reference)
# --------------------------------------------------------------------------------------------------
def cpu_sample(geom, f0, g_cpu, tile_stride=16, threads=None):
    """One bounded sample of the config-2 step on the host: preprocess + binning of all 200k Gaussians,
    blend forward+backward (autograd) on every `tile_stride`-th tile; returns (seconds, extrapolated seconds)."""
    from oracle import torch_oracle
    from oracle.c_oracle import OracleSettings
    if threads:
        torch.set_num_threads(threads)
    fr = geom.frame(f0)
    st = OracleSettings(image_height=fr.image_height, image_width=fr.image_width, x_min=fr.x_min, y_min=fr.y_min,
                        scale=fr.scale, threshold=THRESHOLD, bg=np.zeros(3, np.float32),
                        viewmatrix=fr.view_matrix.permute(1, 0).numpy().copy(), campos=fr.cam_pos.numpy())
    T = ((fr.image_width + 15) // 16) * ((fr.image_height + 15) // 16)
    subset = np.arange(0, T, tile_stride)
    t0 = time.perf_counter()
    fwd = torch_oracle.forward(st, g_cpu["means3D"], g_cpu["opacities"], g_cpu["scales"], g_cpu["rotations"],
                               colors_precomp=g_cpu["colors_precomp"], requires_grad=True, tile_subset=subset)
    t1 = time.perf_counter()
    dL = torch.ones_like(fwd["color"])
    torch_oracle.backward(fwd, dL)
    t2 = time.perf_counter()
    # split: per-Gaussian + binning work is done in full; only the per-tile blend work scales with the tile count
    total = t2 - t0
    full = fwd["t_pre"] + (total - fwd["t_pre"]) * (T / len(subset))
    return total, full, len(subset), T


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, geom, f0, g = build_scene(1, "cpu")
    stride = 16
    times = []
    fulls = []
    for i in range(args.warmup + args.steps):
        total, full, ns, T = cpu_sample(geom, f0, g, tile_stride=stride, threads=cores)
        if i >= args.warmup:
            times.append(total)
            fulls.append(full)
    ms = 1000.0 * sum(times) / len(times)
    # a full step blends all T tiles: the per-tile part of the sample is scaled by T/ns, the per-Gaussian
    # part (done in full inside the sample) is not
    full_ms = 1000.0 * sum(fulls) / len(fulls)
    value = 1000.0 / full_ms
    sample = (f"config 2 scene on the host: preprocess+binning of all {cfg['P']} Gaussians, blend fwd+bwd (autograd) on "
              f"{ns} of {T} tiles (every {stride}th); blend time scaled x{T / ns:.1f} to a whole view")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": full_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sampled_ms_per_step": ms},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# product arm
# --------------------------------------------------------------------------------------------------
def run_product_arm(args, rank, local_rank, world):
    import torch.distributed as dist
    from gsvc_b200 import _lib
    from gsvc_b200.rasterizer import GaussianRasterizer
    from gsvc_b200.sharding import GRAD_LAYOUT, allreduce_grads, pack_grads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py product arm needs a CUDA device (no CPU fallback exists)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L = _lib.lib()

    cfg, geom, f0, g = build_scene(world, device)
    frame_id = f0 + rank                     # frame-sharded window: rank r renders frame f0 + r of the shared set
    rs = settings_for(geom, frame_id, device)
    rast = GaussianRasterizer(raster_settings=rs)
    P, W, H = cfg["P"], cfg["W"], cfg["H"]
    N, T = W * H, ((W + 15) // 16) * ((H + 15) // 16)
    params = {k: v.clone().requires_grad_(True) for k, v in g.items()}
    dL = torch.randn((3, H, W), generator=torch.Generator().manual_seed(100 + rank)).to(device)
    grad_buf = torch.empty((P, 14), dtype=torch.float32, device=device)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # 256 MiB > 126 MB L2

    def forward_only(p=params):
        with torch.no_grad():
            return rast(means3D=p["means3D"], means2D=p["means3D"], shs=None, colors_precomp=p["colors_precomp"],
                        opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)

    def train_step(p=params):
        means2D = torch.zeros_like(p["means3D"], requires_grad=True)
        color, radii, n = rast(means3D=p["means3D"], means2D=means2D, shs=None, colors_precomp=p["colors_precomp"],
                               opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"],
                               cov3D_precomp=None)
        grads = torch.autograd.grad(color, [p[k] for k, _ in GRAD_LAYOUT], grad_outputs=dL)
        if world > 1:
            pack_grads({k: gr for (k, _), gr in zip(GRAD_LAYOUT, grads)}, out=grad_buf)
            allreduce_grads(grad_buf)
        return color, radii, n, grads

    def sync_all():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    def timed(fn, steps, collect_stages=False):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed before each."""
        sync_all()
        total_ms, stage_ms = 0.0, {}
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize(device)
            total_ms += e0.elapsed_time(e1)
            if collect_stages:
                for k, v in _lib.stage_times().items():
                    stage_ms.setdefault(k, []).append(v)
        sync_all()
        t = torch.tensor([total_ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), stage_ms, out

    # ---- warm-up (also sets the instance-capacity hint so no step re-sizes its buffers)
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        out = train_step()
    torch.cuda.synchronize(device)
    num_rendered = out[2]
    V = int((out[1] > 0).sum().item())

    clocks = ClockSampler(local_rank)
    clocks.start()
    _lib.stage_timing(True)
    L.gsvc_rast_launch_count(1)
    total_ms, stage_ms, _ = timed(train_step, args.steps, collect_stages=True)
    launches = int(L.gsvc_rast_launch_count(1))
    fwd_ms, fwd_stage_ms, _ = timed(forward_only, args.steps, collect_stages=True)
    _lib.stage_timing(False)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    host_in = {k: v.detach().cpu().pin_memory() for k, v in g.items()}
    host_img = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
    host_grads = torch.empty((P, 14), dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * 4 for v in host_in.values())
    d2h = host_img.numel() * 4 + host_grads.numel() * 4

    def e2e_step():
        p = {k: v.to(device, non_blocking=True).requires_grad_(True) for k, v in host_in.items()}
        color, radii, n, grads = train_step(p)
        if world == 1:
            pack_grads({k: gr for (k, _), gr in zip(GRAD_LAYOUT, grads)}, out=grad_buf)
        host_img.copy_(color.detach(), non_blocking=True)
        host_grads.copy_(grad_buf, non_blocking=True)
        return n

    for _ in range(3):
        e2e_step()
    e2e_ms, _, _ = timed(e2e_step, args.steps)
    clk = clocks.stop()

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * 1000.0 / ms_per_step
        peak, peak_src = load_peaks()
        alg = algorithmic_bytes(P, V, num_rendered, N, T)
        stage_avg = {k: sum(v) / len(v) for k, v in stage_ms.items()}
        dom = max(stage_avg, key=stage_avg.get)
        achieved = alg[dom] / (stage_avg[dom] * 1e-3) / 1e9
        pairs = 256.0 * num_rendered
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "P": P, "V": V, "R": num_rendered, "N": N, "T": T,
                       "frames": f"frame {f0}+rank of F=600, front view", "l2": "256 MiB flush between timed steps",
                       "parallelism": f"frame-sharded x{world}" + (", NCCL fp32 sum all-reduce of [P,14] grads per step" if world > 1 else "")},
            "fwd_frames_per_s": world * 1000.0 * args.steps / fwd_ms,
            "fwd_ms_per_view": fwd_ms / args.steps,
            "stage_ms": {k: round(v, 5) for k, v in stage_avg.items()},
            "fwd_stage_ms": {k: round(sum(v) / len(v), 5) for k, v in fwd_stage_ms.items()},
            "pairs_per_s_fwd": pairs / (sum(fwd_stage_ms["render_forward"]) / len(fwd_stage_ms["render_forward"]) * 1e-3),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes": alg[dom],
                         "note": "blend kernels are FP32/issue-bound by construction (256 pixel-Gaussian pairs per 40 B instance); see DESIGN.md §4"},
            "e2e": {"value": world * 1000.0 * args.steps / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            g_cpu = {k: v.detach().cpu() for k, v in g.items()}
            stride = 16
            cpu_sample(geom, f0, g_cpu, tile_stride=64, threads=cores)  # warm-up
            total, full, ns, Tt = cpu_sample(geom, f0, g_cpu, tile_stride=stride, threads=cores)
            line["cpu_baseline"] = {
                "value": 1.0 / full, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": (f"pure-PyTorch oracle, same scene: preprocess+binning of all {P} Gaussians, blend fwd+bwd on "
                           f"{ns} of {Tt} tiles (every {stride}th) took {total:.2f} s; blend part scaled x{Tt / ns:.1f} "
                           f"to a whole view = {full:.1f} s")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="gsvc", choices=["gsvc", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    run_product_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

"""Per-tensor and per-Gaussian gradient error of the CUDA path against the referee oracle (tests/parity.py) on the
scenes the GPU suite uses, printed as a table — the calibration record behind ROW_FAIL_MAX / ROW_HARD.
Usage (GPU box): python scripts/parity_report.py [full]      (full: adds BASELINE config 2 at 1080p, two variants)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests import parity
from tests.scenes import make_scene, np_inputs, product_settings, stretch_gaussians
from gsvc_b200.rasterizer import GaussianRasterizer

dev = torch.device("cuda:0")


def report(name, scene):
    gi = np_inputs(scene["gaussians"])
    fo = parity.oracle_forward(scene["oracle_settings"], gi)
    g = {k: v.to(dev).requires_grad_(True) for k, v in scene["gaussians"].items()}
    m2d = torch.zeros_like(g["means3D"], requires_grad=True)
    color, radii, n = GaussianRasterizer(raster_settings=product_settings(scene, dev))(
        means3D=g["means3D"], means2D=m2d, shs=None, colors_precomp=g["colors_precomp"], opacities=g["opacities"],
        scales=g["scales"], rotations=g["rotations"], cov3D_precomp=None)
    assert n == fo["num_rendered"]
    err = np.abs(color.detach().cpu().numpy() - fo["color"])[:, ~fo["fragile"]].max(initial=0.0)
    dL = parity.masked_dL(fo, torch.randn(color.shape, generator=torch.Generator().manual_seed(5)))
    color.backward(torch.as_tensor(dL).to(dev))
    go = parity.oracle_backward(fo, dL)
    vis = fo["radii"] > 0
    print(f"{name}: R={n} visible={int(vis.sum())} fragile_px={fo['fragile'].mean():.2e} fwd_err={err:.2e}")
    got = {k: g[k].grad for k in parity.GRAD_NAMES}
    got["means2D"] = m2d.grad
    for k, v in got.items():
        a = v.detach().cpu().numpy().reshape(len(vis), -1).astype(np.float64)[vis]
        b = np.asarray(go[k], np.float64).reshape(len(vis), -1)[vis]
        s = parity.grad_stats(a, b)
        ratio = np.abs(a - b).max(axis=1) / (parity.ROW_RTOL * np.abs(b).max(axis=1) + parity.ROW_ATOL * s["scale"] + 1e-30)
        print(f"   {k:15s} rel={s['rel']:.2e} rows_missing={s['row_fail']:.2e} worst_row={s['row_worst']:.2f} "
              f"p99={np.quantile(ratio, 0.99):.2f} p999={np.quantile(ratio, 0.999):.2f}")


if __name__ == "__main__":
    for back in (False, True):
        report(f"20k 256x256 back={back}", make_scene(P=20000, W=256, H=256, F=256, back=back, seed=1))
        report(f"20k 256x256 3% needles back={back}", make_scene(P=20000, W=256, H=256, F=256, back=back, seed=1, needle_mix=0.03))
    for st in (2.0, 4.0, 8.0, 16.0):
        report(f"axes {st * st:.0f}:1", make_scene(P=int(20000 / st ** 2), W=256, H=256, F=256, seed=3, stretch=st))
    big = make_scene(P=600, W=320, H=200, F=320, seed=41, bg=(0.7, 0.2, 0.4), scale_modifier=1.7)
    big["gaussians"]["scales"] *= 6.0
    big["gaussians"]["opacities"] *= 0.3
    report("600 screen-filling splats", big)
    if "full" in sys.argv[1:]:
        for mix in (None, 0.05):
            report(f"config 2 (1080p, 200k) needle_mix={mix}", make_scene(P=200000, W=1920, H=1080, F=600, seed=2, bg=(0, 0, 0), needle_mix=mix))

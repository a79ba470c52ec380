"""Timing of the binning stages on a skewed scene: one tile with ~33 000 instances in an
otherwise ordinary 256x256 frame, against the same frame without the hot tile.  Prints the per-stage times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gsvc_b200 import _lib
from gsvc_b200.rasterizer import GaussianRasterizer
from tests.scenes import make_scene, product_settings
from tests.test_gpu_parity import _skewed_scene

dev = torch.device("cuda:0")
for name, scene in (("ordinary 70k Gaussians, 256x256", make_scene(P=70000, W=256, H=256, F=256, seed=77)),
                    ("one hot tile (50k of the 70k Gaussians on it)", _skewed_scene()),
                    ("one hot tile, a third of it at one depth", _skewed_scene(ties=True))):
    g = {k: v.to(dev) for k, v in scene["gaussians"].items()}
    rast = GaussianRasterizer(raster_settings=product_settings(scene, dev))
    kw = dict(means3D=g["means3D"], means2D=g["means3D"], shs=None, colors_precomp=g["colors_precomp"],
              opacities=g["opacities"], scales=g["scales"], rotations=g["rotations"], cov3D_precomp=None)
    with torch.no_grad():
        for _ in range(3):
            _, _, n = rast(**kw)
        torch.cuda.synchronize()
        _lib.stage_timing(True)
        for _ in range(20):
            rast(**kw)
        torch.cuda.synchronize()
        st = _lib.stage_times()
        _lib.stage_timing(False)
    print(f"{name}: R={n} " + " ".join(f"{k}={v * 1e3:.1f}us" for k, v in st.items()))

"""Print forward / gradient errors of the CUDA path vs the C oracle on a few seeded scenes (margin check)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import c_oracle
from tests.scenes import make_scene, np_inputs, product_settings
from gsvc_b200.rasterizer import GaussianRasterizer

dev = torch.device("cuda:0")
for cfg in (dict(P=20000, W=256, H=256, F=256, seed=1), dict(P=20000, W=256, H=256, F=256, seed=1, back=True),
            dict(P=60000, W=640, H=360, F=600, seed=4)):
    scene = make_scene(**cfg)
    gi = np_inputs(scene["gaussians"])
    fo = c_oracle.forward(scene["oracle_settings"], gi["means3D"], gi["opacities"], gi["scales"], gi["rotations"],
                          colors_precomp=gi["colors_precomp"])
    g = {k: v.to(dev).requires_grad_(True) for k, v in scene["gaussians"].items()}
    m2d = torch.zeros_like(g["means3D"], requires_grad=True)
    rast = GaussianRasterizer(raster_settings=product_settings(scene, dev))
    color, radii, n = rast(means3D=g["means3D"], means2D=m2d, shs=None, colors_precomp=g["colors_precomp"],
                           opacities=g["opacities"], scales=g["scales"], rotations=g["rotations"], cov3D_precomp=None)
    dL = torch.randn(color.shape, generator=torch.Generator().manual_seed(5))
    color.backward(dL.to(dev))
    go = c_oracle.backward(fo, dL.numpy())
    solid = ~fo["fragile"]
    err = np.abs(color.detach().cpu().numpy() - fo["color"])
    print(cfg, "R", n, "fwd max err solid", err[:, solid].max(), "fragile px", int(fo["fragile"].sum()),
          "fragile max", err[:, ~solid].max(initial=0))
    ok = ~go["touched_fragile"]
    for k in ("means3D", "scales", "rotations", "opacities", "colors_precomp"):
        a = g[k].grad.cpu().numpy().reshape(len(ok), -1)[ok]
        b = go[k].reshape(len(ok), -1)[ok]
        print(f"   {k:15s} rel {np.abs(a - b).max() / np.abs(b).max():.3e}")
    a = m2d.grad.cpu().numpy()[ok]
    print(f"   {'means2D':15s} rel {np.abs(a - go['means2D'][ok]).max() / np.abs(go['means2D']).max():.3e}")

"""Shared scene builders for the tests: one seeded synthetic scene → oracle settings + product settings."""
from __future__ import annotations

import numpy as np
import torch

from gsvc_b200.frames import CubeGeometry, synthetic_gaussians
from oracle.c_oracle import OracleSettings


def make_scene(P=3000, W=96, H=64, F=96, frame=None, seed=7, back=False, bg=(0.1, 0.2, 0.3), threshold=0.05,
               window=1, scale_modifier=1.0):
    geom = CubeGeometry(W, H, F)
    frame = F // 2 if frame is None else frame
    fr = geom.frame(frame)
    g = synthetic_gaussians(P, geom, frame, frame + window - 1, threshold=threshold, seed=seed)
    vm = fr.view_matrix_s if back else fr.view_matrix          # stored transposed, like the reference
    V = vm.permute(1, 0)                                        # renderer.py:77
    st = OracleSettings(image_height=H, image_width=W, x_min=fr.x_min, y_min=fr.y_min, scale=fr.scale,
                        threshold=threshold, bg=np.asarray(bg, np.float32), scale_modifier=scale_modifier,
                        viewmatrix=V.numpy().copy(), sh_degree=0, campos=fr.cam_pos.numpy())
    return dict(geom=geom, frame=fr, gaussians=g, oracle_settings=st, view_matrix_stored=vm)


def product_settings(scene, device, debug=False, sh_degree=0):
    """Build GaussianRasterizationSettings exactly as renderer.py:63-83 does."""
    from gsvc_b200.rasterizer import GaussianRasterizationSettings
    st = scene["oracle_settings"]
    fr = scene["frame"]
    return GaussianRasterizationSettings(
        image_height=int(st.image_height), image_width=int(st.image_width), x_min=fr.x_min, y_min=fr.y_min,
        scale=fr.scale, threshold=st.threshold, bg=torch.tensor(st.bg, dtype=torch.float32, device=device),
        scale_modifier=st.scale_modifier, viewmatrix=scene["view_matrix_stored"].permute(1, 0).to(device),
        sh_degree=sh_degree, campos=fr.cam_pos, prefiltered=False, debug=debug)


def np_inputs(g):
    return {k: v.numpy() for k, v in g.items()}

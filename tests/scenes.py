"""Shared scene builders for the tests: one seeded synthetic scene → oracle settings + product settings."""
from __future__ import annotations

import numpy as np
import torch

from gsvc_b200.frames import CubeGeometry, synthetic_gaussians
from oracle.c_oracle import OracleSettings


def stretch_gaussians(g, stretch=1.0, needle_mix=None, seed=0):
    """Elongate Gaussians in place: axis 0 times `stretch`, axis 1 divided by it (axis ratio stretch^2 : 1 on top of
    the generator's own spread).  `needle_mix` = fraction of the Gaussians that get a random stretch from
    {2, 4, 8, 16} (axis ratios 4:1 .. 256:1) instead — an ordinary scene with needles in it."""
    P = g["scales"].shape[0]
    if needle_mix:
        rng = np.random.default_rng(seed)
        sel = torch.as_tensor(rng.random(P) < needle_mix)
        st = torch.as_tensor(rng.choice([2.0, 4.0, 8.0, 16.0], P)).float()
        g["scales"][sel, 0] *= st[sel]
        g["scales"][sel, 1] /= st[sel]
    elif stretch != 1.0:
        g["scales"][:, 0] *= stretch
        g["scales"][:, 1] /= stretch
    return g


def make_scene(P=3000, W=96, H=64, F=96, frame=None, seed=7, back=False, bg=(0.1, 0.2, 0.3), threshold=0.05,
               window=1, scale_modifier=1.0, stretch=1.0, needle_mix=None):
    geom = CubeGeometry(W, H, F)
    frame = F // 2 if frame is None else frame
    fr = geom.frame(frame)
    g = synthetic_gaussians(P, geom, frame, frame + window - 1, threshold=threshold, seed=seed)
    stretch_gaussians(g, stretch, needle_mix, seed)
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


def golden_scene(cfg):
    """Scene of a golden fixture from its stored cfg: make_scene keys plus, for the v2 fixtures, `sh_degree` /
    `sh_seed` (seeded SH coefficients [P,16,3] replace colors_precomp).  Returns (scene, numpy inputs, sh_degree)."""
    cfg = dict(cfg)
    deg, sh_seed = cfg.pop("sh_degree", None), cfg.pop("sh_seed", None)
    scene = make_scene(**cfg)
    gi = np_inputs(scene["gaussians"])
    if deg is not None:
        P = gi["means3D"].shape[0]
        gi["shs"] = (torch.randn(P, 16, 3, generator=torch.Generator().manual_seed(sh_seed)) * 0.4).numpy()
        gi.pop("colors_precomp")
        scene["oracle_settings"].sh_degree = int(deg)
    return scene, gi, deg

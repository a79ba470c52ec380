"""Mint the golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference holds NO golden vectors for this path and its rasterizer cannot be built or imported
(un-vendored dependency), so these fixtures are produced by the repo's own C oracle
(oracle/splat_oracle.c) on seeded synthetic scenes; the independent PyTorch oracle is checked
against the same files in tests/test_oracle.py.  They pin the oracle against drift — they do not
pin it against the (absent) reference: parity stays "unpinned".

Two generations: the three round-1 fixtures (default-mode oracle; small, no fragile pixels) are kept as minted, and
`*_v2.npz` (round 2) come from the REFEREE oracle with the seed gradient zeroed on its fragile pixels, so that every
visible Gaussian's gradient is part of the fixture (tests/parity.py): a full-width 1080p strip at ~190 instances per
tile, an ordinary scene with 3 % needles (axis ratios 4:1 .. 256:1), and degree-3 SH colours — each with a non-empty
fragile set.  `python tests/golden/make_golden.py` re-mints only the v2 files; `--all` the old ones too.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle  # noqa: E402
from tests.scenes import make_scene, np_inputs  # noqa: E402

CASES = {
    "front_96x64": dict(P=1500, W=96, H=64, F=96, seed=101, back=False),
    "back_96x64": dict(P=1500, W=96, H=64, F=96, seed=101, back=True),
    "ragged_50x34": dict(P=800, W=50, H=34, F=64, seed=202, back=False, bg=(0.3, 0.1, 0.6)),
}


CASES_V2 = {
    "strip_1920x64_dense_v2": dict(P=30000, W=1920, H=64, F=600, seed=303, back=False),
    "needles_256x256_v2": dict(P=20000, W=256, H=256, F=256, seed=404, back=True, needle_mix=0.03, bg=(0.2, 0.4, 0.1)),
    "sh3_160x96_v2": dict(P=6000, W=160, H=96, F=128, seed=505, back=False, sh_degree=3, sh_seed=506),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mint_v2(out_dir):
    from tests import parity
    from tests.scenes import golden_scene
    for name, cfg in CASES_V2.items():
        scene, gi, deg = golden_scene(cfg)
        fo = parity.oracle_forward(scene["oracle_settings"], gi)
        H, W = cfg["H"], cfg["W"]
        dL = parity.masked_dL(fo, torch.randn(3, H, W, generator=torch.Generator().manual_seed(cfg["seed"] + 1)))
        go = parity.oracle_backward(fo, dL)
        colour = {"g_shs": go["shs"].astype(np.float32)} if deg is not None else {"g_colors": go["colors_precomp"].astype(np.float32)}
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            cfg=np.array(repr(cfg)), referee=np.int32(1), color=fo["color"], radii=fo["radii"],
            num_rendered=np.int64(fo["num_rendered"]), n_contrib=fo["n_contrib"].astype(np.uint16),
            fragile=np.packbits(fo["fragile"]), keys_sha=np.array(sha(fo["bin"]["keys"])),
            point_list_sha=np.array(sha(fo["bin"]["point_list"])), ranges_sha=np.array(sha(fo["bin"]["ranges"])),
            dL_seed=np.int64(cfg["seed"] + 1),
            g_means3D=go["means3D"].astype(np.float32), g_scales=go["scales"].astype(np.float32),
            g_rotations=go["rotations"].astype(np.float32), g_opacities=go["opacities"].astype(np.float32),
            g_means2D=go["means2D"].astype(np.float32), **colour)
        print(name, "R", fo["num_rendered"], "visible", int((fo["radii"] > 0).sum()), "fragile pixels",
              int(fo["fragile"].sum()), "max tile list", int(np.diff(fo["bin"]["ranges"].astype(np.int64), axis=1).max()))


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    mint_v2(out_dir)
    if "--all" not in sys.argv:
        return
    for name, cfg in CASES.items():
        scene = make_scene(**cfg)
        gi = np_inputs(scene["gaussians"])
        fo = c_oracle.forward(scene["oracle_settings"], gi["means3D"], gi["opacities"], gi["scales"], gi["rotations"],
                              colors_precomp=gi["colors_precomp"])
        H, W = cfg["H"], cfg["W"]
        dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(cfg["seed"] + 1)).numpy()
        go = c_oracle.backward(fo, dL)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            cfg=np.array(repr(cfg)), color=fo["color"], radii=fo["radii"], num_rendered=np.int64(fo["num_rendered"]),
            n_contrib=fo["n_contrib"].astype(np.uint16), fragile=fo["fragile"],
            keys_sha=np.array(sha(fo["bin"]["keys"])), point_list_sha=np.array(sha(fo["bin"]["point_list"])),
            ranges_sha=np.array(sha(fo["bin"]["ranges"])), dL=dL.astype(np.float32),
            g_means3D=go["means3D"].astype(np.float32), g_scales=go["scales"].astype(np.float32),
            g_rotations=go["rotations"].astype(np.float32), g_opacities=go["opacities"].astype(np.float32),
            g_colors=go["colors_precomp"].astype(np.float32), g_means2D=go["means2D"].astype(np.float32),
            touched_fragile=go["touched_fragile"])
        print(name, "R", fo["num_rendered"], "visible", int((fo["radii"] > 0).sum()))


if __name__ == "__main__":
    main()

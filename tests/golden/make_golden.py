"""Mint the golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference holds NO golden vectors for this path and its rasterizer cannot be built or imported
(un-vendored dependency), so these fixtures are produced by the repo's own C oracle
(oracle/splat_oracle.c) on seeded synthetic scenes; the independent PyTorch oracle is checked
against the same files in tests/test_oracle.py.  They pin the oracle against drift — they do not
pin it against the (absent) reference: parity stays "unpinned".
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


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
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

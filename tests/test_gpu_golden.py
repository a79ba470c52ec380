"""The CUDA path against the COMMITTED golden fixtures (tests/golden/*.npz, minted by tests/golden/make_golden.py
from the C oracle; the reference itself holds no vectors for this path — SURVEY.md §8c).  Nothing of the oracle runs
here: the stored image, radii, instance count, hashes of the sorted keys / point list / tile ranges, contributor
counts and gradients are compared with what libgsvc_rast.so produces for the scene the fixture names."""
import ast
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from tests import parity
from tests.scenes import golden_scene, make_scene, product_settings

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = (("means3D", "g_means3D"), ("scales", "g_scales"), ("rotations", "g_rotations"), ("opacities", "g_opacities"),
         ("colors_precomp", "g_colors"))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


V1 = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not p.endswith("_v2.npz"))
V2 = sorted(glob.glob(os.path.join(GOLDEN, "*_v2.npz")))


@pytest.mark.parametrize("path", V2, ids=os.path.basename)
def test_cuda_path_reproduces_the_v2_fixture(cuda_device, path):
    """Round-2 fixtures (referee oracle, seed gradient zeroed on its fragile pixels): a 1080p-wide dense strip, an
    ordinary scene with needles, degree-3 SH colours, each with fragile pixels.  Integer stages by hash, image 1e-5
    off the stored fragile set, and EVERY visible Gaussian's gradient at the per-tensor and per-Gaussian bars."""
    from gsvc_b200.rasterizer import GaussianRasterizer, RasterState
    z = np.load(path)
    cfg = ast.literal_eval(str(z["cfg"]))
    scene, gi, deg = golden_scene(cfg)
    H, W = cfg["H"], cfg["W"]
    frag = np.unpackbits(z["fragile"])[:H * W].reshape(H, W).astype(bool)
    assert frag.any() and int(z["referee"]) == 1
    rs = product_settings(scene, cuda_device, sh_degree=deg or 0)
    g = {k: torch.as_tensor(v).to(cuda_device) for k, v in gi.items()}
    colour = dict(shs=g["shs"]) if deg is not None else dict(colors_precomp=g["colors_precomp"])
    st = RasterState(rs, g["means3D"], g["opacities"], scales=g["scales"], rotations=g["rotations"], **colour)
    keys, pl, ranges = st.export_keys()
    fT, n_contrib = st.export_image()
    assert st.num_rendered == int(z["num_rendered"])
    np.testing.assert_array_equal(st.radii.cpu().numpy(), z["radii"])
    assert _sha(keys.cpu().numpy().view(np.uint64)) == str(z["keys_sha"])
    assert _sha(pl.cpu().numpy().view(np.uint32)) == str(z["point_list_sha"])
    assert _sha(ranges.cpu().numpy().view(np.uint32)) == str(z["ranges_sha"])
    np.testing.assert_array_equal(n_contrib.cpu().numpy().astype(np.uint16)[~frag], z["n_contrib"][~frag])
    names = [k for k in ("means3D", "scales", "rotations", "opacities", "colors_precomp", "shs") if k in g]
    p = {k: g[k].clone().requires_grad_(True) for k in names}
    m2d = torch.zeros_like(p["means3D"], requires_grad=True)
    color, radii, n = GaussianRasterizer(raster_settings=rs)(
        means3D=p["means3D"], means2D=m2d, shs=p.get("shs"), colors_precomp=p.get("colors_precomp"),
        opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)
    assert n == int(z["num_rendered"])
    fo = dict(color=z["color"], fragile=frag, radii=z["radii"])
    parity.check_forward(fo, color, max_fragile=1e-2)
    dL = parity.masked_dL(fo, torch.randn(3, H, W, generator=torch.Generator().manual_seed(int(z["dL_seed"]))))
    grads = torch.autograd.grad(color, [p[k] for k in names] + [m2d], grad_outputs=torch.as_tensor(dL).to(cuda_device))
    stored = dict(means3D="g_means3D", scales="g_scales", rotations="g_rotations", opacities="g_opacities",
                  colors_precomp="g_colors", shs="g_shs", means2D="g_means2D")
    got = dict(zip(names + ["means2D"], grads))
    parity.check_grads(fo, {k: z[stored[k]] for k in got}, got)


@pytest.mark.parametrize("path", V1, ids=os.path.basename)
def test_cuda_path_reproduces_the_golden_fixture(cuda_device, path):
    from gsvc_b200.rasterizer import GaussianRasterizer, RasterState
    z = np.load(path)
    scene = make_scene(**ast.literal_eval(str(z["cfg"])))
    rs = product_settings(scene, cuda_device)
    g = {k: v.to(cuda_device) for k, v in scene["gaussians"].items()}
    # integer stages: exact (hashes of the exported arrays are the fixture's)
    st = RasterState(rs, g["means3D"], g["opacities"], colors_precomp=g["colors_precomp"], scales=g["scales"],
                     rotations=g["rotations"])
    keys, pl, ranges = st.export_keys()
    _, n_contrib = st.export_image()
    assert st.num_rendered == int(z["num_rendered"])
    np.testing.assert_array_equal(st.radii.cpu().numpy(), z["radii"])
    assert _sha(keys.cpu().numpy().view(np.uint64)) == str(z["keys_sha"])
    assert _sha(pl.cpu().numpy().view(np.uint32)) == str(z["point_list_sha"])
    assert _sha(ranges.cpu().numpy().view(np.uint32)) == str(z["ranges_sha"])
    solid = ~z["fragile"]
    np.testing.assert_array_equal(n_contrib.cpu().numpy().astype(np.uint16)[solid], z["n_contrib"][solid])
    # the drop-in call: image 1e-5, gradients 1e-4 of the tensor's largest entry
    p = {k: g[k].clone().requires_grad_(True) for k, _ in NAMES}
    m2d = torch.zeros_like(p["means3D"], requires_grad=True)
    color, radii, n = GaussianRasterizer(raster_settings=rs)(
        means3D=p["means3D"], means2D=m2d, shs=None, colors_precomp=p["colors_precomp"], opacities=p["opacities"],
        scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)
    assert n == int(z["num_rendered"])
    np.testing.assert_array_equal(radii.cpu().numpy(), z["radii"])
    assert np.abs(color.detach().cpu().numpy() - z["color"])[:, solid].max() <= 1e-5
    grads = torch.autograd.grad(color, [p[k] for k, _ in NAMES] + [m2d], grad_outputs=torch.as_tensor(z["dL"]).to(cuda_device))
    ok = ~z["touched_fragile"]
    P = ok.shape[0]
    for (k, zk), gr in zip(NAMES + (("means2D", "g_means2D"),), grads):
        a = gr.cpu().numpy().reshape(P, -1)[ok]
        b = z[zk].astype(np.float64).reshape(P, -1)[ok]
        assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max() + 1e-12, k

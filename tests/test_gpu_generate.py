"""SURVEY.md §8f row f2, second half: the fused generator epilogue (gsvc_b200.generate, C-ABI gsvc_gen_epilogue_*)
against the reference's PyTorch expression (guassian.py:147-153, 251-293 restated in generate.reference_epilogue):
selection mask and row order exact, values to fp32 rounding, gradients of every input to 1e-6 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(N, K, n_vis_frac, device, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    anchor = torch.rand(N, 3, generator=g)
    grid_offsets = 0.3 * r(N, K, 3)
    grid_scaling = torch.exp(0.3 * r(N, 6) - 3.0)
    masks = (torch.rand(N, K, 1, generator=g) > 0.3).float()
    vis = torch.nonzero(torch.rand(N, generator=g) < n_vis_frac).flatten().to(torch.int32)
    n = int(vis.numel())
    nop = torch.tanh(r(n, K))                      # the opacity MLP ends in tanh: about half are <= 0
    color = torch.sigmoid(r(n, K * 3))
    scale_rot = r(n, K * 7)
    noff = 0.1 * r(n, K * 3)
    lo, hi = torch.full((1, 3), 0.05), torch.full((1, 3), 0.95)      # some positions get clamped
    t = [anchor, grid_offsets, grid_scaling, masks, nop, color, scale_rot, noff]
    return [x.to(device) for x in t], vis.to(device), lo.to(device), hi.to(device)


@pytest.mark.parametrize("N,K,frac", [(1000, 10, 0.3), (257, 1, 1.0), (5000, 5, 0.0), (100_000, 10, 0.15), (31, 3, 0.5)])
@pytest.mark.parametrize("gathered", [False, True])
def test_epilogue_matches_the_pytorch_expression(cuda_device, N, K, frac, gathered):
    from gsvc_b200.generate import neural_gaussians_epilogue, reference_epilogue
    t, vis, lo, hi = _inputs(N, K, frac, cuda_device, seed=N + K)
    if gathered:                                     # the caller gathered the per-anchor rows itself
        idx = vis.long()
        t[:4] = [x[idx].contiguous() for x in t[:4]]
        vis = None
    a = [x.clone().requires_grad_(True) for x in t]
    b = [x.clone().requires_grad_(True) for x in t]
    got = neural_gaussians_epilogue(a[0], a[1], a[2], a[3], vis, a[4], a[5], a[6], a[7], lo, hi)
    ref = reference_epilogue(b[0], b[1], b[2], b[3], vis, b[4], b[5], b[6], b[7], lo, hi, K)
    assert torch.equal(got.mask, ref.mask)                                   # same Gaussians, same order
    assert got.xyz.shape == ref.xyz.shape and got.opacity.shape == ref.opacity.shape
    assert torch.equal(got.neural_opacity, ref.neural_opacity) and torch.equal(got.opacity, ref.opacity)
    assert torch.equal(got.color, ref.color)
    for name in ("xyz", "scaling", "rot"):
        x, y = getattr(got, name), getattr(ref, name)
        assert (x - y).abs().max().item() <= 2e-7 * max(1.0, y.abs().max().item()) if y.numel() else True, name
    if ref.xyz.numel() == 0:
        return
    gen = torch.Generator().manual_seed(3)
    w = [torch.randn(x.shape, generator=gen).to(cuda_device) for x in ref[:6]]
    loss = lambda o: sum((getattr(o, f) * wi).sum() for f, wi in zip(("xyz", "color", "opacity", "scaling", "rot", "neural_opacity"), w))
    ga = torch.autograd.grad(loss(got), a)
    gb = torch.autograd.grad(loss(ref), b)
    for name, x, y in zip(("anchor", "grid_offsets", "grid_scaling", "masks", "neural_opacity", "color", "scale_rot", "neural_offset"), ga, gb):
        assert x.shape == y.shape, name
        assert (x - y).abs().max().item() <= 2e-6 * max(1e-6, y.abs().max().item()), name


def test_epilogue_feeds_the_rasterizer(cuda_device):
    """The whole f2 chain without a host synchronisation: visible_filter_compact -> (stub MLP) -> fused epilogue ->
    rasterizer call, and a backward through all of it down to the anchors."""
    from gsvc_b200.generate import neural_gaussians_epilogue
    from gsvc_b200.rasterizer import GaussianRasterizer
    from tests.scenes import make_scene, product_settings
    scene = make_scene(P=4000, W=128, H=96, F=128, seed=3)
    N, K = 4000, 4
    g = {k: v.to(cuda_device) for k, v in scene["gaussians"].items()}
    anchor = g["means3D"].clone().requires_grad_(True)
    scaling6 = torch.cat([g["scales"], g["scales"]], dim=1).requires_grad_(True)
    offsets = (0.5 * torch.randn(N, K, 3, generator=torch.Generator().manual_seed(1))).to(cuda_device).requires_grad_(True)
    masks = torch.ones(N, K, 1, device=cuda_device)
    rast = GaussianRasterizer(raster_settings=product_settings(scene, cuda_device))
    idx, _ = rast.visible_filter_compact(means3D=anchor.detach(), scales=scaling6.detach()[:, :3], rotations=g["rotations"])
    n = int(idx.numel())
    w = torch.randn(3, 7 * K + 7 * K, generator=torch.Generator().manual_seed(2)).to(cuda_device).requires_grad_(True)
    feat = anchor.index_select(0, idx.long())
    h = torch.tanh(feat @ w)                                                       # the stub "MLP"
    nop, col, sr, noff = h[:, :K], torch.sigmoid(h[:, K:4 * K]), h[:, 4 * K:11 * K], 0.1 * h[:, 11 * K:14 * K]
    gg = neural_gaussians_epilogue(anchor, offsets, scaling6, masks, idx, nop, col, sr, noff,
                                   torch.full((3,), -10.0), torch.full((3,), 10.0))
    assert gg.xyz.shape[0] == int(gg.mask.sum()) > 0 and gg.neural_opacity.shape == (n * K, 1)
    means2D = torch.zeros_like(gg.xyz, requires_grad=True)
    image, radii, num = rast(means3D=gg.xyz, means2D=means2D, shs=None, colors_precomp=gg.color, opacities=gg.opacity,
                             scales=gg.scaling, rotations=gg.rot, cov3D_precomp=None)
    image.square().sum().backward()
    for t in (anchor, scaling6, offsets, w):
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0
    assert (anchor.grad[torch.ones(N, dtype=torch.bool, device=cuda_device).index_fill(0, idx.long(), False)] == 0).all()

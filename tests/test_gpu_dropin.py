"""The reference's call sequence, restated against the import-path shim (no GSVC model: duck-typed stand-ins):
prefilter_voxel() -> visibility mask over anchors -> per-anchor Gaussians of the visible anchors -> rasterizer ->
RenderResults fields -> loss.backward() -> the densification statistic.  Every rasterizer-facing line mirrors
/root/reference/ortho_gaussian_renderer/preprocess.py:58-108 and renderer.py:63-119 (keyword names, the non-contiguous
`scales[:, :3]` / `view_matrix.permute(1, 0)` arguments, CPU campos, `zeros_like(...) + 0` screen-space points)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import c_oracle
from oracle.c_oracle import OracleSettings
from gsvc_b200.frames import CubeGeometry, synthetic_gaussians

pytestmark = pytest.mark.gpu


class _Pipe:
    debug = False
    compute_cov3D_python = False


class _Cfg:
    threshold = 0.05
    sh_degree = 0


class _Anchors:
    """Stand-in for GaussianModel: anchors with a 6-column scaling (guassian.py:285: columns 3: are the Gaussians'
    own scale factors) and raw quaternions normalised on access (scene/gaussian_model.py get_rotation)."""

    def __init__(self, g, device):
        self.model_config = _Cfg()
        self._anchor = g["means3D"].to(device).requires_grad_(True)
        self._scaling = torch.cat([g["scales"], g["scales"] * 0.5], dim=1).to(device).requires_grad_(True)   # [N,6]
        self._rotation = (g["rotations"] * 1.7).to(device).requires_grad_(True)                               # un-normalised
        self._color = g["colors_precomp"].to(device).requires_grad_(True)
        self._opacity = g["opacities"].to(device).requires_grad_(True)

    get_anchor = property(lambda s: s._anchor)
    get_scaling = property(lambda s: s._scaling)
    get_rotation = property(lambda s: F.normalize(s._rotation))


def prefilter_voxel(frame, pc, pipe, bg_color, scaling_modifier=1.0):
    from diff_gaussian_rasterization.cuda_ortho_gaussian_rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    raster_settings = GaussianRasterizationSettings(                                  # preprocess.py:58-79
        image_height=int(frame.image_height), image_width=int(frame.image_width), x_min=frame.x_min, y_min=frame.y_min,
        scale=frame.scale, threshold=pc.model_config.threshold, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=frame.view_matrix.permute(1, 0).cuda(), sh_degree=pc.model_config.sh_degree, campos=frame.cam_pos,
        prefiltered=False, debug=pipe.debug)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)                  # preprocess.py:81
    means3D = pc.get_anchor
    scales, rotations, cov3D_precomp = pc.get_scaling, pc.get_rotation, None
    radii_pure = rasterizer.visible_filter(means3D=means3D, scales=scales[:, :3], rotations=rotations,
                                           cov3D_precomp=cov3D_precomp)            # preprocess.py:99-104
    return radii_pure > 0                                                            # preprocess.py:108


def render(frame, pc, pipe, bg_color, scaling_modifier=1.0, retain_grad=True):
    from diff_gaussian_rasterization.cuda_ortho_gaussian_rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    visible_mask = prefilter_voxel(frame, pc, pipe, bg_color)                          # renderer.py:29
    # generate_neural_gaussians stand-in: one Gaussian per visible anchor (boolean-mask gathers, guassian.py:147-153)
    xyz, color, opacity = pc.get_anchor[visible_mask], pc._color[visible_mask], pc._opacity[visible_mask]
    scaling, rot = pc.get_scaling[visible_mask][:, 3:] * 2.0, pc.get_rotation[visible_mask]
    screenspace_points = torch.zeros_like(xyz, dtype=pc.get_anchor.dtype, requires_grad=True, device="cuda") + 0
    if retain_grad:
        screenspace_points.retain_grad()                                               # renderer.py:37-42
    raster_settings = GaussianRasterizationSettings(                                   # renderer.py:63-83
        image_height=int(frame.image_height), image_width=int(frame.image_width), x_min=frame.x_min, y_min=frame.y_min,
        scale=frame.scale, threshold=pc.model_config.threshold, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=frame.view_matrix.permute(1, 0).cuda(), sh_degree=pc.model_config.sh_degree, campos=frame.cam_pos,
        prefiltered=False, debug=pipe.debug)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)                   # renderer.py:85
    rendered_image, radii, num_rendered = rasterizer(                                  # renderer.py:90-98
        means3D=xyz, means2D=screenspace_points, shs=None, colors_precomp=color, opacities=opacity, scales=scaling,
        rotations=rot, cov3D_precomp=None)
    return dict(rendered_image=rendered_image, viewspace_points=screenspace_points, visibility_filter=radii > 0,
                visible_mask=visible_mask, radii=radii, active_gaussains=(radii > 0).sum(), num_rendered=num_rendered,
                inputs=(xyz, color, opacity, scaling, rot))


def test_reference_call_sequence_front_and_back(cuda_device):
    torch.cuda.set_device(cuda_device)
    W, H, Fr, N = 320, 192, 320, 30000
    geom = CubeGeometry(W, H, Fr)
    frame = geom.frame(Fr // 2)
    g = synthetic_gaussians(N, geom, Fr // 2, Fr // 2 + 1, seed=71)
    pc = _Anchors(g, cuda_device)
    background = torch.zeros(3, dtype=torch.float32, device="cuda")                    # pipeline/train.py:328
    out_f = render(frame, pc, _Pipe(), background)
    frame.view_matrix, frame.view_matrix_s = frame.view_matrix_s.cuda(), frame.view_matrix.cuda()   # train.py:358
    out_b = render(frame, pc, _Pipe(), background)
    image = (out_f["rendered_image"] + torch.flip(out_b["rendered_image"], dims=(-1,))) / 2          # train.py:366-375
    assert image.shape == (3, H, W) and isinstance(out_f["num_rendered"], int) and out_f["radii"].dtype == torch.int32
    target = torch.rand((3, H, W), generator=torch.Generator().manual_seed(1)).cuda()
    loss = (image - target).abs().mean()
    loss.backward()                                                                    # train.py:462
    for p in (pc._anchor, pc._scaling, pc._rotation, pc._color, pc._opacity):
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0
    assert (pc._scaling.grad[:, :3] == 0).all()         # the filter's scales[:, :3] are not differentiated through
    # the densification statistic of scene/gaussian_model.py:1311-1314
    vp, vis = out_f["viewspace_points"], out_f["visibility_filter"]
    grad_norm = torch.norm(vp.grad[vis, :2], dim=-1, keepdim=True)
    assert grad_norm.shape == (int(vis.sum()), 1) and torch.isfinite(grad_norm).all() and (vp.grad[:, 2] == 0).all()
    assert int(out_f["active_gaussains"]) == int(vis.sum()) > 0
    # the forward of the front view against the oracle, on the arguments the call sequence actually produced
    xyz, color, opacity, scaling, rot = [t.detach().cpu().numpy() for t in out_f["inputs"]]
    fr = geom.frame(Fr // 2)
    st = OracleSettings(image_height=H, image_width=W, x_min=fr.x_min, y_min=fr.y_min, scale=fr.scale, threshold=0.05,
                        bg=np.zeros(3, np.float32), viewmatrix=fr.view_matrix.permute(1, 0).numpy().copy(),
                        campos=fr.cam_pos.numpy())
    fo = c_oracle.forward(st, xyz, opacity, scaling, rot, colors_precomp=color, referee=True)
    assert fo["num_rendered"] == out_f["num_rendered"]
    np.testing.assert_array_equal(out_f["radii"].cpu().numpy(), fo["radii"])
    err = np.abs(out_f["rendered_image"].detach().cpu().numpy() - fo["color"])[:, ~fo["fragile"]]
    assert err.max() <= 1e-5
    # and the anchor-level mask against the oracle's visible_filter on the [:, :3] scales
    ref_mask = c_oracle.visible_filter(st, g["means3D"].numpy(), g["scales"].numpy(),
                                       F.normalize(g["rotations"] * 1.7).numpy()) > 0
    np.testing.assert_array_equal(out_f["visible_mask"].cpu().numpy(), ref_mask)

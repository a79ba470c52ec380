"""Maximum-size probe (run by hand on the GPU; recorded in profiles/): tens of millions of Gaussians over the whole
cube depth against one 4K frame — 64-bit addressing of every per-Gaussian array, multi-CTA scans, the compaction at
scale — checked through size-independent properties:

  * visible_filter radii == the forward's radii; fused compaction == nonzero(radii > 0);
  * the frame rendered from ALL P Gaussians == the frame rendered from only the compacted visible subset, bit for
    bit (culled rows contribute nothing and the emission order of the rest is unchanged), same instance count;
  * gradients of the full call: exactly zero on culled rows, equal to the subset call's on the visible rows
    (up to the order of the float atomics);
  * the toast (front + back view in one chain, 2P virtual Gaussians) == the two single calls composed.

Usage: python scripts/probe_large.py [P_millions=40] [W=3840] [H=2160]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gsvc_b200.frames import CubeGeometry
from gsvc_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
from gsvc_b200.views import render_toast

NAMES = ("means3D", "colors_precomp", "opacities", "scales", "rotations")


def scene(P, geom, thr, dev, seed=1):
    """The SURVEY.md §8d distributions, generated on the device (the CPU generator of frames.synthetic_gaussians
    would spend minutes here): x, y over 1.1x the image extent, z over the WHOLE cube depth."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    fr0, fr1 = geom.frame(0), geom.frame(geom.frames - 1)
    u = torch.rand((P, 3), generator=gen, device=dev)
    w, h = geom.width / fr0.scale, geom.height / fr0.scale
    x = fr0.x_min + (u[:, 0] * 1.1 - 0.05) * w
    y = fr0.y_min + (u[:, 1] * 1.1 - 0.05) * h
    z = (fr0.z - 1.5 * thr) + u[:, 2] * ((fr1.z + 1.5 * thr) - (fr0.z - 1.5 * thr))
    sig = torch.exp(0.6931 + 0.6 * torch.randn((P, 3), generator=gen, device=dev)).clamp_(0.3, 30.0) / fr0.scale
    q = torch.nn.functional.normalize(torch.randn((P, 4), generator=gen, device=dev), dim=1)
    return dict(means3D=torch.stack([x, y, z], 1).contiguous(), scales=sig.contiguous(), rotations=q.contiguous(),
                opacities=(0.05 + 0.95 * torch.rand((P, 1), generator=gen, device=dev)),
                colors_precomp=torch.rand((P, 3), generator=gen, device=dev))


def settings(geom, fid, thr, dev, back):
    fr = geom.frame(fid)
    vm = fr.view_matrix_s if back else fr.view_matrix
    return GaussianRasterizationSettings(
        image_height=geom.height, image_width=geom.width, x_min=fr.x_min, y_min=fr.y_min, scale=fr.scale,
        threshold=thr, bg=torch.tensor([0.1, 0.2, 0.3], device=dev), scale_modifier=1.0,
        viewmatrix=vm.permute(1, 0).to(dev), sh_degree=0, campos=fr.cam_pos, prefiltered=False, debug=False)


def call(rs, g, dL=None):
    p = {k: g[k].detach().requires_grad_(dL is not None) for k in NAMES}
    color, radii, n = GaussianRasterizer(raster_settings=rs)(
        means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"], requires_grad=dL is not None), shs=None,
        colors_precomp=p["colors_precomp"], opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"],
        cov3D_precomp=None)
    grads = torch.autograd.grad(color, [p[k] for k in NAMES], grad_outputs=dL) if dL is not None else None
    return color.detach(), radii, n, grads


def main():
    P = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 40_000_000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
    dev, thr = torch.device("cuda:0"), 0.05
    geom = CubeGeometry(W, H, 600)
    t0 = time.time()
    g = scene(P, geom, thr, dev)
    torch.cuda.synchronize()
    print(f"P = {P:,} Gaussians on {W}x{H} ({sum(v.numel() * 4 for v in g.values()) / 2**30:.1f} GiB of parameters), "
          f"generated in {time.time() - t0:.1f} s", flush=True)
    for fid, back in ((300, False), (599, True)):
        rs = settings(geom, fid, thr, dev, back)
        rast = GaussianRasterizer(raster_settings=rs)
        vf = rast.visible_filter(means3D=g["means3D"], scales=g["scales"], rotations=g["rotations"], cov3D_precomp=None)
        idx, vr = rast.visible_filter_compact(means3D=g["means3D"], scales=g["scales"], rotations=g["rotations"])
        assert torch.equal(vr, vf) and torch.equal(idx.long(), torch.nonzero(vf > 0).flatten())
        dL = torch.randn((3, H, W), generator=torch.Generator(device=dev).manual_seed(fid), device=dev)
        torch.cuda.synchronize(); t1 = time.time()
        color, radii, n, grads = call(rs, g, dL)
        torch.cuda.synchronize(); t2 = time.time()
        assert torch.equal(radii, vf), "visible_filter radii != forward radii"
        sub = {k: g[k].index_select(0, idx.long()) for k in NAMES}
        color_s, radii_s, n_s, grads_s = call(rs, sub, dL)
        assert n == n_s and torch.equal(color, color_s), "full frame != frame of the compacted subset"
        assert torch.equal(radii_s, vf.index_select(0, idx.long()))
        culled = vf == 0
        worst = 0.0
        for k, a, b in zip(NAMES, grads, grads_s):
            assert not bool(a[culled].count_nonzero()), f"{k}: non-zero gradient on a culled row"
            assert bool(torch.isfinite(a).all())
            rel = float((a.index_select(0, idx.long()) - b).abs().max() / b.abs().max())
            worst = max(worst, rel)
            assert rel <= 4e-5, (k, rel)
        print(f"frame {fid} {'back' if back else 'front'}: visible {idx.numel():,} ({100.0 * idx.numel() / P:.1f} %), "
              f"num_rendered {n:,}; full == compacted subset (image bit-exact, gradients within {worst:.1e}); "
              f"first full forward+backward call {1e3 * (t2 - t1):.1f} ms; peak memory "
              f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    # toast over the full set: 2P virtual Gaussians in one chain
    f, b = settings(geom, 300, thr, dev, False), settings(geom, 300, thr, dev, True)
    with torch.no_grad():
        img, radii2, n2 = render_toast(f, b, means3D=g["means3D"], opacities=g["opacities"],
                                       colors_precomp=g["colors_precomp"], scales=g["scales"], rotations=g["rotations"])
        cf, rf, nf, _ = call(f, g)
        cb, rb, nb, _ = call(b, g)
    ref = 0.5 * cf + 0.5 * torch.flip(cb, dims=[-1])
    assert n2 == nf + nb and torch.equal(radii2[0], rf) and torch.equal(radii2[1], rb)
    err = float((img - ref).abs().max())
    assert err <= 2e-7, err
    print(f"toast of frame 300 over all {P:,} Gaussians: num_rendered {n2:,} = {nf:,} + {nb:,}, max |toast - composed| "
          f"{err:.1e}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    print("large probe ok")


if __name__ == "__main__":
    main()

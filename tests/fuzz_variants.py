"""Randomised parity sweep of the input VARIANTS the main sweep (tests/fuzz_parity.py) leaves fixed: spherical-harmonic
colours of every degree (forward clamp mask and SH backward), a precomputed 3D covariance instead of scales/rotations,
the toast composition (front + flip_x(back))/2 in one batched chain, and visible_filter with its fused
compaction and slab index range (radii and indices exact), each against the C oracle: images 1e-5 off
fragile pixels, gradients 1e-4 relative, radii / instance counts exact.
Usage: python tests/fuzz_variants.py [n_cases=30] [seed=0]"""
import dataclasses
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import c_oracle
from tests import parity
from tests.scenes import make_scene, np_inputs, product_settings
from gsvc_b200.rasterizer import GaussianRasterizer
from gsvc_b200.views import render_toast


DIAG = False


def _rel(a, b):
    return float(np.abs(a - b).max(initial=0.0) / (np.abs(b).max(initial=0.0) + 1e-30))


def _check_grads(case, what, names, grads, go, vis, P, worst, bw_abs):
    """Every visible Gaussian is compared (the seed gradient is zeroed on the fragile pixels, tests/parity.py).
    Per tensor: 1e-4 of its largest entry; when that fails (tiny scenes: a gradient that is a sum of signed per-pixel
    terms can cancel to far below any one term and no other Gaussian sets the scale), the cancellation-aware bar
    instead — 1e-5 of the same gradient taken with |dL| (no cancellation over pixels) — counted and reported.  The
    per-Gaussian criterion of tests/parity.py is recorded beside it."""
    go_abs = None
    for k, gr in zip(names, grads):
        full = gr.cpu().numpy().reshape(P, -1)
        assert not full[~vis].any(), (case, what, k, "a culled Gaussian received a gradient")
        a = full[vis].astype(np.float64)
        b = go[k].reshape(P, -1)[vis]
        if not vis.any():
            continue
        r = _rel(a, b)
        if r > 1e-4:
            go_abs = bw_abs() if go_abs is None else go_abs
            r_abs = float(np.abs(a - b).max() / (np.abs(go_abs[k].reshape(P, -1)[vis]).max() + 1e-30))
            if r_abs <= 1e-5:
                worst["cancelled"] = worst.get("cancelled", 0) + 1
                continue
        if r > 1e-4 and DIAG:
            i = np.unravel_index(np.abs(a - b).argmax(), a.shape)
            gid = np.nonzero(vis)[0][i[0]]
            print(f"DIAG {what} {k}: rel {r:.3e} at gaussian {gid} comp {i[1]}: got {a[i]:.9e} ref {b[i]:.9e} "
                  f"max|ref| {np.abs(b).max():.3e}; row got {a[i[0]]} ref {b[i[0]]}")
            continue
        assert r <= 1e-4, (case, what, k, r)
        st = parity.grad_stats(a, b)
        worst["row_fail"] = max(worst.get("row_fail", 0.0), st["row_fail"] if st["n"] >= 200 else 0.0)
        worst["rows"] = worst.get("rows", 0) + st["n"]
        if r > worst["grad"]:
            worst["grad"], worst["grad_at"] = r, f"case {case} {what} {k} P={P} max|ref|={np.abs(b).max():.2e}"


def _cov3d(gi, sm):
    """A valid covariance (xx, xy, xz, yy, yz, zz) per Gaussian; both sides take it as an input, so float64 is fine."""
    q = gi["rotations"].astype(np.float64)
    r, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                  2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                  2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = R * (sm * gi["scales"].astype(np.float64))[:, None, :]
    S = M @ M.transpose(0, 2, 1)
    return np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1).astype(np.float32)


def _filter_case(case, rng, W, H, F, sm, back, dev):
    """visible_filter, its fused compaction and the slab index range against the oracle's radii: anchors over a
    random span of frames, counts around the warp / CTA boundaries of the compaction, z-sorted for the range."""
    from gsvc_b200.frames import CubeGeometry, slab_index_range, synthetic_gaussians, z_interval_table
    from gsvc_b200.rasterizer import GaussianRasterizationSettings
    P = int(rng.choice([1, 31, 32, 33, 255, 256, 257, 1000, 4097, 20000, 70001]))
    thr = float(rng.choice([0.02, 0.05, 0.2]))
    geom = CubeGeometry(W, H, F)
    f0 = int(rng.integers(F)); f1 = int(min(F - 1, f0 + rng.integers(1, 40))); fid = int(rng.integers(f0, f1 + 1))
    gs = synthetic_gaussians(P, geom, f0, f1, threshold=thr, seed=int(rng.integers(1 << 30)))
    order = torch.argsort(gs["means3D"][:, 2], stable=True)
    gs = {k: v[order].contiguous() for k, v in gs.items()}
    fr = geom.frame(fid)
    vm = fr.view_matrix_s if back else fr.view_matrix
    st = c_oracle.OracleSettings(image_height=H, image_width=W, x_min=fr.x_min, y_min=fr.y_min, scale=fr.scale,
                                 threshold=thr, scale_modifier=sm, viewmatrix=vm.permute(1, 0).numpy().copy())
    want = c_oracle.visible_filter(st, gs["means3D"].numpy(), gs["scales"].numpy(), gs["rotations"].numpy())
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, x_min=fr.x_min, y_min=fr.y_min, scale=fr.scale, threshold=thr,
        bg=torch.zeros(3, device=dev), scale_modifier=sm, viewmatrix=vm.permute(1, 0).to(dev), sh_degree=0,
        campos=fr.cam_pos, prefiltered=False, debug=False)
    rast = GaussianRasterizer(raster_settings=rs)
    g = {k: v.to(dev) for k, v in gs.items()}
    want_t = torch.as_tensor(want, device=dev)
    want_idx = torch.nonzero(want_t > 0).flatten()
    got = rast.visible_filter(means3D=g["means3D"], scales=g["scales"], rotations=g["rotations"], cov3D_precomp=None)
    assert torch.equal(got, want_t), (case, "filter")
    idx, radii = rast.visible_filter_compact(means3D=g["means3D"], scales=g["scales"], rotations=g["rotations"])
    assert torch.equal(radii, want_t) and torch.equal(idx.long(), want_idx), (case, "compact")
    idx2, none = rast.visible_filter_compact(means3D=g["means3D"], scales=g["scales"], rotations=g["rotations"],
                                             want_radii=False)
    assert none is None and torch.equal(idx2, idx), (case, "compact without radii")
    lo, hi = slab_index_range(z_interval_table(gs["means3D"][:, 2]), fr.z, thr)
    for k in ("means3D", "scales", "rotations"):
        g[k][:lo] = float("nan"); g[k][hi:] = float("nan")
    part = rast.visible_filter(means3D=g["means3D"], scales=g["scales"], rotations=g["rotations"], cov3D_precomp=None,
                               index_range=(lo, hi))
    assert torch.equal(part, want_t), (case, "range", lo, hi)
    idx3, radii3 = rast.visible_filter_compact(means3D=g["means3D"], scales=g["scales"], rotations=g["rotations"],
                                               index_range=(lo, hi))
    assert torch.equal(radii3, want_t) and torch.equal(idx3.long(), want_idx), (case, "range compact", lo, hi)


def run(n_cases=30, seed=0, verbose=True, only_case=None):
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    worst = dict(fwd=0.0, grad=0.0)
    t0 = time.time()
    for case in range(n_cases):
        W = int(rng.choice([16, 33, 100, 160, 250]))
        H = int(rng.choice([16, 40, 96, 130]))
        F = int(rng.choice([64, 128]))
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        P = int(min(40000, max(1, float(rng.choice([4, 20, 100, 400, 900])) * tiles / 3.0)))
        bg = tuple(float(x) for x in rng.random(3)) if rng.integers(2) else (0.0, 0.0, 0.0)
        sm = float(rng.choice([1.0, 0.5, 2.0]))
        sseed = int(rng.integers(1 << 30))
        variant = ("sh", "cov", "toast", "filter", "toast_sh")[case % 5]
        deg = int(rng.integers(4))
        back = bool(rng.integers(2))
        if only_case is not None and case != only_case:
            continue
        if variant == "filter":
            _filter_case(case, np.random.default_rng(sseed), W, H, F, sm, back, dev)
            if verbose:
                print(f"case {case:3d} ok: filter {W}x{H}", flush=True)
            continue
        dL = torch.randn((3, H, W), generator=torch.Generator().manual_seed(case))
        scene = make_scene(P=P, W=W, H=H, F=F, seed=sseed, back=back, bg=bg, scale_modifier=sm)
        gi = np_inputs(scene["gaussians"])
        g = {k: v.to(dev) for k, v in scene["gaussians"].items()}
        if variant == "sh":
            M = (deg + 1) ** 2
            shs = (torch.randn(P, M, 3, generator=torch.Generator().manual_seed(case + 1)) * 0.4)
            st = dataclasses.replace(scene["oracle_settings"], sh_degree=deg)
            fo = c_oracle.forward(st, gi["means3D"], gi["opacities"], gi["scales"], gi["rotations"], shs=shs.numpy(),
                                  referee=True)
            bw = lambda d: c_oracle.backward(fo, d)
            names = ("means3D", "shs", "opacities", "scales", "rotations")
            p = {k: g[k].clone().requires_grad_(True) for k in names if k != "shs"}
            p["shs"] = shs.to(dev).requires_grad_(True)
            m2d = torch.zeros_like(p["means3D"], requires_grad=True)
            color, radii, n = GaussianRasterizer(raster_settings=product_settings(scene, dev, sh_degree=deg))(
                means3D=p["means3D"], means2D=m2d, shs=p["shs"], colors_precomp=None, opacities=p["opacities"],
                scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)
        elif variant == "cov":
            cov = _cov3d(gi, sm)
            fo = c_oracle.forward(scene["oracle_settings"], gi["means3D"], gi["opacities"], cov3D_precomp=cov,
                                  colors_precomp=gi["colors_precomp"], referee=True)
            bw = lambda d: c_oracle.backward(fo, d)
            names = ("means3D", "colors_precomp", "opacities", "cov3D_precomp")
            p = {k: g[k].clone().requires_grad_(True) for k in names if k != "cov3D_precomp"}
            p["cov3D_precomp"] = torch.as_tensor(cov, device=dev).requires_grad_(True)
            m2d = torch.zeros_like(p["means3D"], requires_grad=True)
            color, radii, n = GaussianRasterizer(raster_settings=product_settings(scene, dev))(
                means3D=p["means3D"], means2D=m2d, shs=None, colors_precomp=p["colors_precomp"],
                opacities=p["opacities"], scales=None, rotations=None, cov3D_precomp=p["cov3D_precomp"])
        else:
            use_sh = variant == "toast_sh"
            sc_f = make_scene(P=P, W=W, H=H, F=F, seed=sseed, back=False, bg=bg, scale_modifier=sm)
            sc_b = make_scene(P=P, W=W, H=H, F=F, seed=sseed, back=True, bg=bg, scale_modifier=sm)
            col = {"colors_precomp": gi["colors_precomp"]}
            names = ("means3D", "colors_precomp", "opacities", "scales", "rotations")
            if use_sh:      # SH colours: the per-view camera position enters the colour (and its gradient to means3D)
                shs = (torch.randn(P, (deg + 1) ** 2, 3, generator=torch.Generator().manual_seed(case + 1)) * 0.4)
                col = {"shs": shs.numpy()}
                names = ("means3D", "shs", "opacities", "scales", "rotations")
                for sc in (sc_f, sc_b):
                    sc["oracle_settings"] = dataclasses.replace(sc["oracle_settings"], sh_degree=deg)
            fos = [c_oracle.forward(s["oracle_settings"], gi["means3D"], gi["opacities"], gi["scales"], gi["rotations"],
                                    referee=True, **col) for s in (sc_f, sc_b)]

            def bw(d):
                gos = [c_oracle.backward(fos[0], 0.5 * d),
                       c_oracle.backward(fos[1], np.ascontiguousarray(0.5 * d[:, :, ::-1]))]
                return {k: gos[0][k] + gos[1][k] for k in names}
            p = {k: g[k].clone().requires_grad_(True) for k in names if k != "shs"}
            if use_sh:
                p["shs"] = shs.to(dev).requires_grad_(True)
            ckw = {"shs": p["shs"]} if use_sh else {"colors_precomp": p["colors_precomp"]}
            d_ = deg if use_sh else 0
            color, radii2, n = render_toast(product_settings(sc_f, dev, sh_degree=d_), product_settings(sc_b, dev, sh_degree=d_),
                                            means3D=p["means3D"], opacities=p["opacities"], scales=p["scales"],
                                            rotations=p["rotations"], **ckw)
            assert np.array_equal(radii2[0].cpu().numpy(), fos[0]["radii"]), case
            assert np.array_equal(radii2[1].cpu().numpy(), fos[1]["radii"]), case
            fo = dict(num_rendered=fos[0]["num_rendered"] + fos[1]["num_rendered"],
                      color=0.5 * (fos[0]["color"] + fos[1]["color"][:, :, ::-1]),
                      fragile=fos[0]["fragile"] | fos[1]["fragile"][:, ::-1],
                      radii=np.maximum(fos[0]["radii"], fos[1]["radii"]))
            radii = None
        assert fo["fragile"].mean() <= 0.05, (case, variant, float(fo["fragile"].mean()))
        dL = torch.as_tensor(parity.masked_dL(fo, dL))      # no gradient through the pixels fp32 cannot decide
        go = bw(dL.numpy())
        assert n == fo["num_rendered"], (case, variant, n, fo["num_rendered"])
        if radii is not None:
            assert np.array_equal(radii.cpu().numpy(), fo["radii"]), (case, variant)
        err = np.abs(color.detach().cpu().numpy() - fo["color"])[:, ~fo["fragile"]].max(initial=0.0)
        if err > 1e-5 and only_case is not None and not variant.startswith("toast"):
            from tests.fuzz_parity import diag_pixel
            diag_pixel(fo, color.detach().cpu().numpy(), W, bg, f"{variant} W={W} H={H} P={P} back={back} sm={sm} deg={deg}")
            return worst
        assert err <= 1e-5, (case, variant, err)
        worst["fwd"] = max(worst["fwd"], float(err))
        grads = torch.autograd.grad(color, [p[k] for k in names], grad_outputs=dL.to(dev))
        _check_grads(case, variant, names, grads, go, fo["radii"] > 0, P, worst, lambda: bw(np.abs(dL.numpy())))
        if verbose:
            print(f"case {case:3d} ok: {variant:5s} {W}x{H} P={P} R={fo['num_rendered']} deg={deg} back={back} sm={sm} "
                  f"fwd_err={err:.1e}", flush=True)
    print(f"{n_cases} variant cases ok in {time.time() - t0:.0f} s; worst fwd err {worst['fwd']:.2e}, "
          f"worst grad rel err {worst['grad']:.2e}({worst.get('grad_at', '-')}); gradients judged at the cancellation-aware bar: {worst.get('cancelled', 0)}; "
          f"every visible Gaussian compared ({worst.get('rows', 0)} rows), largest share of a tensor's rows missing the per-Gaussian criterion {worst.get('row_fail', 0.0):.1e}")
    return worst


if __name__ == "__main__":
    DIAG = len(sys.argv) > 3
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 30, int(sys.argv[2]) if len(sys.argv) > 2 else 0,
        only_case=int(sys.argv[3]) if len(sys.argv) > 3 else None)

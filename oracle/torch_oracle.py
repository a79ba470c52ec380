"""Pure-PyTorch CPU evaluation of the orthographic TSW splatting equation — TEST INFRASTRUCTURE.

PARITY UNPINNED: the reference's own rasterizer (github.com/actcwlf/ortho_diff_gaussian_rasterization,
/root/reference/README.md:52) is un-vendored, un-pinned and absent; this module restates SURVEY.md
Appendix A / DESIGN.md §2 (U1..U8), independently of oracle/splat_oracle.c: it is vectorised
(cumulative products instead of a sequential loop) and gets its gradients from autograd rather than
from hand-derived formulas, so agreement between the two restatements is a real cross-check.

It is also the "pure-PyTorch CPU evaluation of the same splatting equation" that BASELINE.json's
north_star asks to be reported beside the GPU numbers (bench.py cpu_baseline / --impl reference).

Call-site contract it mirrors: /root/reference/ortho_gaussian_renderer/renderer.py:63-98,
preprocess.py:58-104.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import it.
"""
from __future__ import annotations

import numpy as np
import torch

TILE = 16
ALPHA_MIN = 1.0 / 255.0
ALPHA_MAX = 0.99
T_STOP = 1e-4
LOWPASS = 0.3

# /root/reference/utils/sh_utils.py:26-43
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def _t(x, dtype):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.to(device="cpu", dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def build_rotation(q):
    """(r,x,y,z) → R, /root/reference/utils/general_utils.py:98-119 minus the normalisation
    (the rasterizer uses the quaternion as given; callers normalise, guassian.py:287)."""
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(-1, 3, 3)


def eval_sh_rgb(deg, shs, dirs):
    """shs [P,M,3], dirs [P,3] unit. Basis of /root/reference/utils/sh_utils.py:57-110; returns pre-clamp rgb+0.5."""
    sh = shs.permute(0, 2, 1)  # [P,3,M]
    res = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
        res = res - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            res = (res + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                   + C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
                       + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                       + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14]
                       + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return res + 0.5


def ordered_u32(z32: np.ndarray) -> np.ndarray:
    """U2 order-preserving float32 → uint32."""
    u = z32.astype(np.float32).view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def preprocess(st, means3D, scales=None, rotations=None, cov3D_precomp=None, opacities=None,
               colors_precomp=None, shs=None, dtype=torch.float32):
    """Appendix A.1, vectorised. `st` has the 13 settings fields (OracleSettings or the product NamedTuple
    with host values)."""
    W, H = int(st.image_width), int(st.image_height)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    V = _t(np.asarray(st.viewmatrix, dtype=np.float64), dtype)
    p = means3D
    Wm = V[:3, :3]
    pv = p @ Wm.T + V[:3, 3]
    scale = float(st.scale)
    if cov3D_precomp is not None:
        c6 = cov3D_precomp
        Sigma = torch.stack([c6[:, 0], c6[:, 1], c6[:, 2], c6[:, 1], c6[:, 3], c6[:, 4], c6[:, 2], c6[:, 4], c6[:, 5]],
                            dim=-1).reshape(-1, 3, 3)
    else:
        R = build_rotation(rotations)
        M = R * (float(st.scale_modifier) * scales)[:, None, :]
        Sigma = M @ M.transpose(1, 2)
    Sv = Wm[:2] @ Sigma @ Wm[:2].T  # [P,2,2]
    a = scale * scale * Sv[:, 0, 0] + LOWPASS
    b = scale * scale * Sv[:, 0, 1]
    c = scale * scale * Sv[:, 1, 1] + LOWPASS
    det = a * c - b * b
    with torch.no_grad():
        ok = (pv[:, 2].abs() <= float(st.threshold)) & (det != 0)
    det_safe = torch.where(ok, det, torch.ones_like(det))
    conic = _ConicInverse.apply(a, b, c, det_safe)
    pix = (pv[:, :2] - torch.tensor([float(st.x_min), float(st.y_min)], dtype=dtype)) * scale - 0.5
    with torch.no_grad():
        mid = 0.5 * (a + c)
        lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
        rad = torch.ceil(3.0 * torch.sqrt(lam))
        rminx = torch.clamp(torch.trunc((pix[:, 0] - rad) / TILE), 0, gx)
        rminy = torch.clamp(torch.trunc((pix[:, 1] - rad) / TILE), 0, gy)
        rmaxx = torch.clamp(torch.trunc((pix[:, 0] + rad + (TILE - 1)) / TILE), 0, gx)
        rmaxy = torch.clamp(torch.trunc((pix[:, 1] + rad + (TILE - 1)) / TILE), 0, gy)
        area = (rmaxx - rminx) * (rmaxy - rminy)
        ok = ok & (area > 0)
        radii = torch.where(ok, rad, torch.zeros_like(rad)).to(torch.int32)
        rect = torch.stack([rminx, rminy, rmaxx, rmaxy], dim=-1).to(torch.int64)
        rect = rect * ok[:, None]
    if colors_precomp is not None:
        rgb = colors_precomp
    elif shs is not None:
        d = p - _t(np.asarray(st.campos, dtype=np.float64), dtype)
        d = d / d.norm(dim=-1, keepdim=True)
        rgb = torch.clamp_min(eval_sh_rgb(int(st.sh_degree), shs, d), 0.0)  # clamp: zero grad where < 0
    else:
        rgb = None
    return dict(pv=pv, pix=pix, conic=conic, rgb=rgb, radii=radii, rect=rect, ok=ok, opac=opacities)


def bin_and_sort(st, pre):
    """A.2 with numpy integer arithmetic; stable argsort == stable LSD radix sort order."""
    W, H = int(st.image_width), int(st.image_height)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    rect = pre["rect"].numpy()
    ok = pre["ok"].numpy()
    ids = np.nonzero(ok)[0]
    rw = (rect[ids, 2] - rect[ids, 0]).astype(np.int64)
    rh = (rect[ids, 3] - rect[ids, 1]).astype(np.int64)
    cnt = rw * rh
    R = int(cnt.sum())
    gid = np.repeat(ids, cnt)
    start = np.cumsum(cnt) - cnt
    local = np.arange(R, dtype=np.int64) - np.repeat(start, cnt)
    rw_rep = np.repeat(rw, cnt)
    ty = rect[gid, 1] + local // np.maximum(rw_rep, 1)
    tx = rect[gid, 0] + local % np.maximum(rw_rep, 1)
    tile = (ty * gx + tx).astype(np.uint64)
    depth32 = pre["pv"][:, 2].detach().to(torch.float32).numpy()
    keys = (tile << np.uint64(32)) | ordered_u32(depth32)[gid].astype(np.uint64)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    point_list = gid[order].astype(np.uint32)
    T = gx * gy
    ranges = np.zeros((T, 2), np.uint32)
    if R > 0:
        st_tile = (skeys >> np.uint64(32)).astype(np.int64)
        first = np.nonzero(np.r_[True, st_tile[1:] != st_tile[:-1]])[0]
        last = np.r_[first[1:], R]
        ranges[st_tile[first], 0] = first
        ranges[st_tile[first], 1] = last
    return dict(R=R, keys=skeys, point_list=point_list, ranges=ranges, unsorted_keys=keys,
                unsorted_vals=gid.astype(np.uint32))


class _ConicInverse(torch.autograd.Function):
    """conic = (c, -b, a) / det with the forward exactly as the C oracle and the kernels compute it (one reciprocal,
    three products, in the working precision) and the BACKWARD in double: the derivative of the inverse of a nearly
    singular 2x2 covariance (elongated Gaussians) loses 1e-4..1e-3 in fp32, which is why the kernels
    (preprocess_bwd.cu) and the C oracle run this chain in double too."""

    @staticmethod
    def forward(ctx, a, b, c, det_safe):
        ctx.save_for_backward(a, b, c, det_safe)
        det_inv = 1.0 / det_safe
        return torch.stack([c * det_inv, -b * det_inv, a * det_inv], dim=-1)

    @staticmethod
    def backward(ctx, g):
        a, b, c, det_safe = ctx.saved_tensors
        ad, bd, cd, gd = a.double(), b.double(), c.double(), g.double()
        det = ad * cd - bd * bd
        det = torch.where(det == 0, torch.ones_like(det), det)      # culled rows: their incoming gradient is zero
        i2 = 1.0 / (det * det)
        gA, gB, gC = gd[:, 0], gd[:, 1], gd[:, 2]
        ga = (-cd * cd * gA + bd * cd * gB - bd * bd * gC) * i2
        gb = (2 * bd * cd * gA - (ad * cd + bd * bd) * gB + 2 * ad * bd * gC) * i2
        gc = (-bd * bd * gA + ad * bd * gB - ad * ad * gC) * i2
        return ga.to(a.dtype), gb.to(a.dtype), gc.to(a.dtype), None


def _blend_tiles(st, tiles, ranges, point_list, pix, conic, opac, rgb, bg, dtype, exponent="quadratic"):
    """Blend a batch of tiles, padded to the longest list. Returns color [B,256,3], final_T, n_contrib [B,256]."""
    W = int(st.image_width)
    gx = (W + TILE - 1) // TILE
    B = len(tiles)
    lens = (ranges[tiles, 1].astype(np.int64) - ranges[tiles, 0].astype(np.int64))
    Lmax = int(max(1, lens.max()))
    k = np.arange(Lmax, dtype=np.int64)[None, :]
    valid = k < lens[:, None]
    src = np.where(valid, ranges[tiles, 0].astype(np.int64)[:, None] + k, 0)
    if point_list.shape[0] == 0:
        gid = np.zeros((B, Lmax), np.int64)
    else:
        gid = point_list[src].astype(np.int64)
    gid_t = torch.from_numpy(gid)
    valid_t = torch.from_numpy(valid)
    tx = torch.from_numpy((tiles % gx).astype(np.int64))
    ty = torch.from_numpy((tiles // gx).astype(np.int64))
    lx = torch.arange(TILE).repeat(TILE)
    ly = torch.arange(TILE).repeat_interleave(TILE)
    px = (tx[:, None] * TILE + lx[None, :]).to(dtype)  # [B,256]
    py = (ty[:, None] * TILE + ly[None, :]).to(dtype)
    gxy = pix[gid_t]        # [B,L,2]
    gcon = conic[gid_t]     # [B,L,3]
    gop = opac[gid_t]       # [B,L]
    grgb = rgb[gid_t]       # [B,L,3]
    dx = gxy[:, None, :, 0] - px[:, :, None]
    dy = gxy[:, None, :, 1] - py[:, :, None]
    if exponent == "quadratic":
        power = -0.5 * (gcon[:, None, :, 0] * dx * dx + gcon[:, None, :, 2] * dy * dy) - gcon[:, None, :, 1] * dx * dy
        Gs = torch.exp(power)
    else:
        # The CUDA blend's evaluation order (gsvc_b200/csrc/preprocess.cu feat0/feat3, render.cu neg_falloff_log2):
        # -power*log2(e) as the sum of squares (l11 (dx + rho dy))^2 + (l22 dy)^2, L = (l11 0; l11 rho, l22) the
        # Cholesky factor of the conic times log2(e)/2 with the Schur complement from a compensated A*C - B*B and
        # the shear rho = B/A carried as a sum of two floats, accumulated with FMAs; Gs = 2^-q.  Same function of
        # (A, B, C, dx, dy), different rounding: what tests/test_oracle.py uses to check that the C oracle's
        # `fragile` map covers the pixels where two correct fp32 evaluations may disagree.
        cA, cB, cC = gcon[..., 0], gcon[..., 1], gcon[..., 2]
        kL = 0.5 * 1.4426950408889634
        pac, pbb = cA * cC, cB * cB
        with torch.no_grad():   # exact residuals of the two products (what fmaf(a, c, -a*c) returns)
            rac = (cA.double() * cC.double() - pac.double()).to(dtype)
            rbb = (cB.double() * cB.double() - pbb.double()).to(dtype)
        det_c = (pac - pbb) + (rac - rbb)
        aL = cA * kL
        l11 = aL * torch.rsqrt(aL)
        d22 = torch.clamp(det_c / cA * kL, min=1e-30)
        l22 = d22 * torch.rsqrt(d22)
        rho = cB.double() / cA.double()
        rho_hi = rho.to(dtype)
        rho_lo = (rho - rho_hi.double()).to(dtype)
        # fma(rho_lo, dy, fma(rho_hi, dy, dx)): exact product + sum in double, rounded once per FMA
        t = (dx.double() + rho_hi.double()[:, None, :] * dy.double()).to(dtype)
        t = (t.double() + rho_lo.double()[:, None, :] * dy.double()).to(dtype)
        u = l11[:, None, :] * t
        v = l22[:, None, :] * dy
        q = u * u + v * v
        power = -q                 # <= 0 by construction (in units of log2)
        Gs = torch.exp2(power)
    a_raw = gop[:, None, :] * Gs
    alpha = a_raw + (torch.clamp(a_raw, max=ALPHA_MAX) - a_raw).detach()  # U4 straight-through
    with torch.no_grad():
        keep = valid_t[:, None, :] & (power <= 0) & (alpha >= ALPHA_MIN)
    a_eff = torch.where(keep, alpha, torch.zeros_like(alpha))
    one_m = 1.0 - a_eff
    Tcum = torch.cumprod(one_m, dim=-1)
    Tprev = torch.cat([torch.ones_like(Tcum[..., :1]), Tcum[..., :-1]], dim=-1)
    with torch.no_grad():
        live = Tcum >= T_STOP
        contrib = live & keep
        pos = torch.arange(1, Lmax + 1)[None, None, :]
        n_contrib = (pos * contrib).amax(dim=-1)
    w = torch.where(contrib, a_eff * Tprev, torch.zeros_like(a_eff))
    col = torch.einsum("bpl,blc->bpc", w, grgb)
    final_T = torch.where(live, one_m, torch.ones_like(one_m)).prod(dim=-1)
    col = col + final_T[..., None] * bg[None, None, :]
    return col, final_T, n_contrib, px.long(), py.long()


def forward(st, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None, colors_precomp=None,
            shs=None, dtype=torch.float32, requires_grad=False, elem_budget=6_000_000, tile_subset=None,
            exponent="quadratic"):
    """Full forward. Returns dict(color [3,H,W] torch, radii, num_rendered, keys, point_list, ranges, leaves)."""
    import time as _time
    _t0 = _time.perf_counter()
    leaves = {}

    def leaf(name, x):
        if x is None:
            return None
        t = _t(x, dtype).clone()
        if requires_grad:
            t.requires_grad_(True)
        leaves[name] = t
        return t

    means3D = leaf("means3D", means3D)
    P = means3D.shape[0]
    opacities_l = leaf("opacities", opacities)
    scales = leaf("scales", scales)
    rotations = leaf("rotations", rotations)
    cov3D_precomp = leaf("cov3D_precomp", cov3D_precomp)
    colors_precomp = leaf("colors_precomp", colors_precomp)
    shs = leaf("shs", shs)
    pre = preprocess(st, means3D, scales, rotations, cov3D_precomp, opacities_l.reshape(P), colors_precomp, shs, dtype)
    # means2D receives the pixel gradient (U3 scaling applied by `backward`)
    pixgrad_holder = torch.zeros_like(pre["pix"], requires_grad=requires_grad)
    pix = pre["pix"] + pixgrad_holder
    leaves["_pix_holder"] = pixgrad_holder
    b = bin_and_sort(st, pre)
    t_pre = _time.perf_counter() - _t0
    W, H = int(st.image_width), int(st.image_height)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    bg = _t(np.asarray(st.bg, dtype=np.float64), dtype)
    color = torch.zeros(3, gy * TILE, gx * TILE, dtype=dtype)
    final_T = torch.ones(gy * TILE, gx * TILE, dtype=dtype)
    n_contrib = torch.zeros(gy * TILE, gx * TILE, dtype=torch.int64)
    ranges = b["ranges"]
    lens = ranges[:, 1].astype(np.int64) - ranges[:, 0].astype(np.int64)
    all_tiles = np.arange(gx * gy) if tile_subset is None else np.asarray(tile_subset)
    order = all_tiles[np.argsort(-lens[all_tiles], kind="stable")]
    i = 0
    pieces = []
    while i < len(order):
        L = max(1, int(lens[order[i]]))
        nb = max(1, min(len(order) - i, elem_budget // (256 * L)))
        tiles = order[i:i + nb]
        i += nb
        col, fT, nc, px, py = _blend_tiles(st, tiles, ranges, b["point_list"], pix, pre["conic"], pre["opac"],
                                           pre["rgb"], bg, dtype, exponent)
        pieces.append((col, fT, nc, px, py))
    # scatter pieces into the padded image (index_put keeps autograd)
    cols = torch.cat([p_[0].reshape(-1, 3) for p_ in pieces])
    pxs = torch.cat([p_[3].reshape(-1) for p_ in pieces])
    pys = torch.cat([p_[4].reshape(-1) for p_ in pieces])
    color = color.index_put((torch.arange(3)[None, :].expand(cols.shape[0], 3), pys[:, None].expand(-1, 3),
                             pxs[:, None].expand(-1, 3)), cols)
    with torch.no_grad():
        final_T[pys, pxs] = torch.cat([p_[1].reshape(-1) for p_ in pieces]).detach()
        n_contrib[pys, pxs] = torch.cat([p_[2].reshape(-1) for p_ in pieces])
    color = color[:, :H, :W]
    return dict(color=color, radii=pre["radii"].numpy(), num_rendered=b["R"], keys=b["keys"],
                point_list=b["point_list"], ranges=ranges, final_T=final_T[:H, :W], n_contrib=n_contrib[:H, :W],
                leaves=leaves, pre=pre, bin=b, settings=st, t_pre=t_pre)


def backward(fwd, dL_dout):
    """Autograd backward; returns grads dict (numpy) keyed like the rasterizer inputs; means2D per U3."""
    st = fwd["settings"]
    leaves = fwd["leaves"]
    color = fwd["color"]
    g_out = _t(dL_dout, color.dtype).reshape(color.shape)
    names = [k for k, v in leaves.items() if v is not None and v.requires_grad]
    grads = torch.autograd.grad(color, [leaves[k] for k in names], grad_outputs=g_out, allow_unused=True)
    out = {}
    for k, g in zip(names, grads):
        ref = leaves[k]
        out[k] = (torch.zeros_like(ref) if g is None else g).numpy()
    gp = out.pop("_pix_holder")
    m2d = np.zeros((gp.shape[0], 3), gp.dtype)
    m2d[:, 0] = gp[:, 0] * 0.5 * int(st.image_width)
    m2d[:, 1] = gp[:, 1] * 0.5 * int(st.image_height)
    out["means2D"] = m2d
    out["_dL_dpix"] = gp
    return out

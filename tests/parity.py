"""The parity checker every GPU test goes through (round-2 hardening; VERDICT r1 "weak" 1a-1d).

Oracle side: the C restatement in REFEREE mode (oracle/splat_oracle.c orc_set_power_mode: the blend exponent in
double from the fp32 conic = the exact value of SPEC's formula, `fragile` = only the pixels where a well-conditioned
fp32 evaluation cannot decide one of the discontinuous tests).  Checker side, no escape hatches:
  * forward: every non-fragile pixel within 1e-5; the fragile share of the image is bounded per test;
  * gradients: the seed gradient dL/dimage is ZEROED ON THE FRAGILE PIXELS (masked_dL) for both sides.  Pixels are
    independent in the blend, so whichever way a fragile decision falls it then contributes nothing, and EVERY
    visible Gaussian is compared — no Gaussian is set aside (round 1 excluded every Gaussian that touched a fragile
    pixel: 3-20 % of them, 70 % for screen-filling splats).  What is not exercised is the gradient of <= 0.5 % of
    the pixels, whose code path is that of every other pixel;
  * per tensor: (a) max|a-b| <= 1e-4 max|b| (the north-star bar); (b) PER GAUSSIAN, row-wise:
    |a-b| <= ROW_RTOL |b_row| + ROW_ATOL max|b| for EVERY visible Gaussian (ROW_FAIL_MAX = 0; the share of rows
    that miss it and the worst row in units of its tolerance are returned for the sweep records);
  * culled Gaussians must have exactly zero gradients; a scene with nothing visible must have all-zero gradients.
"""
from __future__ import annotations

import numpy as np

from oracle import c_oracle

FWD_ATOL = 1e-5          # north_star: forward pixels within 1e-5 absolute
GRAD_RTOL = 1e-4         # north_star: gradients within 1e-4 relative (of the tensor's largest entry)
ROW_RTOL = 1e-4          # per-Gaussian criterion: |a - b| <= ROW_RTOL * max_c|b_row| + ROW_ATOL * max|b|
ROW_ATOL = 1e-6
ROW_FAIL_MAX = 0.0       # no Gaussian may miss the per-Gaussian criterion (measured on B200: worst row 0.44x its
ROW_HARD = 1.0           # tolerance over the sweeps in profiles/r3_fuzz_*.txt, needles of 256:1 axes included)
MAX_FRAGILE = 2e-3       # share of the pixels the oracle may call undecidable on an ordinary scene

GRAD_NAMES = ("means3D", "scales", "rotations", "opacities", "colors_precomp")


def oracle_forward(st, gi, referee=True, **over):
    """C oracle forward on numpy inputs `gi` (keys as GaussianRasterizer's arguments), referee mode by default."""
    gi = dict(gi)
    gi.update(over)
    return c_oracle.forward(st, gi["means3D"], gi["opacities"], gi.get("scales"), gi.get("rotations"),
                            cov3D_precomp=gi.get("cov3D_precomp"), colors_precomp=gi.get("colors_precomp"),
                            shs=gi.get("shs"), referee=referee)


def masked_dL(fo, dL, fragile_extra=None):
    """dL/dimage [.., H, W] with the oracle's fragile pixels zeroed (numpy float32 copy)."""
    dL = np.array(dL.detach().cpu().numpy() if hasattr(dL, "detach") else dL, dtype=np.float32, copy=True)
    frag = fo["fragile"] if fragile_extra is None else (fo["fragile"] | fragile_extra)
    dL[..., frag] = 0.0
    return dL


def oracle_backward(fo, dL):
    """dL must come from masked_dL: the fragile pixels carry no gradient."""
    dL = np.asarray(dL, np.float32)
    assert not dL[..., fo["fragile"]].any(), "seed gradient is not masked on the fragile pixels (parity.masked_dL)"
    return c_oracle.backward(fo, dL, narrow_touched=True)


def check_forward(fo, color, fT=None, nc=None, max_fragile=MAX_FRAGILE, fragile_extra=None):
    """Pixels: <= 1e-5 off the fragile set, <= one flipped decision on it; n_contrib exact and final_T within 1e-5
    off it.  Returns dict(err, fragile)."""
    got = color.detach().cpu().numpy() if hasattr(color, "detach") else np.asarray(color)
    frag = fo["fragile"] if fragile_extra is None else (fo["fragile"] | fragile_extra)
    err = np.abs(got - fo["color"])
    solid = float(err[:, ~frag].max(initial=0.0))
    assert solid <= FWD_ATOL, f"max abs err {solid:.3e} on non-fragile pixels"
    share = float(frag.mean())
    assert share <= max_fragile, f"the oracle calls {share:.2e} of the pixels fragile (bound {max_fragile:.1e})"
    if frag.any():
        assert err[:, frag].max() <= 2.0 / 255.0 + 1e-3      # one flipped alpha-floor / stop decision at most
    if nc is not None:
        nc = nc.cpu().numpy() if hasattr(nc, "cpu") else np.asarray(nc)
        np.testing.assert_array_equal(nc.view(np.uint32)[~frag], fo["n_contrib"][~frag])
    if fT is not None:
        fT = fT.cpu().numpy() if hasattr(fT, "cpu") else np.asarray(fT)
        assert np.abs(fT - fo["final_T"])[~frag].max(initial=0.0) <= 1e-5
    return dict(err=solid, fragile=share)


def grad_stats(a, b):
    """a (got), b (oracle) [n, w] over the compared rows: max-normalised error, share of rows missing the
    per-Gaussian criterion, worst row in units of its own tolerance."""
    scale = float(np.abs(b).max(initial=0.0))
    if a.size == 0:
        return dict(rel=0.0, row_fail=0.0, row_worst=0.0, scale=scale, n=0)
    diff = np.abs(a - b).max(axis=1)
    tol = ROW_RTOL * np.abs(b).max(axis=1) + ROW_ATOL * scale + 1e-30
    ratio = diff / tol
    return dict(rel=float(diff.max() / (scale + 1e-30)), row_fail=float((ratio > 1.0).mean()),
                row_worst=float(ratio.max()), scale=scale, n=int(a.shape[0]))


SMALL_SCENE = 50         # below this many visible Gaussians a tensor may be judged at the cancellation-aware bar


def check_grads(fo, go, got, names=None, rel_tol=GRAD_RTOL, row_fail_max=ROW_FAIL_MAX, row_hard=ROW_HARD, what="",
                cancel_ref=None):
    """`got`: name -> tensor / array with P rows (gradients for a masked_dL seed); `go`: oracle_backward result.
    Every visible Gaussian is compared.  Asserts the bars in the module docstring; returns dict(compared, rel,
    row_fail, row_worst) (worst over the tensors) for the sweep records.

    `cancel_ref` (randomised sweeps only): a callable returning the oracle backward for |dL|.  A scene of a handful
    of Gaussians has no "largest gradient" to normalise by — each entry is a sum of signed per-pixel terms that can
    cancel to far below any one of them, and then 1e-4 of the RESULT is below the fp32 rounding of its terms.  For
    scenes with fewer than SMALL_SCENE visible Gaussians a tensor that misses the bars is re-judged against the same
    gradient taken with |dL| (no cancellation over pixels): |a-b| <= 1e-5 of its largest entry.  Counted in
    out["cancelled"] and reported by the sweeps; never used by the fixed-scene tests."""
    vis = fo["radii"] > 0
    P = vis.shape[0]
    names = tuple(got.keys()) if names is None else names
    out = dict(compared=1.0, n=int(vis.sum()), rel=0.0, row_fail=0.0, row_worst=0.0, cancelled=0)
    go_abs = None
    for k in names:
        a = got[k].detach().cpu().numpy() if hasattr(got[k], "detach") else np.asarray(got[k])
        a = a.reshape(P, -1).astype(np.float64)
        b = np.asarray(go[k], np.float64).reshape(P, -1)
        assert np.isfinite(a).all(), f"{what}{k}: non-finite gradient"
        assert not a[~vis].any(), f"{what}{k}: a culled Gaussian received a gradient"
        if not vis.any():
            continue
        s = grad_stats(a[vis], b[vis])
        if (cancel_ref is not None and int(vis.sum()) < SMALL_SCENE and
                (s["rel"] > rel_tol or s["row_fail"] > row_fail_max or s["row_worst"] > row_hard)):
            go_abs = go_abs if go_abs is not None else cancel_ref()
            ref = float(np.abs(np.asarray(go_abs[k], np.float64).reshape(P, -1)[vis]).max(initial=0.0))
            err = float(np.abs(a[vis] - b[vis]).max())
            assert err <= 1e-5 * ref, f"{what}{k}: error {err:.3e} > 1e-5 of the uncancelled gradient {ref:.3e}"
            out["cancelled"] += 1
            continue
        assert s["rel"] <= rel_tol, f"{what}{k}: max-normalised error {s['rel']:.3e} > {rel_tol:.0e}"
        assert s["row_fail"] <= row_fail_max, (f"{what}{k}: {s['row_fail']:.4f} of {s['n']} Gaussians miss "
                                               f"|a-b| <= {ROW_RTOL:.0e}|b_row| + {ROW_ATOL:.0e} max|b|")
        assert s["row_worst"] <= row_hard, f"{what}{k}: a Gaussian is {s['row_worst']:.1f}x outside its own tolerance"
        out["rel"] = max(out["rel"], s["rel"])
        out["row_fail"] = max(out["row_fail"], s["row_fail"])
        out["row_worst"] = max(out["row_worst"], s["row_worst"])
    return out

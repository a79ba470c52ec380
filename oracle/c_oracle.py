"""ctypes front-end of oracle/splat_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED (see splat_oracle.c header): the reference's rasterizer source is an
un-vendored dependency (/root/reference/README.md:52); this oracle restates the algorithm
frozen in SURVEY.md Appendix A / DESIGN.md §2.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsplat_oracle.so")
TILE = 16


class _Settings(C.Structure):
    _fields_ = [
        ("W", C.c_int32), ("H", C.c_int32),
        ("x_min", C.c_float), ("y_min", C.c_float), ("scale", C.c_float),
        ("threshold", C.c_float), ("scale_modifier", C.c_float),
        ("bg", C.c_float * 3), ("V", C.c_float * 16),
        ("sh_degree", C.c_int32), ("sh_M", C.c_int32), ("campos", C.c_float * 3),
    ]


@dataclass
class OracleSettings:
    """Same 13 fields as GaussianRasterizationSettings (renderer.py:63-83), host-side values."""
    image_height: int
    image_width: int
    x_min: float
    y_min: float
    scale: float
    threshold: float
    bg: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    scale_modifier: float = 1.0
    viewmatrix: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))  # logical V[r][c]
    sh_degree: int = 0
    campos: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    prefiltered: bool = False
    debug: bool = False

    @property
    def grid(self):
        return ((self.image_width + TILE - 1) // TILE, (self.image_height + TILE - 1) // TILE)

    def to_c(self, sh_M: int = 0) -> _Settings:
        s = _Settings()
        s.W, s.H = int(self.image_width), int(self.image_height)
        s.x_min, s.y_min, s.scale = float(self.x_min), float(self.y_min), float(self.scale)
        s.threshold, s.scale_modifier = float(self.threshold), float(self.scale_modifier)
        s.bg[:] = [float(v) for v in np.asarray(self.bg, np.float32).reshape(3)]
        s.V[:] = [float(v) for v in np.asarray(self.viewmatrix, np.float32).reshape(16)]
        s.sh_degree, s.sh_M = int(self.sh_degree), int(sh_M)
        s.campos[:] = [float(v) for v in np.asarray(self.campos, np.float32).reshape(3)]
        return s


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (building the checker is not using it)."""
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_count_instances.restype = C.c_int64
    return _lib


def set_threads(n: int) -> int:
    """OpenMP threads of the oracle's parallel loops.  torchrun exports OMP_NUM_THREADS=1 to its workers, which
    libgomp reads once at load time — a bench leg that wants every host core has to ask for them explicitly."""
    L = lib()
    L.omp_set_num_threads(C.c_int(int(n)))
    L.omp_get_max_threads.restype = C.c_int
    return int(L.omp_get_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


def key_bits(st: OracleSettings) -> int:
    gx, gy = st.grid
    n = gx * gy
    return 32 + max(1, int(np.ceil(np.log2(max(n, 2)))))


def preprocess(st: OracleSettings, means3D, scales=None, rotations=None, cov3D_precomp=None,
               opacities=None, colors_precomp=None, shs=None):
    means3D = _f32(means3D, (-1, 3))
    P = means3D.shape[0]
    scales, rotations = _f32(scales, (P, 3)), _f32(rotations, (P, 4))
    cov3D_precomp = _f32(cov3D_precomp, (P, 6))
    opacities = _f32(opacities, (P,))
    colors_precomp = _f32(colors_precomp, (P, 3))
    sh_M = 0
    if shs is not None:
        shs = _f32(shs)
        sh_M = shs.shape[1]
        shs = shs.reshape(P, sh_M, 3)
    out = dict(
        radii=np.zeros(P, np.int32), depth=np.zeros(P, np.float32), xy=np.zeros((P, 2), np.float32),
        conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
        rect=np.zeros((P, 4), np.int32), tiles_touched=np.zeros(P, np.int32),
        clamped=np.zeros((P, 3), np.uint8), cov3D=np.zeros((P, 6), np.float32),
    )
    cs = st.to_c(sh_M)
    lib().orc_preprocess(C.byref(cs), C.c_int(P), _p(means3D), _p(scales), _p(rotations), _p(cov3D_precomp),
                         _p(opacities), _p(colors_precomp), _p(shs), _p(out["radii"]), _p(out["depth"]),
                         _p(out["xy"]), _p(out["conic_opacity"]), _p(out["rgb"]), _p(out["rect"]),
                         _p(out["tiles_touched"]), _p(out["clamped"]), _p(out["cov3D"]))
    out["_inputs"] = dict(means3D=means3D, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp,
                          opacities=opacities, colors_precomp=colors_precomp, shs=shs, sh_M=sh_M)
    return out


def visible_filter(st: OracleSettings, means3D, scales=None, rotations=None, cov3D_precomp=None):
    """preprocess.py:99-104 — radii only (U7)."""
    return preprocess(st, means3D, scales, rotations, cov3D_precomp)["radii"]


def bin_and_sort(st: OracleSettings, pre):
    """A.2: duplicateWithKeys → stable sort → identifyTileRanges."""
    P = pre["radii"].shape[0]
    R = int(lib().orc_count_instances(C.c_int(P), _p(pre["tiles_touched"])))
    keys = np.zeros(max(R, 1), np.uint64)
    vals = np.zeros(max(R, 1), np.uint32)
    cs = st.to_c()
    lib().orc_duplicate_with_keys(C.byref(cs), C.c_int(P), _p(pre["radii"]), _p(pre["rect"]), _p(pre["depth"]),
                                  _p(keys), _p(vals))
    unsorted_keys, unsorted_vals = keys[:R].copy(), vals[:R].copy()
    lib().orc_sort_pairs(C.c_int64(R), C.c_int(key_bits(st)), _p(keys), _p(vals))
    gx, gy = st.grid
    ranges = np.zeros((gx * gy, 2), np.uint32)
    lib().orc_tile_ranges(C.c_int64(R), _p(keys), C.c_int(gx * gy), _p(ranges))
    return dict(R=R, keys=keys[:R], point_list=vals[:R], ranges=ranges,
                unsorted_keys=unsorted_keys, unsorted_vals=unsorted_vals)


def forward(st: OracleSettings, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None,
            colors_precomp=None, shs=None, frag_eps: float = 4e-6, referee: bool = False):
    """Full forward: returns dict(color[3,H,W], radii[P], num_rendered, + all intermediate state).
    `referee`: evaluate the blend exponent in double (the exact value of SPEC's formula on the fp32 conic) and
    report as `fragile` only what a well-conditioned fp32 implementation cannot decide — see splat_oracle.c
    orc_set_power_mode.  Default off: the fp32 three-term form every GPU parity test is held to."""
    pre = preprocess(st, means3D, scales, rotations, cov3D_precomp, opacities, colors_precomp, shs)
    b = bin_and_sort(st, pre)
    H, W = st.image_height, st.image_width
    color = np.zeros((3, H, W), np.float32)
    final_T = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.uint32)
    fragile = np.zeros((H, W), np.uint8)
    pl = b["point_list"] if b["R"] > 0 else np.zeros(1, np.uint32)
    cs = st.to_c(pre["_inputs"]["sh_M"])
    lib().orc_set_power_mode(C.c_int(1 if referee else 0))
    lib().orc_render_forward(C.byref(cs), _p(b["ranges"]), _p(pl), _p(pre["xy"]), _p(pre["conic_opacity"]),
                             _p(pre["rgb"]), _p(color), _p(final_T), _p(n_contrib), _p(fragile),
                             C.c_float(frag_eps))
    return dict(color=color, radii=pre["radii"], num_rendered=b["R"], final_T=final_T, n_contrib=n_contrib,
                fragile=fragile.astype(bool), pre=pre, bin=b, settings=st, referee=bool(referee))


def backward(fwd, dL_dout, narrow_touched: bool = False):
    """A.4: returns grads dict (float64) for means3D, means2D, scales, rotations | cov3D_precomp,
    colors_precomp | shs, opacities, plus `touched_fragile[P]`."""
    st: OracleSettings = fwd["settings"]
    pre, b = fwd["pre"], fwd["bin"]
    inp = pre["_inputs"]
    P = pre["radii"].shape[0]
    dL_dout = _f32(dL_dout, (3, st.image_height, st.image_width))
    d_pix = np.zeros((P, 2), np.float64)
    d_conic = np.zeros((P, 3), np.float64)
    d_op = np.zeros(P, np.float64)
    d_rgb = np.zeros((P, 3), np.float64)
    touched = np.zeros(P, np.uint8)
    fragile = np.ascontiguousarray(fwd["fragile"].astype(np.uint8))
    pl = b["point_list"] if b["R"] > 0 else np.zeros(1, np.uint32)
    cs = st.to_c(inp["sh_M"])
    lib().orc_set_touched_mode(C.c_int(1 if narrow_touched else 0))   # see splat_oracle.c: default = whole tile list
    lib().orc_set_power_mode(C.c_int(1 if fwd.get("referee") else 0))     # replay on the forward's own values
    lib().orc_render_backward(C.byref(cs), _p(b["ranges"]), _p(pl), _p(pre["xy"]), _p(pre["conic_opacity"]),
                              _p(pre["rgb"]), _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(dL_dout), C.c_int(P),
                              _p(d_pix), _p(d_conic), _p(d_op), _p(d_rgb), _p(fragile), _p(touched))
    have_cov = inp["cov3D_precomp"] is not None
    have_sh = inp["shs"] is not None
    g = dict(
        means3D=np.zeros((P, 3)), means2D=np.zeros((P, 3)), scales=np.zeros((P, 3)), rotations=np.zeros((P, 4)),
        cov3D_precomp=np.zeros((P, 6)), opacities=d_op.reshape(P, 1),
        shs=np.zeros((P, max(inp["sh_M"], 1), 3)), colors_precomp=np.zeros((P, 3)),
    )
    lib().orc_preprocess_backward(
        C.byref(cs), C.c_int(P), _p(pre["radii"]), _p(inp["means3D"]), _p(inp["scales"]), _p(inp["rotations"]),
        _p(pre["cov3D"]), _p(inp["shs"]), _p(pre["clamped"]), _p(d_pix), _p(d_conic), _p(d_rgb),
        C.c_int(1 if have_cov else 0), _p(g["means3D"]), _p(g["means2D"]), _p(g["scales"]), _p(g["rotations"]),
        _p(g["cov3D_precomp"]), _p(g["shs"]) if have_sh else None,
        None if have_sh else _p(g["colors_precomp"]))
    g["touched_fragile"] = touched.astype(bool)
    g["_dL_dpix"], g["_dL_dconic"], g["_dL_drgb"] = d_pix, d_conic, d_rgb
    return g

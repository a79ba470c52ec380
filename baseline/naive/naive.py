"""ctypes front-end of baseline/naive/naive_rast.cu — TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

The upstream-design GPU chain (cub::DeviceScan + duplicateWithKeys + cub::DeviceRadixSort::SortPairs + one-pixel-
per-thread blend + per-pixel-atomic backward), same SPEC as the product, built for sm_100a.  Used by
tests/test_gpu_baseline.py (second referee for sorted keys / point list / tile ranges; sanity of its own image and
gradients against the oracle) and by bench.py's `gpu_baseline` leg.  Only tests/ and bench.py may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libnaive_rast.so")
STAGES = ("preprocess", "scan", "duplicate", "radix_sort", "tile_ranges", "render_forward", "render_backward",
          "preprocess_backward")


class Settings(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("x_min", C.c_float), ("y_min", C.c_float), ("scale", C.c_float),
                ("threshold", C.c_float), ("scale_modifier", C.c_float), ("bg", C.c_float * 3), ("V", C.c_float * 16)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "naive_rast.cu")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.naive_forward.restype = C.c_longlong
        _lib.naive_forward.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8
        _lib.naive_backward.restype = C.c_int
        _lib.naive_backward.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
        _lib.naive_export.argtypes = [C.c_void_p] * 4
        _lib.naive_stage_times.argtypes = [C.c_void_p]
    return _lib


def settings_from(rs) -> Settings:
    """From a GaussianRasterizationSettings-shaped tuple (renderer.py:63-83); the view matrix is read once to the host."""
    s = Settings()
    s.W, s.H = int(rs.image_width), int(rs.image_height)
    s.x_min, s.y_min, s.scale = float(rs.x_min), float(rs.y_min), float(rs.scale)
    s.threshold, s.scale_modifier = float(rs.threshold), float(rs.scale_modifier)
    s.bg[:] = [float(v) for v in torch.as_tensor(rs.bg).detach().cpu().reshape(3).tolist()]
    s.V[:] = [float(v) for v in rs.viewmatrix.detach().cpu().float().reshape(-1).tolist()]   # logical V[r][c], row-major
    return s


class NaiveRasterizer:
    """forward(...) -> (color [3,H,W], radii [P], num_rendered); backward(dL) -> dict of gradients (state of the last
    forward is kept inside the library, the way autograd keeps the upstream buffers)."""

    def __init__(self, rs):
        self.s = settings_from(rs)
        self.H, self.W = int(rs.image_height), int(rs.image_width)

    def forward(self, means3D, scales, rotations, opacities, colors):
        L = lib()
        dev = means3D.device
        self.P = P = int(means3D.shape[0])
        self.inputs = [t.contiguous().float() for t in (means3D, scales, rotations, opacities, colors)]
        m, sc, rot, op, col = self.inputs
        color = torch.empty((3, self.H, self.W), dtype=torch.float32, device=dev)
        radii = torch.zeros((max(P, 1),), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        R = L.naive_forward(C.byref(self.s), P, m.data_ptr(), sc.data_ptr(), rot.data_ptr(), op.data_ptr(), col.data_ptr(),
                            color.data_ptr(), radii.data_ptr(), st)
        if R < 0:
            raise RuntimeError("naive_forward failed")
        self.R = int(R)
        return color, radii[:P], self.R

    def backward(self, dL):
        L = lib()
        P, dev = self.P, dL.device
        _, sc, rot, _, _ = self.inputs
        g = dict(means3D=torch.empty((P, 3), device=dev), means2D=torch.empty((P, 3), device=dev),
                 colors_precomp=torch.empty((P, 3), device=dev), opacities=torch.empty((P, 1), device=dev),
                 scales=torch.empty((P, 3), device=dev), rotations=torch.empty((P, 4), device=dev))
        dL = dL.contiguous().float()
        rc = L.naive_backward(C.byref(self.s), P, sc.data_ptr(), rot.data_ptr(), dL.data_ptr(), g["means3D"].data_ptr(),
                              g["means2D"].data_ptr(), g["colors_precomp"].data_ptr(), g["opacities"].data_ptr(),
                              g["scales"].data_ptr(), g["rotations"].data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError("naive_backward failed")
        return g

    def export(self):
        """Sorted 64-bit keys [R], point list [R], tile ranges [T,2] of the last forward."""
        L = lib()
        dev = self.inputs[0].device
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        keys = torch.zeros(max(self.R, 1), dtype=torch.int64, device=dev)
        pl = torch.zeros(max(self.R, 1), dtype=torch.int32, device=dev)
        ranges = torch.zeros((T, 2), dtype=torch.int32, device=dev)
        if L.naive_export(keys.data_ptr(), pl.data_ptr(), ranges.data_ptr(), torch.cuda.current_stream(dev).cuda_stream) != 0:
            raise RuntimeError("naive_export failed")
        return keys[:self.R], pl[:self.R], ranges


def timing(enable: bool) -> None:
    lib().naive_timing(1 if enable else 0)


def stage_times() -> dict:
    buf = (C.c_float * 8)()
    lib().naive_stage_times(buf)
    return {k: float(buf[i]) for i, k in enumerate(STAGES) if buf[i] >= 0}

"""TEST INFRASTRUCTURE — ctypes/numpy front-end of the CPU oracle (oracle/surfel_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this module.  It restates the reference rasterizer (DSR/cuda_rasterizer/{forward,backward,rasterizer_impl}.cu,
see the citations in the .cpp) and exposes EVERY intermediate the parity tests compare: radii, means2D, depths,
transMat, normal_opacity, rgb, clamped, tiles_touched, point_offsets, unsorted/sorted keys, point_list, ranges,
final_T/dist1/dist2, n_contrib/median_contributor, the 3+8 output planes and all gradient tensors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsurfel_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only; no CUDA involved)."""
    src = os.path.join(_HERE, "surfel_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libsurfel_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_forward_f32.restype = ctypes.c_int64
        _lib.orc_forward_f64.restype = ctypes.c_int64
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


@dataclass
class ForwardState:
    """Everything the reference keeps in geomBuffer / binningBuffer / imgBuffer, plus the outputs."""
    P: int
    W: int
    H: int
    num_rendered: int
    dtype: np.dtype
    inputs: dict = field(default_factory=dict)
    radii: np.ndarray = None
    means2D: np.ndarray = None
    depths: np.ndarray = None
    transMat: np.ndarray = None
    normal_opacity: np.ndarray = None
    rgb: np.ndarray = None
    clamped: np.ndarray = None
    tiles_touched: np.ndarray = None
    point_offsets: np.ndarray = None
    out_color: np.ndarray = None
    out_others: np.ndarray = None
    final_T: np.ndarray = None
    n_contrib: np.ndarray = None
    ranges: np.ndarray = None
    keys_unsorted: np.ndarray = None
    vals_unsorted: np.ndarray = None
    keys_sorted: np.ndarray = None
    point_list: np.ndarray = None


def _as(a, dt):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a), dtype=dt)
    return a if a.size else None


def forward(*, bg, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, image_height,
            image_width, shs=None, sh_degree=0, colors_precomp=None, scales=None, rotations=None,
            transMat_precomp=None, precision: str = "f32") -> ForwardState:
    """Oracle of RasterizeGaussiansCUDA (DSR/rasterize_points.cu:39-141)."""
    dt = np.float32 if precision == "f32" else np.float64
    fn = lib().orc_forward_f32 if precision == "f32" else lib().orc_forward_f64
    creal = ctypes.c_float if precision == "f32" else ctypes.c_double
    means3D = _as(means3D, dt)
    P = 0 if means3D is None else means3D.shape[0]
    H, W = int(image_height), int(image_width)
    shs_a = _as(shs, dt)
    M = 0 if shs_a is None else shs_a.shape[1]
    inp = dict(bg=_as(bg, dt), means3D=means3D, shs=shs_a, colors_precomp=_as(colors_precomp, dt),
               opacities=_as(opacities, dt), scales=_as(scales, dt), rotations=_as(rotations, dt),
               transMat_precomp=_as(transMat_precomp, dt), viewmatrix=_as(viewmatrix, dt),
               projmatrix=_as(projmatrix, dt), campos=_as(campos, dt), tanfovx=float(tanfovx),
               tanfovy=float(tanfovy), sh_degree=int(sh_degree), M=M)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    st = ForwardState(P=P, W=W, H=H, num_rendered=0, dtype=np.dtype(dt), inputs=inp)
    st.radii = np.zeros(P, np.int32)
    st.means2D = np.zeros((P, 2), dt)
    st.depths = np.zeros(P, dt)
    st.transMat = np.zeros((P, 9), dt)
    st.normal_opacity = np.zeros((P, 4), dt)
    st.rgb = np.zeros((P, 3), dt)
    st.clamped = np.zeros((P, 3), np.uint8)
    st.tiles_touched = np.zeros(P, np.uint32)
    st.point_offsets = np.zeros(P, np.uint32)
    st.out_color = np.zeros((3, H, W), dt)
    st.out_others = np.zeros((8, H, W), dt)
    st.final_T = np.zeros((3, H, W), dt)
    st.n_contrib = np.zeros((2, H, W), np.uint32)
    st.ranges = np.zeros((gx * gy, 2), np.uint32)
    if P == 0:
        st.out_color[:] = inp["bg"][:, None, None] * 0  # reference returns zeros for P == 0 (rasterize_points.cu:92,106)
        st.keys_unsorted = np.zeros(0, np.uint64); st.vals_unsorted = np.zeros(0, np.uint32)
        st.keys_sorted = np.zeros(0, np.uint64); st.point_list = np.zeros(0, np.uint32)
        return st
    cap = max(1024, 4 * P)
    while True:
        st.keys_unsorted = np.zeros(cap, np.uint64)
        st.vals_unsorted = np.zeros(cap, np.uint32)
        st.keys_sorted = np.zeros(cap, np.uint64)
        st.point_list = np.zeros(cap, np.uint32)
        Rn = fn(P, inp["sh_degree"], M, W, H, _ptr(inp["bg"]), _ptr(means3D), _ptr(inp["shs"]),
                _ptr(inp["colors_precomp"]), _ptr(inp["opacities"]), _ptr(inp["scales"]), _ptr(inp["rotations"]),
                _ptr(inp["transMat_precomp"]), _ptr(inp["viewmatrix"]), _ptr(inp["projmatrix"]), _ptr(inp["campos"]),
                creal(tanfovx), creal(tanfovy), _ptr(st.radii), _ptr(st.means2D), _ptr(st.depths), _ptr(st.transMat),
                _ptr(st.normal_opacity), _ptr(st.rgb), _ptr(st.clamped), _ptr(st.tiles_touched), _ptr(st.point_offsets),
                _ptr(st.out_color), _ptr(st.out_others), _ptr(st.final_T), _ptr(st.n_contrib), _ptr(st.ranges),
                ctypes.c_int64(cap), _ptr(st.keys_unsorted), _ptr(st.vals_unsorted), _ptr(st.keys_sorted),
                _ptr(st.point_list))
        if Rn <= cap:
            break
        cap = int(Rn)
    st.num_rendered = int(Rn)
    for k in ("keys_unsorted", "vals_unsorted", "keys_sorted", "point_list"):
        setattr(st, k, getattr(st, k)[:Rn])
    return st


def backward(st: ForwardState, dL_dout_color, dL_dout_others) -> dict:
    """Oracle of RasterizeGaussiansBackwardCUDA (DSR/rasterize_points.cu:143-240).

    Returns the reference's 8 gradient tensors plus dL_dnormal (internal) under the reference's names."""
    dt = st.dtype.type
    prec64 = st.dtype == np.float64
    fn = lib().orc_backward_f64 if prec64 else lib().orc_backward_f32
    creal = ctypes.c_double if prec64 else ctypes.c_float
    i = st.inputs
    P, M = st.P, i["M"]
    g = dict(dL_dmeans2D=np.zeros((P, 3), dt), dL_dnormal=np.zeros((P, 3), dt), dL_dopacity=np.zeros((P, 1), dt),
             dL_dcolors=np.zeros((P, 3), dt), dL_dmeans3D=np.zeros((P, 3), dt), dL_dtransMat=np.zeros((P, 9), dt),
             dL_dsh=np.zeros((P, M, 3), dt), dL_dscales=np.zeros((P, 2), dt), dL_drotations=np.zeros((P, 4), dt))
    if P == 0:
        return g
    dpix = np.ascontiguousarray(dL_dout_color, dtype=dt)
    doth = np.ascontiguousarray(dL_dout_others, dtype=dt)
    fn(P, i["sh_degree"], M, st.W, st.H, ctypes.c_int64(st.num_rendered), _ptr(i["bg"]), _ptr(i["means3D"]),
       _ptr(i["shs"]), _ptr(i["colors_precomp"]), _ptr(i["scales"]), _ptr(i["rotations"]), _ptr(i["transMat_precomp"]),
       _ptr(i["viewmatrix"]), _ptr(i["projmatrix"]), _ptr(i["campos"]), creal(i["tanfovx"]), creal(i["tanfovy"]),
       _ptr(st.radii), _ptr(st.means2D), _ptr(st.transMat), _ptr(st.normal_opacity), _ptr(st.rgb), _ptr(st.clamped),
       _ptr(st.final_T), _ptr(st.n_contrib), _ptr(st.ranges), _ptr(np.ascontiguousarray(st.point_list)),
       _ptr(dpix), _ptr(doth), _ptr(g["dL_dmeans2D"]), _ptr(g["dL_dnormal"]), _ptr(g["dL_dopacity"]),
       _ptr(g["dL_dcolors"]), _ptr(g["dL_dmeans3D"]), _ptr(g["dL_dtransMat"]), _ptr(g["dL_dsh"]),
       _ptr(g["dL_dscales"]), _ptr(g["dL_drotations"]))
    return g


def mark_visible(means3D, viewmatrix) -> np.ndarray:
    m = np.ascontiguousarray(means3D, np.float32)
    v = np.ascontiguousarray(viewmatrix, np.float32)
    out = np.zeros(m.shape[0], np.uint8)
    if m.shape[0]:
        lib().orc_mark_visible_f32(m.shape[0], _ptr(m), _ptr(v), _ptr(out))
    return out.astype(bool)

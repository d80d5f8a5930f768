"""Seeded synthetic "D-NeRF-shaped" scenes and cameras.

Pure numpy so the CPU oracle tests, the GPU parity tests and bench.py all draw exactly the same inputs.
Camera matrices follow the reference's conventions (scene/cameras.py:55-59, utils/graphics_utils.py:42-74):
``world_view_transform`` and ``full_proj_transform`` are stored transposed (row-vector convention).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

FOVX_DNERF = 0.6911112070083618  # camera_angle_x of the D-NeRF synthetic sets


@dataclass
class SyntheticCamera:
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: np.ndarray  # (4,4) f32, transposed
    projection_matrix: np.ndarray     # (4,4) f32, transposed
    full_proj_transform: np.ndarray   # (4,4) f32
    camera_center: np.ndarray         # (3,) f32
    fid: float

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """utils/graphics_utils.py:51-74 (z_sign = +1), returned NON-transposed."""
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    bottom, left = -top, -right
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(position, width: int, height: int, fovx: float = FOVX_DNERF, fid: float = 0.0,
                   target=(0.0, 0.0, 0.0), znear: float = 0.01, zfar: float = 100.0) -> SyntheticCamera:
    c = np.asarray(position, np.float64)
    z = np.asarray(target, np.float64) - c
    z /= np.linalg.norm(z)
    up = np.array([0.0, 0.0, 1.0])
    if abs(np.dot(up, z)) > 0.999:
        up = np.array([0.0, 1.0, 0.0])
    x = np.cross(up, z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    Rc2w = np.stack([x, y, z], axis=1)
    Rt = np.eye(4)
    Rt[:3, :3] = Rc2w.T
    Rt[:3, 3] = -Rc2w.T @ c
    wv = np.float32(Rt).T.copy()
    focal = width / (2 * math.tan(fovx / 2))
    fovy = 2 * math.atan(height / (2 * focal))
    proj = projection_matrix(znear, zfar, fovx, fovy).T.copy()
    full = (wv @ proj).astype(np.float32)
    center = np.linalg.inv(wv.astype(np.float64))[3, :3].astype(np.float32)
    return SyntheticCamera(width, height, fovx, fovy, wv, proj, full, center, float(fid))


def fibonacci_cameras(n: int, width: int, height: int, radius: float = 4.0, fovx: float = FOVX_DNERF):
    cams = []
    golden = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(n):
        zc = 1.0 - 2.0 * (i + 0.5) / n
        r = math.sqrt(max(0.0, 1.0 - zc * zc))
        th = golden * i
        pos = radius * np.array([math.cos(th) * r, math.sin(th) * r, zc])
        cams.append(look_at_camera(pos, width, height, fovx, fid=i / n))
    return cams


@dataclass
class SyntheticScene:
    """Canonical surfel parameters in the reference's raw (pre-activation) parameterisation
    (scene/gaussian_model.py:170-177) plus control nodes (utils/time_utils.py:805-811, 899-915)."""
    xyz: np.ndarray            # (P,3)
    features_dc: np.ndarray    # (P,1,3)
    features_rest: np.ndarray  # (P,15,3)
    scaling: np.ndarray        # (P,2) log-scale
    rotation: np.ndarray       # (P,4) unnormalised (w,x,y,z)
    opacity: np.ndarray        # (P,1) logit
    feature: np.ndarray        # (P,hyper_dim)
    nodes: Optional[np.ndarray] = None        # (M,3+hyper_dim)
    node_radius: Optional[np.ndarray] = None  # (M,) log
    node_weight: Optional[np.ndarray] = None  # (M,1) logit
    sh_degree: int = 3

    @property
    def P(self) -> int:
        return self.xyz.shape[0]


def make_scene(P: int, seed: int, s_med: float = 0.005, n_nodes: int = 0, hyper_dim: int = 8,
               sh_degree: int = 3) -> SyntheticScene:
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(P, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyz = (d * rng.uniform(size=(P, 1)) ** (1.0 / 3.0)).astype(np.float32)
    scaling = (math.log(s_med) + 0.5 * rng.normal(size=(P, 2))).astype(np.float32)
    rotation = rng.normal(size=(P, 4)).astype(np.float32)
    u = rng.uniform(0.05, 0.95, size=(P, 1))
    opacity = np.log(u / (1 - u)).astype(np.float32)
    ncoef = (sh_degree + 1) ** 2
    f_dc = rng.normal(size=(P, 1, 3)).astype(np.float32)
    f_rest = (0.1 * rng.normal(size=(P, 15, 3))).astype(np.float32)
    if ncoef < 16:
        f_rest[:, ncoef - 1:, :] = 0
    feature = np.full((P, hyper_dim), -1e-2, np.float32)
    sc = SyntheticScene(xyz, f_dc, f_rest, scaling, rotation, opacity, feature, sh_degree=sh_degree)
    if n_nodes > 0:
        dn = rng.normal(size=(n_nodes, 3))
        dn /= np.linalg.norm(dn, axis=1, keepdims=True)
        nxyz = dn * rng.uniform(size=(n_nodes, 1)) ** (1.0 / 3.0)
        sc.nodes = np.concatenate([nxyz, np.full((n_nodes, hyper_dim), 1e-2)], axis=1).astype(np.float32)
        scene_range = float(xyz.max() - xyz.min())
        sc.node_radius = np.full((n_nodes,), math.log(0.1 * scene_range + 1e-7), np.float32)
        sc.node_weight = np.zeros((n_nodes, 1), np.float32)
    return sc


def activated(sc: SyntheticScene):
    """The render() glue of gaussian_renderer/__init__.py:83-122 with zero deformation, in numpy:
    means3D, opacity=sigmoid, scales=exp, rotations=normalize, shs=cat(dc,rest)."""
    opac = 1.0 / (1.0 + np.exp(-sc.opacity.astype(np.float64)))
    scales = np.exp(sc.scaling.astype(np.float64))
    q = sc.rotation.astype(np.float64)
    q = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
    shs = np.concatenate([sc.features_dc, sc.features_rest], axis=1)
    return dict(means3D=sc.xyz.copy(), opacities=opac.astype(np.float32), scales=scales.astype(np.float32),
                rotations=q.astype(np.float32), shs=shs.astype(np.float32))


# BASELINE.json configs.  C1 is the CPU-runnable case, C3 the headline.
CONFIGS = {
    "C1": dict(P=10_000, W=256, H=256, s_med=0.02, n_nodes=0, K=0, seed=1235),
    "C2": dict(P=100_000, W=800, H=800, s_med=0.005, n_nodes=0, K=0, seed=1236),
    "C3": dict(P=300_000, W=800, H=800, s_med=0.005, n_nodes=512, K=4, seed=1237),
    "C4": dict(P=200_000, W=800, H=800, s_med=0.008, n_nodes=512, K=4, seed=1238),
    "C5": dict(P=1_000_000, W=1600, H=1600, s_med=0.005, n_nodes=2048, K=4, seed=1239),
    # small cases for parity tests
    "T0": dict(P=2_000, W=128, H=96, s_med=0.03, n_nodes=64, K=4, seed=7),
    "T1": dict(P=20_000, W=320, H=240, s_med=0.01, n_nodes=128, K=4, seed=11),
}

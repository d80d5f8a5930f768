"""Minimal containers with the reference's attribute names, so render()/DeformModel.step() can be driven without
the reference's data loaders (scene/gaussian_model.py:37-130, scene/cameras.py:18-59).  The real
``scene.GaussianModel`` / ``scene.cameras.Camera`` objects work with render() unchanged."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import synthetic as syn


class SurfelModel(nn.Module):
    """Canonical 2-D Gaussian parameters in the reference's raw parameterisation (gaussian_model.py:170-177)."""

    def __init__(self, scene: syn.SyntheticScene, device="cuda", with_motion_mask: bool = False):
        super().__init__()
        t = lambda a: nn.Parameter(torch.as_tensor(a, dtype=torch.float32, device=device).contiguous())
        self._xyz = t(scene.xyz)
        self._features_dc = t(scene.features_dc)
        self._features_rest = t(scene.features_rest)
        self._scaling = t(scene.scaling)
        self._rotation = t(scene.rotation)
        self._opacity = t(scene.opacity)
        self.feature = t(scene.feature)
        self.max_sh_degree = 3
        self.active_sh_degree = scene.sh_degree
        self.with_motion_mask = with_motion_mask
        self.max_radii2D = torch.zeros((self._xyz.shape[0],), device=device)

    @property
    def motion_mask(self):
        if self.with_motion_mask:
            return torch.sigmoid(self.feature[..., -1:])
        return torch.ones_like(self._xyz[..., :1])

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    def get_rotation_bias(self, rotation_bias=None):
        rotation_bias = rotation_bias if rotation_bias is not None else 0.
        return torch.nn.functional.normalize(self._rotation + rotation_bias)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    def raster_parameters(self):
        return [self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling, self._rotation, self.feature]


class ViewCamera:
    """Device-resident camera with the attributes render() reads (scene/cameras.py:18-59)."""

    def __init__(self, cam: syn.SyntheticCamera, device="cuda", uid: int = 0):
        self.uid = uid
        self.image_width, self.image_height = cam.image_width, cam.image_height
        self.FoVx, self.FoVy = cam.FoVx, cam.FoVy
        t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device)
        self.world_view_transform = t(cam.world_view_transform)
        self.projection_matrix = t(cam.projection_matrix)
        self.full_proj_transform = t(cam.full_proj_transform)
        self.camera_center = t(cam.camera_center)
        self.fid = t(np.array([cam.fid], np.float32))
        self.zfar, self.znear = 100.0, 0.01


class PipelineParams:
    """arguments/__init__.py:90-96."""
    convert_SHs_python = False
    compute_cov3D_python = False
    depth_ratio = 0.0
    debug = False

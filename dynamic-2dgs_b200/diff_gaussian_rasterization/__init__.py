"""Drop-in replacement for the reference package ``diff_gaussian_rasterization`` — the depth/alpha variant of the 3-D
Gaussian rasterizer (DGR/diff_gaussian_rasterization/__init__.py, DGR = submodules/diff-gaussian-rasterization): same names,
signatures, return values ``(color, radii, depth, alpha)`` and error behaviour, backed by the sm_100a kernels of libd2gs.so.

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

is the import the reference keeps commented out at gaussian_renderer/__init__.py:15 (render_flow, :222-337, needs it).
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from d2gs_b200 import gs3d as _gs3d
from d2gs_b200 import raster as _raster


def cpu_deep_copy_tuple(input_tuple):
    return _raster.cpu_deep_copy_tuple(input_tuple)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    """Reference: DGR/diff_gaussian_rasterization/__init__.py:20-41.  Returns (color, radii, depth, alpha)."""
    return _gs3d.rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)


class GaussianRasterizationSettings(NamedTuple):
    """Reference: DGR/diff_gaussian_rasterization/__init__.py:159-171 (field order is part of the API)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    """Reference: DGR/diff_gaussian_rasterization/__init__.py:173-224."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _raster.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   raster_settings)

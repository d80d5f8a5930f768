"""Drop-in replacement for the reference package ``diff_surfel_rasterization``
(DSR/diff_surfel_rasterization/__init__.py): same names, signatures, return values and error behaviour,
backed by the hand-written sm_100a kernels of libd2gs.so instead of the pybind ``_C`` module.

    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer

Put ``dynamic-2dgs_b200/`` on ``sys.path`` ahead of the reference's site-packages install and
``gaussian_renderer/__init__.py:14`` picks this module up unchanged.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from d2gs_b200 import raster as _raster


def cpu_deep_copy_tuple(input_tuple):
    return _raster.cpu_deep_copy_tuple(input_tuple)


def rasterize_gaussians(
    means3D,
    means2D,
    sh,
    colors_precomp,
    opacities,
    scales,
    rotations,
    cov3Ds_precomp,
    raster_settings,
):
    """Reference: DSR/diff_surfel_rasterization/__init__.py:21-42. Returns (color, radii, allmap)."""
    return _raster.rasterize_surfels(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizationSettings(NamedTuple):
    """Reference: DSR/diff_surfel_rasterization/__init__.py:158-170 (field order is part of the API)."""
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
    """Reference: DSR/diff_surfel_rasterization/__init__.py:172-222."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # boolean per point: in front of the near plane of this camera
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

"""TEST / BASELINE INFRASTRUCTURE — the reference's per-frame pipeline, restated in eager torch around the
UNMODIFIED reference rasterizer extension (oracle/_ref).

Only tests/ and bench.py (``--impl reference``) may import this.  The reference's own Python modules cannot be
imported here or on the GPU box (train_gui.py / scene / utils pull in pytorch3d, plyfile, simple_knn, ... which
are not installed; SURVEY.md header), so their *torch-level op sequence* is restated op for op:
  * ControlNodeWarp.forward  utils/time_utils.py:1133-1233  -> oracle/deform_oracle.py (knn_points -> explicit topk)
  * render() glue            gaussian_renderer/__init__.py:41-219, utils/point_utils.py:9-38
  * rasterizer               the reference CUDA extension itself, through its own GaussianRasterizer module
Runs on whatever device its tensors live on (CUDA for the reference arm of the bench, CPU impossible because the
reference rasterizer asserts CUDA tensors: rasterize_points.cu:27-28).
"""
from __future__ import annotations

import math

import torch

from . import deform_oracle as do


def depths_to_points(view, depthmap):
    c2w = (view.world_view_transform.T).inverse()
    W, H = view.image_width, view.image_height
    fx = W / (2 * math.tan(view.FoVx / 2.))
    fy = H / (2 * math.tan(view.FoVy / 2.))
    dev = depthmap.device
    intrins = torch.tensor([[fx, 0., W / 2.], [0., fy, H / 2.], [0., 0., 1.0]]).float().to(dev)
    grid_x, grid_y = torch.meshgrid(torch.arange(W), torch.arange(H), indexing='xy')
    points = torch.stack([grid_x, grid_y, torch.ones_like(grid_x)], dim=-1).reshape(-1, 3).float().to(dev)
    rays_d = points @ intrins.inverse().T @ c2w[:3, :3].T
    rays_o = c2w[:3, 3]
    return depthmap.reshape(-1, 1) * rays_d + rays_o


def depth_to_normal(view, depth):
    points = depths_to_points(view, depth).reshape(*depth.shape[1:], 3)
    output = torch.zeros_like(points)
    dx = torch.cat([points[2:, 1:-1] - points[:-2, 1:-1]], dim=0)
    dy = torch.cat([points[1:-1, 2:] - points[1:-1, :-2]], dim=1)
    normal_map = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    output[1:-1, 1:-1, :] = normal_map
    return output, points


def render_reference(ref_mod, view, pc, bg_color, d_xyz, d_rotation, d_scaling, debug=False):
    """The default branch of the reference render() (no d_color/d_opacity, no detach flags, no depth filtering)."""
    xyz = pc.get_xyz
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    tanfovx, tanfovy = math.tan(view.FoVx * 0.5), math.tan(view.FoVy * 0.5)
    rs = ref_mod.GaussianRasterizationSettings(
        image_height=int(view.image_height), image_width=int(view.image_width), tanfovx=tanfovx, tanfovy=tanfovy,
        bg=bg_color, scale_modifier=1.0, viewmatrix=view.world_view_transform, projmatrix=view.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=view.camera_center, prefiltered=False, debug=debug)
    rasterizer = ref_mod.GaussianRasterizer(raster_settings=rs)
    means3D = xyz + d_xyz
    opacity = pc.get_opacity
    scales = pc.get_scaling + d_scaling
    rotations = pc.get_rotation_bias(d_rotation)
    shs = pc.get_features
    rendered_image, radii, allmap = rasterizer(means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=None,
                                               opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=None)
    rets = {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}
    mask = 1
    render_alpha = allmap[1:2]
    render_normal = allmap[2:5]
    render_normal = (render_normal.permute(1, 2, 0) @ (view.world_view_transform[:3, :3].T)).permute(2, 0, 1)
    render_normal = render_normal * mask
    render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
    render_depth_expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
    render_dist = allmap[6:7] * mask
    depth_ratio = 1
    surf_depth = render_depth_expected * (1 - depth_ratio) + depth_ratio * render_depth_median
    surf_depth = surf_depth * mask
    surf_normal, surf_point = depth_to_normal(view, surf_depth)
    surf_normal = surf_normal.permute(2, 0, 1)
    surf_point = surf_point.permute(2, 0, 1)
    surf_normal = surf_normal * render_alpha.detach()
    surf_normal = surf_normal * mask
    rets.update({'alpha': render_alpha, 'rend_normal': render_normal, 'rend_dist': render_dist, 'depth': surf_depth,
                 'surf_normal': surf_normal, 'surf_point': surf_point, "bg_color": bg_color})
    return rets


def deform_reference(net_params, nodes, node_radius_log, node_weight_logit, xyz, fid, feature, motion_mask, K, hyper_dim,
                     local_frame=True, knn_mode="mm"):
    """ControlNodeWarp.forward in eager torch (time_utils.py:1133-1233); ``fid``: 0-d or (1,) time tensor."""
    t = fid.reshape(1, 1).expand(nodes.shape[0], 1)
    out = do.control_node_warp_forward(net_params, nodes, node_radius_log, node_weight_logit, xyz, t, feature, motion_mask,
                                       K, hyper_dim, local_frame=local_frame, knn_mode=knn_mode)
    return {"d_xyz": out["d_xyz"], "d_rotation": out["d_rotation"], "d_scaling": out["d_scaling"]}

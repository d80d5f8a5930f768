"""Drop-in replacement for the reference's ``gaussian_renderer`` package (gaussian_renderer/__init__.py):
``render`` (:41-219) and ``render_flow`` (:222-337) with the same signatures and returned dicts, backed by the B200
rasterizers; ``network_gui`` (gaussian_renderer/network_gui.py) and the ``GaussianModel`` re-export (:16) so that
``from gaussian_renderer import render, network_gui, render_flow`` (train_gui.py:18) and
``from gaussian_renderer import GaussianModel`` (render_mesh.py:22) resolve when this package shadows the reference's.
(Alternatively leave the reference's package in place and call ``d2gs_b200.install_into_reference()``, which swaps
``render`` / ``render_flow`` inside it.)

Differences that do not change results:
  * SH coefficients are handed to the rasterizer as the two parameter tensors (DC, rest) instead of a fresh
    P x 16 x 3 concatenation every frame (:114,122) whenever the model exposes ``_features_dc/_features_rest``;
  * the per-camera ray table of depth_to_normal (utils/point_utils.py:9-25: meshgrid + two 3x3 inverses rebuilt per
    call) is cached per camera pose.
"""
import math

import torch

from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from d2gs_b200 import raster as _raster
from d2gs_b200 import epilogue as _epilogue

_RAY_CACHE = {}
# True: exp / sigmoid / normalize and the deformation deltas are applied inside the per-surfel kernels whenever the model
# allows it (raw-parameter mode).  False: always the reference's eager op sequence in front of the rasterizer — the
# rasterizer then sees bit-identical inputs to the reference's, which the parity tests use to compare bit for bit.
FUSED_ACTIVATIONS = True


def __getattr__(name):
    """Lazy attributes of the reference package that live outside the hot path: ``GaussianModel`` is the reference's own
    class (scene/gaussian_model.py, importable whenever the trainer is), ``network_gui`` the viewer socket module."""
    if name == "GaussianModel":
        from scene.gaussian_model import GaussianModel
        return GaussianModel
    if name == "network_gui":
        import importlib
        return importlib.import_module(__name__ + ".network_gui")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def _plain_getters(pc) -> bool:
    """The fused raw-parameter path evaluates exp(_scaling), normalize(_rotation + d), sigmoid(_opacity) inside the kernel.
    That equals the model's getters only if the model has not overridden them: the reference's StandardGaussianModel
    (scene/gaussian_model.py) replaces get_scaling by an isotropic mean (and has `all_the_same`), so any object whose
    class redefines a getter relative to the base that owns `_scaling` — or that carries an `all_the_same` attribute —
    takes the eager op sequence instead."""
    if hasattr(pc, "all_the_same"):
        return False
    cls = type(pc)
    for getter in ("get_scaling", "get_opacity", "get_rotation_bias", "get_rotation", "get_xyz"):
        owners = [k for k in cls.__mro__ if getter in vars(k)]
        if len(owners) > 1:        # redefined somewhere down the hierarchy
            return False
    return True


def quaternion_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product a*b of (w,x,y,z) quaternions, returned with a non-negative real part
    (used only by the editing branch that passes ``d_rotation_bias``)."""
    aw, av = a[..., :1], a[..., 1:]
    bw, bv = b[..., :1], b[..., 1:]
    w = aw * bw - (av * bv).sum(-1, keepdim=True)
    v = aw * bv + bw * av + torch.cross(av, bv, dim=-1)
    q = torch.cat([w, v], -1)
    return torch.where(q[..., :1] < 0, -q, q)


def _camera_rays(view):
    """rays_d (H*W,3), rays_o (3): utils/point_utils.py:9-24, cached per camera pose."""
    key = (view.world_view_transform.data_ptr(), int(view.image_width), int(view.image_height), float(view.FoVx), float(view.FoVy),
           view.world_view_transform._version)
    hit = _RAY_CACHE.get(key)
    if hit is not None:
        return hit
    dev = view.world_view_transform.device
    c2w = (view.world_view_transform.T).inverse()
    W, H = int(view.image_width), int(view.image_height)
    fx = W / (2 * math.tan(view.FoVx / 2.))
    fy = H / (2 * math.tan(view.FoVy / 2.))
    intrins = torch.tensor([[fx, 0., W / 2.], [0., fy, H / 2.], [0., 0., 1.0]]).float().to(dev)
    grid_x, grid_y = torch.meshgrid(torch.arange(W, device=dev), torch.arange(H, device=dev), indexing='xy')
    points = torch.stack([grid_x, grid_y, torch.ones_like(grid_x)], dim=-1).reshape(-1, 3).float()
    rays_d = points @ intrins.inverse().T @ c2w[:3, :3].T
    rays_o = c2w[:3, 3]
    if len(_RAY_CACHE) > 512:
        _RAY_CACHE.clear()
    _RAY_CACHE[key] = (rays_d, rays_o)
    return rays_d, rays_o


def depths_to_points(view, depthmap):
    rays_d, rays_o = _camera_rays(view)
    return depthmap.reshape(-1, 1) * rays_d + rays_o


def depth_to_normal(view, depth):
    """utils/point_utils.py:27-38."""
    points = depths_to_points(view, depth).reshape(*depth.shape[1:], 3)
    output = torch.zeros_like(points)
    dx = points[2:, 1:-1] - points[:-2, 1:-1]
    dy = points[1:-1, 2:] - points[1:-1, :-2]
    normal_map = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    output[1:-1, 1:-1, :] = normal_map
    return output, points


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, d_xyz, d_rotation, d_scaling, d_opacity=None, d_color=None,
           scaling_modifier=1.0, override_color=None, random_bg_color=False, render_motion=False, detach_xyz=False,
           detach_scale=False, detach_rot=False, detach_opacity=False, d_rot_as_res=True, scale_const=None,
           d_rotation_bias=None, force_visible=False, depth_filtering=False):
    """Render the scene (reference: gaussian_renderer/__init__.py:41-219).  Background tensor must be on the GPU."""
    xyz = pc.get_xyz
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    bg = bg_color if not random_bg_color else torch.rand_like(bg_color)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False, debug=pipe.debug)

    means2D = screenspace_points

    if pipe.compute_cov3D_python:
        raise NotImplementedError("compute_cov3D_python is broken for 2-D scales in the reference (utils/general_utils.py:167)")

    # ---- colour inputs -----------------------------------------------------------------------------------------
    shs, sh_rest, colors_precomp = None, None, None
    if render_motion:
        colors_precomp = torch.zeros_like(xyz)
        colors_precomp[..., :1] = pc.motion_mask
        colors_precomp[..., -1:] = 1 - pc.motion_mask
    else:
        has_dc = d_color is not None and type(d_color) is not float
        split = hasattr(pc, "_features_dc") and hasattr(pc, "_features_rest") and not pipe.convert_SHs_python
        if split:
            shs = pc._features_dc + d_color[:, None] if has_dc else pc._features_dc
            sh_rest = pc._features_rest
        else:
            feats = pc.get_features
            sh_features = torch.cat([feats[:, :1] + d_color[:, None], feats[:, 1:]], dim=1) if has_dc else feats
            if pipe.convert_SHs_python:
                from d2gs_b200.sh import eval_sh
                shs_view = sh_features.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
                dir_pp = (xyz - viewpoint_camera.camera_center.repeat(sh_features.shape[0], 1))
                dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
                colors_precomp = torch.clamp_min(eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized) + 0.5, 0.0)
            else:
                shs = sh_features

    # ---- geometry inputs: fused path (activations + deltas inside the per-surfel kernel) when the model exposes the
    # ---- reference's raw parameters with its standard activations; otherwise the reference's eager op sequence.
    def _delta_ok(d, like):
        return (not torch.is_tensor(d) and d == 0) or (torch.is_tensor(d) and d.shape == like.shape)

    fused = (FUSED_ACTIVATIONS and all(hasattr(pc, n) for n in ("_scaling", "_rotation", "_opacity"))
             and getattr(pc, "scaling_activation", torch.exp) is torch.exp
             and getattr(pc, "opacity_activation", torch.sigmoid) is torch.sigmoid
             and getattr(pc, "rotation_activation", torch.nn.functional.normalize) is torch.nn.functional.normalize
             and scale_const is None and d_opacity is None and d_rotation_bias is None
             and _delta_ok(d_xyz, xyz) and _delta_ok(d_scaling, pc._scaling) and _delta_ok(d_rotation, pc._rotation)
             and pc._scaling.dim() == 2 and pc._scaling.shape[1] == 2 and _plain_getters(pc))
    if fused:
        det = lambda t, flag: (t.detach() if flag and torch.is_tensor(t) else t)
        tz = lambda d: d if torch.is_tensor(d) else None
        rendered_image, radii, allmap = _raster.rasterize_surfels_raw(
            det(xyz, detach_xyz), det(tz(d_xyz), detach_xyz), det(pc._scaling, detach_scale), det(tz(d_scaling), detach_scale),
            det(pc._rotation, detach_rot), det(tz(d_rotation), detach_rot), det(pc._opacity, detach_opacity), means2D,
            shs, sh_rest, colors_precomp, raster_settings)
    else:
        means3D = xyz + d_xyz
        if scale_const is not None:
            opacity = torch.ones_like(pc.get_opacity)
        else:
            opacity = pc.get_opacity if d_opacity is None else pc.get_opacity + d_opacity
        scales = pc.get_scaling + d_scaling
        rotations = pc.get_rotation_bias(d_rotation)
        if d_rotation_bias is not None:
            rotations = quaternion_multiply(d_rotation_bias, rotations)
        if detach_xyz:
            means3D = means3D.detach()
        if detach_rot:
            rotations = rotations.detach()
        if detach_scale:
            scales = scales.detach()
        if detach_opacity:
            opacity = opacity.detach()
        if scale_const is not None:
            scales = scale_const * torch.ones_like(scales)
        rendered_image, radii, allmap = _raster.rasterize_surfels(means3D, means2D, shs, colors_precomp, opacity, scales,
                                                                  rotations, None, raster_settings, sh_rest=sh_rest)

    rets = {"render": rendered_image, "viewspace_points": means2D, "visibility_filter": radii > 0, "radii": radii}

    if not depth_filtering:
        # fused image-space epilogue: one kernel forward / one backward (d2gs_b200/epilogue.py)
        render_alpha, render_normal, render_dist, surf_depth, surf_normal, surf_point = _epilogue.render_epilogue(allmap, viewpoint_camera)
        pipe.depth_ratio = 1
    else:
        whitebackground = torch.tensor([1, 1, 1], dtype=torch.float32, device=xyz.device)
        if bg_color.equal(whitebackground):
            mask = (1 - (torch.all(rendered_image >= 0.95, dim=0)).to(torch.int))
        else:
            mask = (1 - (torch.all(rendered_image <= 0.05, dim=0)).to(torch.int))
        render_alpha = allmap[1:2]
        render_normal = allmap[2:5]
        render_normal = (render_normal.permute(1, 2, 0) @ (viewpoint_camera.world_view_transform[:3, :3].T)).permute(2, 0, 1)
        render_normal = render_normal * mask
        render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
        render_depth_expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
        render_dist = allmap[6:7] * mask
        pipe.depth_ratio = 1
        surf_depth = render_depth_expected * (1 - pipe.depth_ratio) + (pipe.depth_ratio) * render_depth_median
        surf_depth = surf_depth * mask
        surf_normal, surf_point = depth_to_normal(viewpoint_camera, surf_depth)
        surf_normal = surf_normal.permute(2, 0, 1)
        surf_point = surf_point.permute(2, 0, 1)
        surf_normal = surf_normal * (render_alpha).detach()
        surf_normal = surf_normal * mask

    rets.update({'alpha': render_alpha, 'rend_normal': render_normal, 'rend_dist': render_dist, 'depth': surf_depth,
                 'surf_normal': surf_normal, 'surf_point': surf_point, "bg_color": bg})
    return rets


def render_flow(pc, viewpoint_camera1, viewpoint_camera2, d_xyz1, d_xyz2, d_rotation1, d_scaling1, scaling_modifier=1.0,
                compute_cov3D_python=False, scale_const=None, d_rot_as_res=True, **kwargs):
    """Splat the per-surfel screen-space motion between (t1, camera1) and (t2, camera2): reference
    gaussian_renderer/__init__.py:222-337 (called by train_gui.py:351 for the optical-flow loss).

    The reference body unpacks four outputs from a rasterizer call — the interface of its depth/alpha 3-D Gaussian
    rasterizer (submodules/diff-gaussian-rasterization) — while the module imports the surfel rasterizer, which returns
    three.  Here the model decides: 3-column scales go to the 3-D drop-in (diff_gaussian_rasterization, 4 outputs); the
    2-column scales of a surfel model go to the surfel rasterizer and ``depth`` / ``alpha`` are planes 0 / 1 of its
    allmap (sum of w*depth, sum of w) — the same quantities the 3-D rasterizer returns."""
    xyz = pc.get_xyz
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    tanfovx = math.tan(viewpoint_camera1.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera1.FoVy * 0.5)

    # per-surfel flow in normalised device coordinates; centres are detached (:252), the deltas carry the gradient
    canon = xyz.detach()
    ones = torch.ones_like(canon[..., :1])
    proj2 = (viewpoint_camera2 if viewpoint_camera2 is not None else viewpoint_camera1).full_proj_transform
    uvz2 = torch.cat([canon + d_xyz2, ones], dim=-1) @ proj2
    uvz2 = uvz2[..., :3] / uvz2[..., -1:]
    uvz1 = torch.cat([canon + d_xyz1, ones], dim=-1) @ viewpoint_camera1.full_proj_transform
    uvz1 = uvz1[..., :3] / uvz1[..., -1:]
    flow = uvz2 - uvz1
    flow = torch.cat([flow[..., :2], pc.motion_mask.expand_as(flow[..., -1:])], dim=-1)   # third channel: motion mask (:266)

    means3D = xyz + d_xyz1
    opacity = pc.get_opacity
    base_rot = pc.get_rotation
    if d_rot_as_res:
        rotations = base_rot + d_rotation1
    else:
        rotations = base_rot if type(d_rotation1) is float else quaternion_multiply(d_rotation1, base_rot)
    scales = torch.ones_like(pc.get_scaling) * scale_const if scale_const is not None else pc.get_scaling + d_scaling1
    if compute_cov3D_python and scale_const is None:
        raise NotImplementedError("compute_cov3D_python is not supported by the B200 drop-in (use the scales / rotations path)")
    settings = dict(image_height=int(viewpoint_camera1.image_height), image_width=int(viewpoint_camera1.image_width),
                    tanfovx=tanfovx, tanfovy=tanfovy, bg=torch.zeros_like(flow[0]), scale_modifier=scaling_modifier,
                    viewmatrix=viewpoint_camera1.world_view_transform, projmatrix=viewpoint_camera1.full_proj_transform,
                    sh_degree=0, campos=viewpoint_camera1.camera_center, prefiltered=False, debug=False)
    if scales.shape[-1] == 3:
        import diff_gaussian_rasterization as dgr
        rendered_image, radii, rendered_depth, rendered_alpha = dgr.GaussianRasterizer(dgr.GaussianRasterizationSettings(**settings))(
            means3D=means3D, means2D=screenspace_points, shs=None, colors_precomp=flow, opacities=opacity, scales=scales,
            rotations=rotations, cov3D_precomp=None)
    else:
        rendered_image, radii, allmap = GaussianRasterizer(GaussianRasterizationSettings(**settings))(
            means3D=means3D, means2D=screenspace_points, shs=None, colors_precomp=flow, opacities=opacity, scales=scales,
            rotations=rotations, cov3D_precomp=None)
        rendered_depth, rendered_alpha = allmap[0:1], allmap[1:2]
    return {"render": rendered_image, "depth": rendered_depth, "alpha": rendered_alpha, "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii}

"""TEST INFRASTRUCTURE — CPU restatement (torch, float32 or float64) of the reference's depth/alpha 3-D Gaussian rasterizer,
DGR = /root/reference/submodules/diff-gaussian-rasterization.  Never imported by the product path.

Forward follows DGR/cuda_rasterizer/forward.cu:74-112 (computeCov2D), :117-150 (computeCov3D, quaternion used as given),
:153-264 (preprocessCUDA: near-plane cull at 0.2, conic, radius = ceil(3 sqrt(lambda_max)), tile rectangle), :270-376
(renderCUDA: power > 0 / alpha < 1/255 rejections, alpha capped at 0.99, early termination at T < 1e-4, outputs colour,
depth = sum w d, alpha = sum w, n_contrib), auxiliary.h:45-62 (ndc2Pix in double, getRect) and rasterizer_impl.cu:70-110
(per-tile lists ordered by (tile, depth bits), ties by index: CUB's radix sort is stable).

Backward = torch autograd of that forward, with the places where the reference's hand-written gradient is NOT the
derivative of its forward restated explicitly so that the oracle reproduces the reference, not calculus:
  * the scale gradient is taken w.r.t. scale_modifier * scale (backward.cu:313-317 omits the factor) -> straight-through;
  * alpha = min(0.99, o G): the reference differentiates as if the cap were absent (backward.cu:503-551) -> straight-through;
  * the clamp of the view-space mean to 1.3 tan(fov) (forward.cu:82-87): gradient 0 through the clamped coordinate and no
    dependence of the clamped value on t.z (backward.cu:171-172, 253-255) -> `_ClampedT`;
  * the conic = inverse covariance: 1 / (det^2 + 1e-7) instead of 1 / det^2 (backward.cu:201-211) -> `_Conic`.
Everything else in DGR/cuda_rasterizer/backward.cu (colour/SH incl. the direction normalisation and the clamp-at-0 rule,
depth, projected mean, covariance -> scale / quaternion) is the derivative of the forward.

The image is processed Gaussian by Gaussian in depth order, vectorised over all pixels: sizes up to a few thousand Gaussians
and ~10^4 pixels finish in seconds.  Pin: tests/golden/gs3d_golden_*.npz hold outputs, intermediates and all eight gradients
of the UNMODIFIED reference extension (oracle/_ref/diff_gaussian_rasterization, built by oracle/build_ref_aux.sh) run on a
B200 by tests/golden/make_gs3d_golden.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]
TILE = 16


class _Conic(torch.autograd.Function):
    """(a, b, c) -> (c, -b, a) / det  (forward.cu:225-230); backward as written in backward.cu:196-211."""

    @staticmethod
    def forward(ctx, a, b, c):
        det = a * c - b * b
        ctx.save_for_backward(a, b, c)
        inv = 1.0 / det
        return c * inv, -b * inv, a * inv

    @staticmethod
    def backward(ctx, gx, gy_true, gz):
        a, b, c = ctx.saved_tensors
        gy = 0.5 * gy_true                       # the reference accumulates half of d/dconic.y (backward.cu:546)
        denom = a * c - b * b
        d2 = 1.0 / (denom * denom + 0.0000001)
        da = d2 * (-c * c * gx + 2 * b * c * gy + (denom - a * c) * gz)
        dc = d2 * (-a * a * gz + 2 * a * b * gy + (denom - a * c) * gx)
        db = d2 * 2 * (b * c * gx - (denom + 2 * b * b) * gy + a * b * gz)
        return da, db, dc


class _ClampedT(torch.autograd.Function):
    """t.x <- clamp(t.x / t.z, +-lim) * t.z  with the reference's gradient: 1 inside the clamp, 0 outside, none to t.z."""

    @staticmethod
    def forward(ctx, tx, tz, lim):
        r = tx / tz
        ctx.save_for_backward((r >= -lim) & (r <= lim))
        return torch.clamp(r, -lim, lim) * tz

    @staticmethod
    def backward(ctx, g):
        (inside,) = ctx.saved_tensors
        return g * inside.to(g.dtype), None, None


def eval_sh(deg, sh, dirs):
    """forward.cu:20-71 without the +0.5 / clamp.  sh: (n, M, 3), dirs: (n, 3) unit."""
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6] +
                   SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3.0 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10] +
                       SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12] +
                       SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14] +
                       SH_C3[6] * x * (xx - 3.0 * yy) * sh[:, 15])
    return res


def cov3d_from_scale_rot(scales, mod, q):
    """forward.cu:117-150.  Rm is GLM's column-major literal read as a math matrix; Sigma = (S Rm)^T (S Rm)."""
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    Rm = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y)], -1),
        torch.stack([2 * (x * y - r * z), 1 - 2 * (x * x + z * z), 2 * (y * z + r * x)], -1),
        torch.stack([2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)], -1)], 1)
    # backward.cu:313-317 returns dL/d(mod * scale) as the scale gradient (no factor `mod`): value mod*s, gradient 1
    s_eff = scales + (mod * scales - scales).detach()
    Mm = s_eff[:, :, None] * Rm
    Sg = Mm.transpose(1, 2) @ Mm
    return torch.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], -1)


def render(means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, H, W, bg, *, shs=None, sh_degree=0,
           colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0, means2D=None,
           dtype=torch.float64):
    """Returns dict(color (3,H,W), depth (1,H,W), alpha (1,H,W), radii (P) int, tiles_touched (P), n_contrib (H,W),
    num_rendered, means2D_pix (P,2), depths (P), conic_opacity (P,4), rgb (P,3)).  Differentiable w.r.t. every float input
    (and ``means2D``, a (P,3) zero tensor whose gradient is the reference's dL_dmeans2D in NDC units)."""
    t = lambda a: None if a is None else (a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))).to(dtype)
    means3D, opacities, view, proj, campos, bg = t(means3D), t(opacities).reshape(-1), t(viewmatrix), t(projmatrix), t(campos), t(bg)
    shs, colors_precomp, scales, rotations, cov3D_precomp = t(shs), t(colors_precomp), t(scales), t(rotations), t(cov3D_precomp)
    P = means3D.shape[0]
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    focal_y, focal_x = H / (2.0 * tanfovy), W / (2.0 * tanfovx)
    if dtype == torch.float32:   # rasterizer_impl.cu:216-217 computes the focal lengths in float
        focal_y = float(np.float32(H) / (np.float32(2.0) * np.float32(tanfovy)))
        focal_x = float(np.float32(W) / (np.float32(2.0) * np.float32(tanfovx)))
    radii = torch.zeros(P, dtype=torch.int64)
    tiles_touched = torch.zeros(P, dtype=torch.int64)
    out = dict(radii=radii, tiles_touched=tiles_touched)

    p_view_all = means3D @ view[:3, :3] + view[3, :3]
    vis = torch.nonzero(p_view_all[:, 2].detach() > 0.2).reshape(-1)          # auxiliary.h:155 (near-plane cull)
    m = means3D[vis]
    p_view = p_view_all[vis]
    hom = torch.cat([m, torch.ones_like(m[:, :1])], 1) @ proj
    p_w = 1.0 / (hom[:, 3] + 0.0000001)
    ndc = hom[:, :2] * p_w[:, None]
    if means2D is not None:
        ndc = ndc + t(means2D)[vis, :2]
    cov3D = cov3D_precomp[vis] if cov3D_precomp is not None else cov3d_from_scale_rot(scales[vis], scale_modifier, rotations[vis])
    # computeCov2D
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tz = p_view[:, 2]
    tx = _ClampedT.apply(p_view[:, 0], tz, limx)
    ty = _ClampedT.apply(p_view[:, 1], tz, limy)
    zero = torch.zeros_like(tz)
    Jm = torch.stack([torch.stack([focal_x / tz, zero, zero], -1), torch.stack([zero, focal_y / tz, zero], -1),
                      torch.stack([-(focal_x * tx) / (tz * tz), -(focal_y * ty) / (tz * tz), zero], -1)], 1)
    Tm = view[:3, :3] @ Jm
    V = torch.stack([torch.stack([cov3D[:, 0], cov3D[:, 1], cov3D[:, 2]], -1), torch.stack([cov3D[:, 1], cov3D[:, 3], cov3D[:, 4]], -1),
                     torch.stack([cov3D[:, 2], cov3D[:, 4], cov3D[:, 5]], -1)], 1)
    cov = Tm.transpose(1, 2) @ V.transpose(1, 2) @ Tm
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = (a * c - b * b).detach()
    con_x, con_y, con_z = _Conic.apply(a, b, c)
    mid = 0.5 * (a + c).detach()
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    lam2 = mid - torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    my_radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(lam, lam2)))
    # ndc2Pix is evaluated in double and rounded to float (auxiliary.h:45-48)
    pix = (((ndc.double() + 1.0) * torch.tensor([W, H], dtype=torch.float64) - 1.0) * 0.5).to(dtype)
    pd, rd = pix.detach(), my_radius
    trunc = lambda v: torch.trunc(v).to(torch.int64)
    rx0 = torch.clamp(trunc((pd[:, 0] - rd) / TILE), 0, gx); ry0 = torch.clamp(trunc((pd[:, 1] - rd) / TILE), 0, gy)
    rx1 = torch.clamp(trunc((pd[:, 0] + rd + TILE - 1) / TILE), 0, gx); ry1 = torch.clamp(trunc((pd[:, 1] + rd + TILE - 1) / TILE), 0, gy)
    area = (rx1 - rx0) * (ry1 - ry0)
    keep = (det != 0) & (area > 0)
    if colors_precomp is None:
        d = m - campos
        d = d / torch.sqrt((d * d).sum(-1, keepdim=True))
        res = eval_sh(sh_degree, shs[vis], d) + 0.5
        rgb = torch.clamp_min(res, 0.0)          # gradient 0 where clamped (backward.cu:33-36)
    else:
        rgb = colors_precomp[vis]
    radii[vis[keep]] = my_radius[keep].to(torch.int64)
    tiles_touched[vis[keep]] = area[keep]
    full = lambda v, shape: torch.zeros(shape, dtype=dtype).index_put((vis[keep],), v[keep].detach())
    out["means2D_pix"] = full(pix, (P, 2)); out["depths"] = full(p_view[:, 2], (P,))
    out["conic_opacity"] = full(torch.stack([con_x, con_y, con_z, opacities[vis]], -1), (P, 4)); out["rgb"] = full(rgb, (P, 3))
    out["num_rendered"] = int(area[keep].sum())

    # depth order; ties by index (stable).  Keys are the float bits of positive depths: same order as the values.
    kept = torch.nonzero(keep).reshape(-1)
    depth_key = p_view[kept, 2].detach().to(torch.float32)
    order = kept[torch.sort(depth_key, stable=True).indices]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pfx, pfy = xs.to(dtype), ys.to(dtype)
    tile_x, tile_y = xs // TILE, ys // TILE
    T = torch.ones((H, W), dtype=dtype)
    C = torch.zeros((3, H, W), dtype=dtype)
    D = torch.zeros((H, W), dtype=dtype)
    A = torch.zeros((H, W), dtype=dtype)
    done = torch.zeros((H, W), dtype=torch.bool)
    count = torch.zeros((H, W), dtype=torch.int64)
    last = torch.zeros((H, W), dtype=torch.int64)
    opa = opacities[vis]
    for g in order.tolist():
        x0, x1, y0, y1 = int(rx0[g]) * TILE, min(int(rx1[g]) * TILE, W), int(ry0[g]) * TILE, min(int(ry1[g]) * TILE, H)
        sl = (slice(y0, y1), slice(x0, x1))
        act = ~done[sl]
        count[sl] += act.to(torch.int64)
        dx, dy = pix[g, 0] - pfx[sl], pix[g, 1] - pfy[sl]
        power = -0.5 * (con_x[g] * dx * dx + con_z[g] * dy * dy) - con_y[g] * dx * dy
        raw = opa[g] * torch.exp(power)
        alpha = raw + (torch.clamp_max(raw, 0.99) - raw).detach()
        ok = act & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
        test_T = T[sl] * (1 - alpha)
        newly_done = ok & (test_T.detach() < 0.0001)
        contrib = ok & ~newly_done
        w = torch.where(contrib, alpha * T[sl], torch.zeros_like(alpha))
        Cn, Dn, An, Tn = C.clone(), D.clone(), A.clone(), T.clone()
        Cn[(slice(None),) + sl] = C[(slice(None),) + sl] + w[None] * rgb[g][:, None, None]
        Dn[sl] = D[sl] + w * p_view[g, 2]
        An[sl] = A[sl] + w
        Tn[sl] = torch.where(contrib, test_T, T[sl])
        C, D, A, T = Cn, Dn, An, Tn
        done[sl] |= newly_done
        last[sl] = torch.where(contrib, count[sl], last[sl])
    out["color"] = C + T[None] * bg[:, None, None]
    out["depth"] = D[None]
    out["alpha"] = A[None]
    out["n_contrib"] = last
    return out


def render_with_grads(inputs: dict, g_color, g_depth, g_alpha, dtype=torch.float64):
    """inputs: the keyword arguments of render() as numpy arrays.  Returns (forward outputs, gradients named like the reference's
    return tuple: means2D, colors_precomp, opacities, means3D, cov3Ds_precomp, sh, scales, rotations)."""
    leaf = {}
    kw = dict(inputs)
    for k in ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp"):
        if kw.get(k) is not None:
            leaf[k] = torch.as_tensor(np.asarray(kw[k])).to(dtype).clone().requires_grad_(True)
            kw[k] = leaf[k]
    leaf["means2D"] = torch.zeros((leaf["means3D"].shape[0], 3), dtype=dtype, requires_grad=True)
    out = render(means2D=leaf["means2D"], dtype=dtype, **kw)
    tt = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)
    loss = (out["color"] * tt(g_color)).sum() + (out["depth"] * tt(g_depth)).sum() + (out["alpha"] * tt(g_alpha)).sum()
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).detach() for k, v in leaf.items()}
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    return out, grads

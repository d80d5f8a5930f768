"""GPU: the fused photometric loss (libd2gs.so: d2gs_loss_forward/backward through d2gs_b200.loss) against the reference's
golden vectors and, at the benchmark's image size, against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import util
from oracle import loss_oracle as lo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.npz")


@pytest.mark.parametrize("case", ["a", "b"])
def test_fused_loss_matches_reference_golden(case, cuda_device):
    from d2gs_b200 import loss as fl
    dev = cuda_device
    g = np.load(GOLD)
    T = lambda k: torch.tensor(g[f"{case}_{k}"], device=dev)
    img, rn, sn, rd = (T(k).requires_grad_(True) for k in ("image", "rend_normal", "surf_normal", "rend_dist"))
    lam, ln, ld = (float(v) for v in g[f"{case}_lambdas"])
    loss, parts = fl.surfel_loss(img, T("gt"), rn, sn, rd, lam, ln, ld, return_parts=True)
    (3.0 * loss).backward()                      # exercises the upstream-gradient scalar
    parts = parts.cpu().numpy()
    for i, k in enumerate(("loss", "l1", "ssim", "normal", "dist")):
        ref = float(g[f"{case}_{k}"])
        assert abs(float(parts[i]) - ref) <= 1e-5 * max(abs(ref), 1e-3), (k, float(parts[i]), ref)      # fp32 sums, different order
    for t, k in ((img, "g_image"), (rn, "g_rend_normal"), (sn, "g_surf_normal"), (rd, "g_rend_dist")):
        assert util.rel_err(t.grad.cpu().numpy() / 3.0, g[f"{case}_{k}"]) < 1e-4, (k, util.rel_err(t.grad.cpu().numpy() / 3.0, g[f"{case}_{k}"]))


def test_fused_loss_at_benchmark_size_and_wrappers(cuda_device):
    from d2gs_b200 import loss as fl
    dev = cuda_device
    H = W = 800
    gen = torch.Generator().manual_seed(11)
    img = torch.rand(3, H, W, generator=gen)
    gt = (img + 0.1 * torch.randn(3, H, W, generator=gen)).clamp(0, 1)
    rn = torch.nn.functional.normalize(torch.randn(3, H, W, generator=gen), dim=0)
    sn = torch.nn.functional.normalize(torch.randn(3, H, W, generator=gen), dim=0) * torch.rand(1, H, W, generator=gen)
    rd = torch.rand(1, H, W, generator=gen) * 1e-3
    cpu = [t.clone().requires_grad_(True) for t in (img, rn, sn, rd)]
    l_ref, _ = lo.surfel_loss(cpu[0], gt, cpu[1], cpu[2], cpu[3], 0.2, 0.02, 1000.0)
    l_ref.backward()
    gpu = [t.clone().to(dev).requires_grad_(True) for t in (img, rn, sn, rd)]
    l_our = fl.surfel_loss(gpu[0], gt.to(dev), gpu[1], gpu[2], gpu[3], 0.2, 0.02, 1000.0)
    l_our.backward()
    assert abs(float(l_our) - float(l_ref)) <= 1e-5 * abs(float(l_ref))
    for a, b, k in zip(gpu, cpu, ("image", "rend_normal", "surf_normal", "rend_dist")):
        assert util.rel_err(a.grad.cpu().numpy(), b.grad.numpy()) < 1e-4, k
    # reference-named wrappers; no normal / distortion maps; inputs that need no gradient
    x, y = img.to(dev), gt.to(dev)
    assert abs(float(fl.l1_loss(x, y)) - float(lo.l1_loss(img, gt))) < 1e-6
    assert abs(float(fl.ssim(x, y)) - float(lo.ssim(img, gt))) < 1e-5
    xr = x.clone().requires_grad_(True)
    fl.surfel_loss(xr, y, None, None, None, 0.2).backward()
    c = img.clone().requires_grad_(True)
    lo.surfel_loss(c, gt, lambda_dssim=0.2)[0].backward()
    assert util.rel_err(xr.grad.cpu().numpy(), c.grad.numpy()) < 1e-4
    with pytest.raises(RuntimeError):
        fl.surfel_loss(img, gt)            # CPU tensors: the product path has no CPU fallback

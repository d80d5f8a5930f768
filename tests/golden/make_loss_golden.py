"""Generates tests/golden/loss_golden.npz with the REFERENCE's own loss functions (utils/loss_utils.py, imported from
/root/reference in the build container; CPU tensors, so no GPU is needed).  Committed together with its output."""
import os, sys
import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from utils.loss_utils import l1_loss, ssim      # noqa: E402  (reference code, read-only)

HERE = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(20240925)
out = {}
for name, (H, W) in {"a": (37, 53), "b": (64, 96)}.items():
    img = torch.rand(3, H, W, generator=g, dtype=torch.float32).requires_grad_(True)
    gt = (img.detach() + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    if name == "b":
        gt[:, :10, :] = img.detach()[:, :10, :]          # exact matches: sign(0) = 0 in the L1 gradient
    rn = torch.nn.functional.normalize(torch.randn(3, H, W, generator=g), dim=0).requires_grad_(True)
    sn = (torch.nn.functional.normalize(torch.randn(3, H, W, generator=g), dim=0) * torch.rand(1, H, W, generator=g)).requires_grad_(True)
    rd = (torch.rand(1, H, W, generator=g) * 1e-3).requires_grad_(True)
    lam, ln, ld = 0.2, 0.02, 1000.0
    Ll1 = l1_loss(img, gt)
    s = ssim(img, gt)
    normal_loss = ln * (1 - (rn * sn).sum(dim=0))[None].mean()       # train_gui.py:296-299
    dist_loss = ld * rd.mean()
    loss = (1.0 - lam) * Ll1 + lam * (1.0 - s) + normal_loss + dist_loss
    loss.backward()
    for k, v in dict(image=img, gt=gt, rend_normal=rn, surf_normal=sn, rend_dist=rd).items():
        out[f"{name}_{k}"] = v.detach().numpy()
    for k, v in dict(loss=loss, l1=Ll1, ssim=s, normal=normal_loss, dist=dist_loss).items():
        out[f"{name}_{k}"] = np.float32(v.item())
    for k, v in dict(g_image=img.grad, g_rend_normal=rn.grad, g_surf_normal=sn.grad, g_rend_dist=rd.grad).items():
        out[f"{name}_{k}"] = v.numpy()
    out[f"{name}_lambdas"] = np.array([lam, ln, ld], np.float32)
np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **out)
print("wrote", os.path.join(HERE, "loss_golden.npz"), {k: getattr(v, "shape", ()) for k, v in out.items() if k.startswith("a_")})

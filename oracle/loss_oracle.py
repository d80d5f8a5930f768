"""TEST INFRASTRUCTURE ONLY - CPU/torch restatement of the reference's photometric loss.

Follows utils/loss_utils.py:18-19 (l1_loss), :33-43 (Gaussian window), :45-76 (ssim/_ssim) and train_gui.py:292-313
(normal-consistency term, distortion term, composition).  Pinned against the reference's own functions by
tests/golden/make_loss_golden.py (run in the build container, where /root/reference is importable); nothing on the
product path may import this module.
"""
from math import exp

import torch
import torch.nn.functional as F


def gaussian_window(window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)], dtype=torch.float32)
    return g / g.sum()


def ssim_map(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    c = img1.shape[-3]
    w1 = gaussian_window(window_size).to(img1.dtype).unsqueeze(1)
    win = (w1 @ w1.t()).unsqueeze(0).unsqueeze(0).expand(c, 1, window_size, window_size).contiguous().to(img1.device)
    x, y = (img1[None] if img1.dim() == 3 else img1), (img2[None] if img2.dim() == 3 else img2)
    pad = window_size // 2
    mu1, mu2 = F.conv2d(x, win, padding=pad, groups=c), F.conv2d(y, win, padding=pad, groups=c)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(x * x, win, padding=pad, groups=c) - mu1_sq
    s2 = F.conv2d(y * y, win, padding=pad, groups=c) - mu2_sq
    s12 = F.conv2d(x * y, win, padding=pad, groups=c) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))


def ssim(img1, img2):
    return ssim_map(img1, img2).mean()


def l1_loss(a, b):
    return (a - b).abs().mean()


def surfel_loss(image, gt, rend_normal=None, surf_normal=None, rend_dist=None, lambda_dssim=0.2, lambda_normal=0.0, lambda_dist=0.0):
    """Returns (loss, dict of parts)."""
    l1 = l1_loss(image, gt)
    s = ssim(image, gt)
    loss = (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - s)
    parts = {"l1": l1, "ssim": s, "normal": torch.zeros(()), "dist": torch.zeros(())}
    if rend_normal is not None and surf_normal is not None and lambda_normal != 0.0:
        parts["normal"] = lambda_normal * (1 - (rend_normal * surf_normal).sum(dim=0))[None].mean()
        loss = loss + parts["normal"]
    if rend_dist is not None and lambda_dist != 0.0:
        parts["dist"] = lambda_dist * rend_dist.mean()
        loss = loss + parts["dist"]
    return loss, parts

"""CPU: the loss oracle (oracle/loss_oracle.py) against the golden vectors produced by the reference's own
utils/loss_utils.py + the train_gui.py composition (tests/golden/make_loss_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.npz")


@pytest.mark.parametrize("case", ["a", "b"])
def test_loss_oracle_matches_reference_golden(case):
    g = np.load(GOLD)
    T = lambda k: torch.tensor(g[f"{case}_{k}"])
    img, rn, sn, rd = (T(k).requires_grad_(True) for k in ("image", "rend_normal", "surf_normal", "rend_dist"))
    lam, ln, ld = (float(v) for v in g[f"{case}_lambdas"])
    loss, parts = lo.surfel_loss(img, T("gt"), rn, sn, rd, lam, ln, ld)
    loss.backward()
    assert abs(float(loss) - float(g[f"{case}_loss"])) <= 1e-6 * abs(float(g[f"{case}_loss"]))
    for k in ("l1", "ssim", "normal", "dist"):
        assert abs(float(parts[k]) - float(g[f"{case}_{k}"])) <= 2e-6 * max(abs(float(g[f"{case}_{k}"])), 1e-3), k
    for t, k in ((img, "g_image"), (rn, "g_rend_normal"), (sn, "g_surf_normal"), (rd, "g_rend_dist")):
        ref = g[f"{case}_{k}"]
        assert np.linalg.norm(t.grad.numpy() - ref) <= 1e-5 * np.linalg.norm(ref), k


def test_window_is_the_reference_window():
    w = lo.gaussian_window()
    assert w.shape == (11,) and abs(float(w.sum()) - 1.0) < 1e-6 and torch.allclose(w, w.flip(0)) and int(w.argmax()) == 5

"""CPU: pieces of bench.py that define the measured workload — the synthetic loss of SURVEY 8(d) (a custom autograd function
with few launches) must equal its plain definition in value and in every gradient."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_synthetic_loss_equals_its_definition():
    b = _bench()
    H, W = 24, 40
    wts = b.loss_weights(H, W, "cpu")
    g = torch.Generator().manual_seed(1)
    out = {k: torch.randn(v.shape, generator=g, requires_grad=True) for k, v in wts.items()}
    gt = torch.rand((3, H, W), generator=g)
    loss = b.synthetic_loss(out, wts, gt)
    (2.5 * loss).backward()                                   # a non-unit upstream gradient reaches every map
    ref_in = {k: v.detach().clone().requires_grad_(True) for k, v in out.items()}
    ref = (ref_in["render"] - gt).abs().mean()
    for k, w in wts.items():
        ref = ref + (ref_in[k] * w).sum()
    (2.5 * ref).backward()
    assert abs(float(loss) - float(ref)) <= 1e-6 * abs(float(ref))
    for k in wts:
        assert torch.allclose(out[k].grad, ref_in[k].grad, rtol=1e-6, atol=1e-9), k


def test_frame_bytes_formula_is_the_survey_figure():
    b = _bench()
    # 979 B per surfel + 316 B per (surfel, tile) instance + 128 B per pixel (+ 156 B per surfel with deformation)
    assert b.frame_bytes(300_000, 800_000, 640_000, True) == 979 * 300_000 + 316 * 800_000 + 128 * 640_000 + 156 * 300_000
    assert b.frame_bytes(10, 0, 0, False) == 9790

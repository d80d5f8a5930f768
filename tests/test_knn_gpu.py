"""GPU: simple_knn._C.distCUDA2 (libd2gs.so: d2gs_knn_mean_dist2) against the oracle, the golden outputs of the reference
extension, and — when oracle/_ref/simple_knn travelled to the box — the live reference extension."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_knn_golden import knn_cases, load_reference_knn  # noqa: E402
from oracle import knn_oracle as ko  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(HERE, "golden", "knn_golden.npz")
RTOL = 1e-6      # squared distances are sums of three fp32 products: the only freedom is the fma contraction


@pytest.mark.parametrize("name", list(knn_cases().keys()))
def test_matches_oracle_and_golden(name, cuda_device):
    from simple_knn._C import distCUDA2
    pts = knn_cases()[name]
    got = distCUDA2(torch.as_tensor(pts, device=cuda_device))
    assert got.dtype == torch.float32 and got.shape == (pts.shape[0],) and got.device.type == "cuda"
    got = got.cpu().numpy()
    np.testing.assert_allclose(got, ko.mean_dist2(pts), rtol=RTOL, atol=1e-12)
    if os.path.exists(GOLDEN):
        np.testing.assert_allclose(got, np.load(GOLDEN)[name], rtol=RTOL, atol=1e-12)


def test_edge_cases(cuda_device):
    from simple_knn._C import distCUDA2
    pts = knn_cases()["tiny7"]
    for n in (0, 1, 2, 3, 4):
        got = distCUDA2(torch.as_tensor(pts[:n], device=cuda_device)).cpu().numpy()
        want = ko.mean_dist2(pts[:n])
        assert got.shape == want.shape and np.array_equal(got, want), (n, got, want)
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(5, 3))                    # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(5, 2, device=cuda_device))
    # a non-contiguous float64 view is accepted like the reference's .contiguous()
    base = torch.as_tensor(knn_cases()["ball5000"], device=cuda_device)
    wide = torch.zeros(5000, 6, device=cuda_device, dtype=torch.float32); wide[:, ::2] = base
    assert torch.equal(distCUDA2(wide[:, ::2]), distCUDA2(base))


def test_large_set_is_exact_and_matches_live_reference(cuda_device):
    """300 k points (the C3 surfel count): exactness through a size-independent property — the result is invariant under a
    permutation of the points and equals the float64 KD-tree answer — plus the live reference extension when present."""
    from scipy.spatial import cKDTree
    from simple_knn._C import distCUDA2
    rng = np.random.default_rng(5)
    P = 300_000
    pts = rng.uniform(-1, 1, size=(P, 3)).astype(np.float32)
    pts[:5000] = pts[5000:10000]                                            # 5000 exact twins
    x = torch.as_tensor(pts, device=cuda_device)
    got = distCUDA2(x)
    perm = torch.randperm(P, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(1))
    assert torch.equal(distCUDA2(x[perm]), got[perm])                      # exact search: no dependence on the order
    d, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=4, workers=-1)
    want = (np.sort(d, axis=1)[:, 1:] ** 2).mean(1)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=3e-6, atol=1e-12)
    ref = load_reference_knn()
    if ref is not None:
        np.testing.assert_allclose(got.cpu().numpy(), ref.distCUDA2(x).cpu().numpy(), rtol=RTOL, atol=1e-12)

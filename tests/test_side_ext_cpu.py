"""CPU: host logic of the two side extensions' drop-ins (simple_knn, diff_gaussian_rasterization) and their C-ABI entry
points — argument validation and error behaviour that needs no GPU (no compute calls)."""
import ctypes as C

import pytest
import torch

from d2gs_b200 import _lib


def test_simple_knn_shim_rejects_what_the_reference_rejects():
    from simple_knn._C import distCUDA2
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        distCUDA2(torch.zeros(5, 3))                       # the reference dereferences a host pointer on the device; we raise
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        distCUDA2([[0.0, 0.0, 0.0]])


def test_knn_and_gs3d_workspace_queries():
    L = _lib.lib()
    b = C.c_size_t(0)
    assert L.d2gs_knn_mean_dist2_workspace(300_000, C.byref(b)) == 0 and b.value > 300_000 * (4 + 16)
    assert L.d2gs_knn_mean_dist2_workspace(0, C.byref(b)) == 0 and b.value > 0
    assert L.d2gs_knn_mean_dist2_workspace(-1, C.byref(b)) < 0 and b"bad" in L.d2gs_last_error()
    assert L.d2gs_knn_mean_dist2(0, None, None, None, 0, None) == 0                       # P = 0: nothing to do
    assert L.d2gs_knn_mean_dist2(10, None, None, None, 0, None) < 0 and b"missing" in L.d2gs_last_error()
    g, i, bn = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert L.d2gs_gs3d_workspace(100_000, 800, 800, 1_000_000, C.byref(g), C.byref(i), C.byref(bn)) == 0
    assert g.value >= 100_000 * (48 + 24 + 1 + 4 + 4) and i.value >= 800 * 800 * 4 + 2500 * 8 and bn.value >= 1_000_000 * 24
    assert L.d2gs_gs3d_workspace(10, 0, 800, 0, C.byref(g), C.byref(i), C.byref(bn)) < 0
    assert L.d2gs_gs3d_forward(None, None) < 0 and L.d2gs_gs3d_backward(None, None) < 0


def test_gs3d_forward_validates_before_launching():
    L = _lib.lib()
    a = _lib.Gs3dFwdArgs()
    a.P, a.width, a.height = 10, 64, 64
    assert L.d2gs_gs3d_forward(C.byref(a), None) < 0 and b"missing outputs" in L.d2gs_last_error()
    a.P = -1
    assert L.d2gs_gs3d_forward(C.byref(a), None) < 0 and b"bad sizes" in L.d2gs_last_error()


def test_runtime_options():
    for name in ("cull", "knn_filter", "tile_sort", "deform_bwd_smem"):
        _lib.set_option(name, 1)
    with pytest.raises(_lib.D2gsError, match="unknown option"):
        _lib.set_option("no_such_switch", 1)


def test_diff_gaussian_rasterization_api_surface():
    import diff_gaussian_rasterization as dgr
    assert dgr.GaussianRasterizationSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                                                         "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    rs = dgr.GaussianRasterizationSettings(8, 8, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
    r = dgr.GaussianRasterizer(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), colors_precomp=x, scales=x)
    with pytest.raises(RuntimeError, match="CUDA tensor"):      # no CPU path (the reference asserts the same in C++)
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))

"""CPU: the C-ABI library loads and exports every symbol include/d2gs.h declares; argument validation works
without a GPU (no compute calls)."""
import ctypes as C
import os
import re

import pytest

import util  # noqa: F401
from d2gs_b200 import _lib

ROOT = util.ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "d2gs.h")).read()
    return sorted(set(re.findall(r"D2GS_API\s+[\w\s\*]+?\b(d2gs_\w+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared_symbols()
    assert len(names) >= 10
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/d2gs.h but not exported by libd2gs.so"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_config_matches_reference_macros():
    c = _lib.config()
    # DSR/cuda_rasterizer/config.h:15-17, auxiliary.h:20-37
    assert (c["num_channels"], c["block_x"], c["block_y"]) == (3, 16, 16)
    assert (c["tight_bbox"], c["render_auxiliary"], c["backface_cull"], c["dual_visible"], c["detach_weight"]) == (0, 1, 1, 1, 0)
    assert c["near_plane"] == 0.2 and c["far_plane"] == 100.0 and c["filter_size"] == 0.7071067811865476
    assert c["sm_arch"] == 100


def test_workspace_sizes_cover_layout():
    g, i, b = _lib.workspace_sizes(1000, 64, 48, 5000)
    assert g >= 1000 * (80 + 1 + 4 + 4)
    assert i >= 64 * 48 * 4 * 5 + 8 * 12
    assert b >= 5000 * 24
    g2, _, b2 = _lib.workspace_sizes(2000, 64, 48, 10000)
    assert g2 > g and b2 > b


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.d2gs_raster_forward(None, None) == -1
    assert b"null" in L.d2gs_last_error()
    a = _lib.RasterFwdArgs()
    a.P, a.width, a.height = -1, 16, 16
    assert L.d2gs_raster_forward(C.byref(a), None) == -1
    assert L.d2gs_raster_backward(None, None) == -1
    assert L.d2gs_deform_forward(None, None) == -1
    assert L.d2gs_deform_backward(None, None) == -1
    assert L.d2gs_mark_visible(-1, None, None, None, None, None) == -1
    with pytest.raises(_lib.D2gsError):
        _lib.check(-1, "probe")


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libd2gs.so")
    with pytest.raises(_lib.D2gsError, match="no CPU fallback"):
        _lib.lib()


def test_dropin_api_surface():
    import inspect
    import diff_surfel_rasterization as d
    assert d.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")
    sig = inspect.signature(d.GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                    "rotations", "cov3D_precomp"]
    sig = inspect.signature(d.rasterize_gaussians)
    assert list(sig.parameters) == ["means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations",
                                    "cov3Ds_precomp", "raster_settings"]
    r = d.GaussianRasterizer(None)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r.forward(None, None, None)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r.forward(None, None, None, shs=1, colors_precomp=1)
    with pytest.raises(Exception, match="scale/rotation pair"):
        r.forward(None, None, None, shs=1)
    with pytest.raises(Exception, match="scale/rotation pair"):
        r.forward(None, None, None, shs=1, scales=1, rotations=1, cov3D_precomp=1)


_STRUCT_PAIRS = {
    "D2gsConfig": "D2gsConfig", "D2gsRasterFwdArgs": "RasterFwdArgs", "D2gsRasterBwdArgs": "RasterBwdArgs",
    "D2gsRasterState": "RasterState", "D2gsEpilogueArgs": "EpilogueArgs", "D2gsLossArgs": "LossArgs",
    "D2gsAdamTensor": "AdamTensor", "D2gsMlpArgs": "MlpArgs", "D2gsDeformFwdArgs": "DeformFwdArgs",
    "D2gsDeformBwdArgs": "DeformBwdArgs", "D2gsGs3dFwdArgs": "Gs3dFwdArgs", "D2gsGs3dBwdArgs": "Gs3dBwdArgs",
    "D2gsGs3dState": "Gs3dState",
}


def _header_structs():
    src = open(os.path.join(ROOT, "include", "d2gs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", src, flags=re.S):
        names = []
        for decl in m.group(2).split(";"):
            for part in decl.strip().split(","):
                nm = re.search(r"(\w+)\s*(\[\d+\])?$", part.strip())
                if nm:
                    names.append(nm.group(1))
        out[m.group(1)] = names
    return out


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The ctypes structures of d2gs_b200/_lib.py against include/d2gs.h as a C compiler lays it out: every struct the
    header declares has a mirror, same field names in the same order, same offsets, same size (the header is plain C)."""
    import shutil
    import subprocess
    structs = _header_structs()
    assert sorted(structs) == sorted(_STRUCT_PAIRS), "a struct of include/d2gs.h has no ctypes mirror listed here"
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "d2gs.h"', "int main(void) {"]
    for cname, fields in structs.items():
        lines.append(f'  printf("{cname} . %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    c_layout = {}
    for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines():
        s, f, v = line.split()
        c_layout.setdefault(s, []).append((f, int(v)))
    for cname, pyname in _STRUCT_PAIRS.items():
        cls = getattr(_lib, pyname)
        py = [(".", C.sizeof(cls))] + [(f[0], getattr(cls, f[0]).offset) for f in cls._fields_]
        assert py == c_layout[cname], (cname, [(a, b) for a, b in zip(py, c_layout[cname]) if a != b][:4])

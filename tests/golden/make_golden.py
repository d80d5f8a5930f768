"""Generates the golden fixtures in tests/golden/ by running the UNMODIFIED reference CUDA extension
(oracle/_ref, built by oracle/build_ref.sh from /root/reference/submodules/diff-surfel-rasterization)
on a B200:   gpurun -- python tests/golden/make_golden.py     -> gpurun_out/golden_<cfg>.npz
The .npz files are then copied into tests/golden/ and committed; nothing at test time reads /root/reference.

Every intermediate the reference keeps in its opaque geomBuffer / binningBuffer / imgBuffer is decoded with the
bump-allocation rule of DSR/cuda_rasterizer/rasterizer_impl.h:21-28 + rasterizer_impl.cu:155-194.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import util  # noqa: E402


def _carve(buf: torch.Tensor, fields):
    """fields: [(name, numpy dtype, count, components)] in allocation order; 128-B alignment of the ADDRESS."""
    base = buf.data_ptr()
    raw = buf.cpu().numpy()
    out, addr = {}, base
    for name, dt, count, comps in fields:
        addr = (addr + 127) & ~127
        nbytes = np.dtype(dt).itemsize * count * comps
        if name is not None:
            a = raw[addr - base: addr - base + nbytes].view(dt)
            out[name] = a.reshape(count, comps).copy() if comps > 1 else a.copy()
        addr += nbytes
    return out


def decode_geom(buf, P, scan_bytes=None):
    # depths, clamped, internal_radii, means2D, transMat, normal_opacity, rgb, tiles_touched, [scan space], point_offsets
    f = [("depths", np.float32, P, 1), ("clamped", np.uint8, P, 3), ("internal_radii", np.int32, P, 1),
         ("means2D", np.float32, P, 2), ("transMat", np.float32, P, 9), ("normal_opacity", np.float32, P, 4),
         ("rgb", np.float32, P, 3), ("tiles_touched", np.uint32, P, 1)]
    out = _carve(buf, f)
    return out


def decode_binning(buf, R):
    f = [("point_list", np.uint32, R, 1), ("point_list_unsorted", np.uint32, R, 1), ("keys_sorted", np.uint64, R, 1),
         ("keys_unsorted", np.uint64, R, 1)]
    return _carve(buf, f)


def decode_img(buf, N):
    f = [("final_T", np.float32, N, 3), ("n_contrib", np.uint32, N, 2), ("ranges", np.uint32, N, 2)]
    o = _carve(buf, f)
    # accum_alpha / n_contrib are plane-major (3 x N), (2 x N)
    o["final_T"] = o["final_T"].reshape(-1)[: 3 * N].reshape(3, N)
    o["n_contrib"] = o["n_contrib"].reshape(-1)[: 2 * N].reshape(2, N)
    return o


def run_reference(ref, act, kw, gc, go, dev):
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
    C = ref._C
    empty = torch.empty(0, device=dev)
    H, W = int(kw["image_height"]), int(kw["image_width"])
    ins = {k: T(v) for k, v in act.items()}
    bg, view, proj, campos = T(kw["bg"]), T(kw["viewmatrix"]), T(kw["projmatrix"]), T(kw["campos"])
    R, color, others, radii, geom, binning, img = C.rasterize_gaussians(
        bg, ins["means3D"], empty, ins["opacities"], ins["scales"], ins["rotations"], 1.0, empty, view, proj,
        float(kw["tanfovx"]), float(kw["tanfovy"]), H, W, ins["shs"], int(kw["sh_degree"]), campos, False, False)
    g = C.rasterize_gaussians_backward(
        bg, ins["means3D"], radii, empty, ins["scales"], ins["rotations"], 1.0, empty, view, proj,
        float(kw["tanfovx"]), float(kw["tanfovy"]), T(gc), T(go), ins["shs"], int(kw["sh_degree"]), campos, geom, R,
        binning, img, False)
    torch.cuda.synchronize()
    names = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales", "dL_drotations")
    P = ins["means3D"].shape[0]
    out = dict(num_rendered=np.int64(R), out_color=color.cpu().numpy(), out_others=others.cpu().numpy(),
               radii=radii.cpu().numpy())
    out.update(decode_geom(geom, P))
    out.update(decode_binning(binning, R))
    im = decode_img(img, H * W)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out["final_T"] = im["final_T"].reshape(3, H, W)
    out["n_contrib"] = im["n_contrib"].reshape(2, H, W)
    out["ranges"] = im["ranges"][:tiles]
    for n, t in zip(names, g):
        out[n] = t.cpu().numpy()
    return out


def main():
    dev = torch.device("cuda:0")
    ref = util.load_reference_ext()
    assert ref is not None, "oracle/_ref is not built (run oracle/build_ref.sh where /root/reference exists)"
    os.makedirs("gpurun_out", exist_ok=True)
    for cfg, cam, deg in (("T0", 3, 3), ("T0", 6, 1)):
        act, kw = util.raster_inputs(cfg, cam_index=cam, sh_degree=deg)
        gc, go = util.upstream_grads(kw["image_height"], kw["image_width"], seed=cam)
        out = run_reference(ref, act, kw, gc, go, dev)
        out.update(meta_cfg=np.array(cfg), meta_cam=np.int64(cam), meta_deg=np.int64(deg), grad_seed=np.int64(cam))
        path = f"gpurun_out/golden_{cfg}_cam{cam}_deg{deg}.npz"
        np.savez_compressed(path, **out)
        print(path, "R =", int(out["num_rendered"]), "visible =", int((out["radii"] > 0).sum()),
              "bytes =", os.path.getsize(path))


if __name__ == "__main__":
    main()

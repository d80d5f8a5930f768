"""Diagnostic (not a pytest): ours vs the live reference extension on one view; counts bit-level mismatches per stage.
    python tests/gpu_parity_diag.py C3 3 [--sync]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dynamic-2dgs_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import numpy as np, torch
import util, make_golden
import test_raster_gpu as trg
from d2gs_b200 import raster, _lib

cfg, cam = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda:0")
ref = util.load_reference_ext()
act, kw = util.raster_inputs(cfg, cam_index=cam, n_cams=100, bg=(0.0, 0.0, 0.0) if cam % 2 else (1.0, 1.0, 1.0))
gc, go = util.upstream_grads(kw["image_height"], kw["image_width"], seed=cam)
g = make_golden.run_reference(ref, act, kw, gc, go, dev)
if "--sync" not in sys.argv:
    raster.set_deferred_count(True, warmup=1, margin=1.5)
    trg.run_ours(act, kw, dev)
o = trg.run_ours(act, kw, dev, gc, go)
st = {k: v.detach().cpu().numpy() for k, v in raster.export_state(o["ctx"]).items()}
rep = {"cfg": cfg, "cam": cam, "R_ours": int(o["ctx"].num_rendered), "R_ref": int(g["num_rendered"])}
vis = g["radii"] > 0
rep["radii_mismatch"] = int((o["radii"].cpu().numpy() != g["radii"]).sum())
for k in ("means2D", "transMat", "normal_opacity", "rgb", "depths"):
    a, b = st[k][vis].view(np.uint32), g[k][vis].view(np.uint32)
    neq = a != b
    rep[k + "_bits_mismatch"] = int(neq.sum())
    rep[k + "_n"] = int(neq.size)
    if neq.any():
        d = np.abs(a.astype(np.int64) - b.astype(np.int64))[neq]
        rep[k + "_max_ulp"] = int(d.max())
        if a.ndim == 2:
            rep[k + "_mismatch_by_col"] = neq.sum(0).tolist()
for k in ("keys_sorted", "point_list", "ranges"):
    rep[k + "_equal"] = bool(np.array_equal(st[k].reshape(-1).view(np.uint32), g[k].reshape(-1).view(np.uint32)))
nc_o, nc_r = st["n_contrib"].view(np.uint32), g["n_contrib"]
d0 = nc_o[0] != nc_r[0]
rep["n_contrib_last_mismatch"] = int(d0.sum())
has = nc_r[0] > 0
rep["n_contrib_median_mismatch"] = int((nc_o[1][has] != nc_r[1][has]).sum())
rep["pixels"] = int(d0.size)
ys, xs = np.nonzero(d0)
rep["examples"] = [{"x": int(x), "y": int(y), "ours": int(nc_o[0][y, x]), "ref": int(nc_r[0][y, x]), "T_ours": float(st["final_T"][0][y, x]),
                    "T_ref": float(g["final_T"][0][y, x])} for y, x in list(zip(ys, xs))[:12]]
fT = st["final_T"].view(np.uint32) != g["final_T"].view(np.uint32)
rep["final_T_bits_mismatch"] = [int(fT[i].sum()) for i in range(3)]
col = o["color"].detach().cpu().numpy()
rep["color_bits_mismatch"] = int((col.view(np.uint32) != g["out_color"].view(np.uint32)).sum())
rep["color_max_abs"] = float(np.abs(col - g["out_color"]).max())
am = o["allmap"].detach().cpu().numpy()
rep["allmap_bits_mismatch_by_plane"] = [int((am[i].view(np.uint32) != g["out_others"][i].view(np.uint32)).sum()) for i in range(8)]
rep["allmap_max_abs_by_plane"] = [float(np.abs(am[i] - g["out_others"][i]).max()) for i in range(8)]
print("DIAG " + json.dumps(rep))

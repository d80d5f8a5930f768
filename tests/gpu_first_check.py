"""First-light GPU script (not a pytest): our CUDA path vs the CPU oracle vs the reference extension."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import util
from oracle import surfel_oracle as so
import diff_surfel_rasterization as ours
from d2gs_b200 import raster

dev = torch.device("cuda:0")
ref = util.load_reference_ext()
print("reference ext:", "loaded" if ref else "ABSENT")
out = {}
for cfg in ("T0", "T1"):
    act, kw = util.raster_inputs(cfg)
    H, W = kw["image_height"], kw["image_width"]
    gc, go = util.upstream_grads(H, W)
    T = lambda a: torch.as_tensor(a, device=dev)
    def run(mod):
        ins = {k: T(v).requires_grad_(True) for k, v in act.items()}
        m2d = torch.zeros_like(ins["means3D"], requires_grad=True)
        rs = util.settings_for(mod, kw, dev)
        r = mod.GaussianRasterizer(rs)
        color, radii, allmap = r(means3D=ins["means3D"], means2D=m2d, opacities=ins["opacities"], shs=ins["shs"],
                                 scales=ins["scales"], rotations=ins["rotations"])
        loss = (color * T(gc)).sum() + (allmap * T(go)).sum()
        loss.backward()
        torch.cuda.synchronize()
        return dict(color=color.detach().cpu().numpy(), radii=radii.cpu().numpy(), allmap=allmap.detach().cpu().numpy(),
                    g_means3D=ins["means3D"].grad.cpu().numpy(), g_means2D=m2d.grad.cpu().numpy(),
                    g_opac=ins["opacities"].grad.cpu().numpy(), g_shs=ins["shs"].grad.cpu().numpy(),
                    g_scales=ins["scales"].grad.cpu().numpy(), g_rot=ins["rotations"].grad.cpu().numpy())
    mine = run(ours)
    st = so.forward(**act, **kw)
    g = so.backward(st, gc, go)
    orc = dict(color=st.out_color, radii=st.radii, allmap=st.out_others, g_means3D=g["dL_dmeans3D"],
               g_means2D=g["dL_dmeans2D"], g_opac=g["dL_dopacity"], g_shs=g["dL_dsh"], g_scales=g["dL_dscales"],
               g_rot=g["dL_drotations"])
    rep = {}
    for k in mine:
        if k == "radii":
            rep["radii_mismatch_vs_oracle"] = int((mine[k] != orc[k]).sum())
        else:
            rep[k + "_rel_vs_oracle"] = util.rel_err(mine[k], orc[k])
    if ref:
        theirs = run(ref)
        for k in mine:
            if k == "radii":
                rep["radii_mismatch_vs_ref"] = int((mine[k] != theirs[k]).sum())
            else:
                rep[k + "_rel_vs_ref"] = util.rel_err(mine[k], theirs[k])
                rep[k + "_maxabs_vs_ref"] = float(np.abs(mine[k] - theirs[k]).max())
                rep[k + "_bitexact_vs_ref"] = bool((mine[k] == theirs[k]).all())
        for k in theirs:
            if k != "radii":
                rep[k + "_ref_rel_vs_oracle"] = util.rel_err(theirs[k], orc[k])
    out[cfg] = rep
    print(cfg, json.dumps(rep, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/first_check.json", "w"), indent=1)

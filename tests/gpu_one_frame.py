"""ncu target (not a pytest): N frames of the C3 rasterizer fwd+bwd only, through the drop-in op."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-2dgs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import util
import diff_surfel_rasterization as ours
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
act, kw = util.raster_inputs(cfg, cam_index=17, n_cams=100, bg=(0, 0, 0))
gc, go = util.upstream_grads(kw["image_height"], kw["image_width"], seed=1)
T = lambda a: torch.as_tensor(a, device=dev)
ins = {k: T(v).requires_grad_(True) for k, v in act.items()}
m2d = torch.zeros_like(ins["means3D"], requires_grad=True)
rs = util.settings_for(ours, kw, dev)
gc, go = T(gc), T(go)
for it in range(n):
    color, radii, allmap = ours.GaussianRasterizer(rs)(means3D=ins["means3D"], means2D=m2d, opacities=ins["opacities"], shs=ins["shs"],
                                                       scales=ins["scales"], rotations=ins["rotations"])
    ((color * gc).sum() + (allmap * go).sum()).backward()
torch.cuda.synchronize()
print("done", float(color.sum()))

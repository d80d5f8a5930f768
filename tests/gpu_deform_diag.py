"""Diagnostic (not a pytest): fused deformation vs the eager oracle at a BASELINE size: neighbour-set equality, delta errors."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dynamic-2dgs_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from d2gs_b200 import deform as dfm, model as mdl, synthetic as syn
from oracle import deform_oracle as do
cfg_name = sys.argv[1] if len(sys.argv) > 1 else "C3"
dev = torch.device("cuda:0")
cfg = syn.CONFIGS[cfg_name]
sc = syn.make_scene(cfg["P"], cfg["seed"], cfg["s_med"], n_nodes=cfg["n_nodes"], hyper_dim=8)
pc = mdl.SurfelModel(sc, dev)
torch.manual_seed(0)
dm = dfm.DeformModel(deform_type="node", is_blender=True, K=4, hyper_dim=8, node_num=cfg["n_nodes"], local_frame=True)
cn = dm.deform
with torch.no_grad():
    cn.nodes.copy_(torch.as_tensor(sc.nodes, device=dev)); cn._node_radius.copy_(torch.as_tensor(sc.node_radius, device=dev))
    for h in (cn.network.gaussian_warp, cn.network.gaussian_rotation, cn.network.local_rotation):
        h.weight.mul_(1e3)
fid = torch.tensor([0.5], device=dev)
with torch.no_grad():
    w, d, i = cn.cal_nn_weight(pc.get_xyz.detach(), feature=pc.feature)
    ours = dm.step(pc.get_xyz.detach(), cn.expand_time(fid), feature=pc.feature, motion_mask=pc.motion_mask)
    net = {k[len("network."):]: v for k, v in cn.named_parameters() if k.startswith("network.")}
    t = fid.reshape(1, 1).expand(cn.nodes.shape[0], 1)
    ref = do.control_node_warp_forward(net, cn.nodes, cn._node_radius, cn._node_weight, pc.get_xyz.detach(), t, pc.feature, pc.motion_mask, 4, 8,
                                       local_frame=True, knn_mode="exact")
    # float64 yardstick for the neighbour sets
    x64 = torch.cat([pc.get_xyz.detach(), pc.feature[:, :8]], -1).double()
    n64 = cn.nodes.detach().double()
    idx64 = torch.cat([torch.sort(((x64[s:s + 8192, None, :] - n64[None]) ** 2).sum(-1), dim=1, stable=True).indices[:, :4] for s in range(0, x64.shape[0], 8192)])
rep = {"cfg": cfg_name}
rep["idx_equal_rows_ours_vs_oracle"] = float((i == ref["nn_idx"]).all(dim=1).float().mean())
rep["idx_equal_rows_ours_vs_f64"] = float((i == idx64).all(dim=1).float().mean())
rep["idx_equal_rows_oracle_vs_f64"] = float((ref["nn_idx"] == idx64).all(dim=1).float().mean())
same = (i == ref["nn_idx"]).all(dim=1)
for k in ("d_xyz", "d_rotation", "d_scaling"):
    a, b = ours[k].double(), ref[k].double()
    rep[k + "_rel"] = float((a - b).norm() / b.norm())
    rep[k + "_rel_same_idx"] = float((a[same] - b[same]).norm() / b[same].norm())
    rep[k + "_absmax"] = float((a - b).abs().max()); rep[k + "_scale"] = float(b.abs().mean())
rep["w_rel_same_idx"] = float((w[same].double() - ref["nn_weight"][same].double()).norm() / ref["nn_weight"][same].double().norm())
print("DDIAG " + json.dumps(rep))

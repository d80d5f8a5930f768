"""A/B harness (not a pytest): times the rasterizer stages of several builds of libd2gs.so on the same views.
    python tests/gpu_ab.py [--cfg C3] [--views 3,17,50,71,97] [--reps 3] default build/variants/libd2gs_x.so ...
Each build runs in its own process (D2GS_LIB selects it; "default" = the in-tree library); prints one line per build with
the per-launch stage times (CUDA events inside the library, `d2gs_profile_*`) and a checksum of outputs / gradients."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(cfg, views, reps):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "dynamic-2dgs_b200"), os.path.join(ROOT, "tests")]
    import numpy as np
    import torch
    import util
    import diff_surfel_rasterization as ours
    from d2gs_b200 import _lib, raster
    dev = torch.device("cuda:0")
    raster.set_deferred_count(True, warmup=1)
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
    act, _ = util.raster_inputs(cfg, cam_index=views[0], n_cams=100, bg=(0, 0, 0))
    ins = {k: T(v).requires_grad_(True) for k, v in act.items()}
    sets = []
    for v in views:
        _, kw = util.raster_inputs(cfg, cam_index=v, n_cams=100, bg=(0, 0, 0))
        sets.append(util.settings_for(ours, kw, dev))
    gc, go = util.upstream_grads(kw["image_height"], kw["image_width"], seed=1)
    gc, go = T(gc), T(go)

    def frame(rs):
        for t in ins.values():
            t.grad = None
        m2d = torch.zeros_like(ins["means3D"], requires_grad=True)
        color, radii, allmap = ours.GaussianRasterizer(rs)(means3D=ins["means3D"], means2D=m2d, opacities=ins["opacities"], shs=ins["shs"],
                                                           scales=ins["scales"], rotations=ins["rotations"])
        ((color * gc).sum() + (allmap * go).sum()).backward()
        return color, allmap
    for rs in sets:
        frame(rs)
    torch.cuda.synchronize()
    _lib.profile_collect(); _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for rs in sets:
            color, allmap = frame(rs)
    e1.record()
    torch.cuda.synchronize()
    st = _lib.profile_collect()
    _lib.profile_enable(False)
    chk = [float(color.double().sum()), float(allmap.double().sum()), float(ins["means3D"].grad.double().abs().sum()),
           float(ins["shs"].grad.double().abs().sum()), float(ins["rotations"].grad.double().abs().sum())]
    ms = {k: round(v[0] / v[1], 4) for k, v in st.items() if v[1]}
    print("AB_JSON " + json.dumps({"stages_ms": ms, "total_ms_per_frame": e0.elapsed_time(e1) / (reps * len(sets)), "checksum": chk}))


if __name__ == "__main__":
    args = sys.argv[1:]
    cfg, views, reps = "C3", [3, 17, 50, 71, 97], 3
    libs = []
    i = 0
    while i < len(args):
        if args[i] == "--cfg":
            cfg = args[i + 1]; i += 2
        elif args[i] == "--views":
            views = [int(x) for x in args[i + 1].split(",")]; i += 2
        elif args[i] == "--reps":
            reps = int(args[i + 1]); i += 2
        elif args[i] == "--one":
            one(cfg, views, reps); sys.exit(0)
        else:
            libs.append(args[i]); i += 1
    for lib in libs:
        env = dict(os.environ)
        lib, _, opts = lib.partition("@")          # "default@tile_order=0": D2GS_OPTIONS for that run
        if opts:
            env["D2GS_OPTIONS"] = opts
        if lib == "default":
            env.pop("D2GS_LIB", None)
        else:
            env["D2GS_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cfg", cfg, "--views", ",".join(map(str, views)), "--reps", str(reps), "--one"],
                           env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("AB_JSON ")]
        if not line:
            print(f"{os.path.basename(lib)}: FAILED\n{r.stdout[-1500:]}\n{r.stderr[-3000:]}", flush=True)
            continue
        d = json.loads(line[0][8:])
        s = d["stages_ms"]
        print(f"{os.path.basename(lib) + ('@' + opts if opts else ''):44s} fwd {s.get('blend_fwd')} bwd {s.get('blend_bwd')} pre_f {s.get('preprocess_fwd')} pre_b {s.get('preprocess_bwd')} "
              f"scan {s.get('scan')} dup {s.get('duplicate')} sort {s.get('sort')} frame {d['total_ms_per_frame']:.3f}  chk {['%.7g' % c for c in d['checksum']]}",
              flush=True)

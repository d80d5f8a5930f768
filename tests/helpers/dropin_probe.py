"""Runs in a subprocess (tests/test_dropin_cpu.py): puts this repo's drop-in packages in FRONT of the reference tree on
sys.path, stubs the third-party modules that are absent from this image (GUI, mesh and metric libraries that the
reference imports at module level but the hot path never touches), then executes the import block of the reference's
train_gui.py / render_mesh.py and reports what every imported name resolved to.  Prints one JSON object."""
import ast
import importlib
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = sys.argv[1]
PKG = os.path.join(ROOT, "dynamic-2dgs_b200")
sys.path[:0] = [PKG, REF]
sys.argv = ["probe"]      # some reference modules build an ArgumentParser at import

ABSENT_OK = {"dearpygui", "open3d", "imageio", "plyfile", "lpips", "trimesh", "kornia", "matplotlib", "skimage", "piq",
             "tinycudann", "mediapy", "roma", "pytorch_msssim", "nvdiffrast", "pymeshlab", "xatlas",
             "tensorboard", "mcubes", "pysdf", "sklearn_extra", "torchmetrics", "ffmpeg", "viser", "nerfview", "OpenGL", "glfw",
             "pygltflib", "moderngl", "pytorch3d"}
stubbed = []


class _Anything(types.ModuleType):
    """A module whose every attribute is a callable stand-in; enough for `from x import y` and class definitions."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None,
                            "__getattr__": lambda self, n: (lambda *a, **k: None)})
        setattr(self, name, v)
        return v


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top in ABSENT_OK:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Anything(spec.name)
        m.__path__ = []
        stubbed.append(spec.name)
        return m

    def exec_module(self, module):
        pass


sys.path.insert(0, PKG)
from d2gs_b200 import pytorch3d_shim  # noqa: E402

shim = pytorch3d_shim.install()           # before the stub finder exists: the real pytorch3d wins if it is installed
sys.meta_path.append(_StubFinder())       # consulted only after the real finders fail

import torch  # noqa: E402

if not torch.cuda.is_available():
    torch.nn.Module.cuda = lambda self, *a, **k: self          # the reference moves modules to the GPU in constructors
    _orig_tensor_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    _orig_to = torch.Tensor.to

    def _to(self, *a, **k):          # the reference's regularisers hard-code device="cuda"
        a = tuple("cpu" if isinstance(v, str) and v.startswith("cuda") else v for v in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return _orig_to(self, *a, **k)
    torch.Tensor.to = _to

import d2gs_b200  # noqa: E402


def import_block(path):
    """Top-level import statements of a reference script, executed one by one."""
    src = open(path).read()
    names, errors = {}, {}
    ns = {}
    for node in ast.parse(src).body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            code = compile(ast.Module([node], []), path, "exec")
            try:
                exec(code, ns)
            except Exception as ex:  # noqa: BLE001
                errors[ast.unparse(node)] = f"{type(ex).__name__}: {ex}"
    for k, v in ns.items():
        if k.startswith("__"):
            continue
        mod = getattr(v, "__module__", None) or getattr(v, "__name__", "")
        f = getattr(sys.modules.get(mod.split(".")[0] if mod else ""), "__file__", None) if mod else None
        names[k] = {"module": mod, "file": getattr(v, "__file__", None) or (sys.modules[mod].__file__ if mod in sys.modules and hasattr(sys.modules[mod], "__file__") else f)}
    return names, errors, ns


out = {"pytorch3d_shim": shim}
names, errors, ns = import_block(os.path.join(REF, "train_gui.py"))
out["train_gui"] = {"names": names, "errors": errors}
names2, errors2, ns2 = import_block(os.path.join(REF, "render_mesh.py"))
out["render_mesh"] = {"names": names2, "errors": errors2}

# ---- bind the deformation classes to the reference's own
ref_dm = d2gs_b200.install_into_reference()
tu = importlib.import_module("utils.time_utils")
Node, Mlp = ref_dm.model_dict["node"], ref_dm.model_dict["mlp"]
from d2gs_b200 import deform as dfm  # noqa: E402

out["bound"] = {
    "node_is_subclass": issubclass(Node, tu.ControlNodeWarp) and issubclass(Node, dfm._FastNodeWarpMixin),
    "mlp_is_subclass": issubclass(Mlp, tu.DeformNetwork) and issubclass(Mlp, dfm._FusedNetworkMixin),
    "forward_is_fast": Node.forward is dfm._FastNodeWarpMixin.forward,
    "cal_nn_weight_is_fast": Node.cal_nn_weight is dfm._FastNodeWarpMixin.cal_nn_weight,
    "inherited": {n: getattr(Node, n) is getattr(tu.ControlNodeWarp, n)
                  for n in ("arap_loss", "densify", "as_gaussians", "init", "state_dict", "load_state_dict", "node_deform",
                            "elastic_loss", "acc_loss", "p2dR", "cal_node_importance", "init_gaussians", "expand_time")},
    "render_is_ours": ns["render"].__module__ if "render" in ns else None,
    "DeformModel_dict_is_patched": ns["DeformModel"].__init__.__globals__["model_dict"]["node"] is Node if "DeformModel" in ns else None,
}

# ---- the bound class on CPU: the calls outside the fast path run the REFERENCE's code (here with the pytorch3d stand-in)
torch.manual_seed(0)
node = Node(is_blender=True, node_num=24, K=3, hyper_dim=2, local_frame=True, with_arap_loss=True)
out["bound"]["network_class"] = [c.__name__ for c in type(node.network).__mro__[:3]]
x = torch.randn(200, 3)
feat = torch.randn(200, 3) * 0.01
with torch.no_grad():
    node.nodes.data[:, :3] = torch.randn(24, 3)
    node._node_radius.data.fill_(-1.0)
    node.network.gaussian_warp.weight.mul_(3e3)      # default head init is ~1e-5: make the deformation (and its ARAP energy) non-trivial
d = node(x, node.expand_time(torch.tensor([0.3])), feat, torch.ones(200, 1), iteration=100)          # CPU -> reference forward, ARAP reg on
out["bound"]["cpu_forward_keys"] = sorted(d.keys())
out["bound"]["cpu_forward_shapes"] = [list(d[k].shape) for k in ("d_xyz", "d_rotation", "d_scaling")]
out["bound"]["reg_loss_is_tensor"] = torch.is_tensor(node.reg_loss) and bool(torch.isfinite(node.reg_loss))
# same inputs through the stand-alone class (plain-torch fallbacks) must give the same weights
alone = dfm.ControlNodeWarp(is_blender=True, node_num=24, K=3, hyper_dim=2, local_frame=True, with_arap_loss=True)
alone.load_state_dict(node.state_dict())
w0, d0, i0 = tu.ControlNodeWarp.cal_nn_weight(node, x, feature=feat)
w1, d1, i1 = alone.cal_nn_weight(x, feature=feat)
out["bound"]["standalone_knn_matches_reference"] = bool(torch.equal(i0, i1) and torch.allclose(w0, w1, atol=1e-6) and torch.allclose(d0, d1, atol=1e-6))
torch.manual_seed(3)
a0 = float(tu.ControlNodeWarp.arap_loss(node, t=torch.tensor(0.4)))
torch.manual_seed(3)
a1 = float(alone.arap_loss(t=torch.tensor(0.4)))
out["bound"]["arap"] = [a0, a1]
out["stubbed"] = sorted(set(s.split(".")[0] for s in stubbed))
print("PROBE_JSON " + json.dumps(out, default=str))

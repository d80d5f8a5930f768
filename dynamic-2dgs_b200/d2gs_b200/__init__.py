"""d2gs_b200 — B200-native (sm_100a) implementation of the Dynamic-2DGS per-frame render hot path.

Python here is plumbing (device memory, streams, autograd glue, torch.distributed); the work is done by the
hand-written CUDA kernels of ``libd2gs.so`` behind the C ABI declared in ``include/d2gs.h``.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "install_into_reference"]


def install_into_reference(patch_renderer: bool = True, stub_pytorch3d: bool = False):
    """Make an importable reference tree (hustvl/Dynamic-2DGS on sys.path) run on the B200 path without editing it:
    see ``d2gs_b200.deform.install_into_reference``.  ``stub_pytorch3d=True`` first registers a plain-torch stand-in for
    the ``pytorch3d`` functions the reference imports at module level (``d2gs_b200.pytorch3d_shim``) when the real
    package is not installed."""
    if stub_pytorch3d:
        from . import pytorch3d_shim
        pytorch3d_shim.install()
    from . import deform
    return deform.install_into_reference(patch_renderer=patch_renderer)

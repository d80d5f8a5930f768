"""d2gs_b200 — B200-native (sm_100a) implementation of the Dynamic-2DGS per-frame render hot path.

Python here is plumbing (device memory, streams, autograd glue, torch.distributed); the work is done by the
hand-written CUDA kernels of ``libd2gs.so`` behind the C ABI declared in ``include/d2gs.h``.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]

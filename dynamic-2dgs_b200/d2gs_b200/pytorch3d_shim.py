"""Plain-torch stand-in for the handful of ``pytorch3d`` functions the reference imports at module level
(utils/time_utils.py:5, utils/deform_utils.py:3-5,12) — pytorch3d is not vendored by the reference, not pinned in its
requirements.txt and not installed in this image.  ``install()`` registers the stand-in in ``sys.modules`` ONLY when the
real package cannot be imported.  None of this is on the hot path: the B200 classes override the one hot call site
(``cal_nn_weight``, utils/time_utils.py:950) with the fused CUDA kernel; what remains are the node-sized regularisers
and editing helpers (M <= a few thousand points), for which an exhaustive search is adequate.

Semantics restated from the published pytorch3d API: ``knn_points(p1, p2, lengths1, lengths2, K, return_nn=False)`` ->
named tuple (dists (N,P1,K) squared L2 ascending, idx (N,P1,K) int64, knn (N,P1,K,D) | None); ``ball_query(p1, p2, K,
radius, return_nn)`` -> first K points of p2 (in index order) within ``radius``, padded with idx -1 / dists 0."""
import sys
import types
from collections import namedtuple

import torch

_KNN = namedtuple("KNN", "dists idx knn")


def knn_points(p1, p2, lengths1=None, lengths2=None, norm: int = 2, K: int = 1, version: int = -1, return_nn: bool = False,
               return_sorted: bool = True):
    if norm != 2:
        raise NotImplementedError("only squared-L2 neighbourhoods are provided")
    dists, idxs, nns = [], [], []
    for b in range(p1.shape[0]):
        a, c = p1[b], p2[b]
        if lengths1 is not None:
            a = a[: int(lengths1[b])]
        if lengths2 is not None:
            c = c[: int(lengths2[b])]
        k = min(K, c.shape[0])
        d = torch.cat([((a[s:s + 4096, None, :] - c[None, :, :]) ** 2).sum(-1) for s in range(0, a.shape[0], 4096)]) if a.shape[0] else \
            a.new_zeros((0, c.shape[0]))
        srt = torch.sort(d, dim=1, stable=True)
        dd, ii = srt.values[:, :k], srt.indices[:, :k]
        if k < K:
            dd = torch.cat([dd, dd.new_zeros(dd.shape[0], K - k)], 1)
            ii = torch.cat([ii, ii.new_zeros(ii.shape[0], K - k)], 1)
        if a.shape[0] < p1.shape[1]:
            pad = p1.shape[1] - a.shape[0]
            dd = torch.cat([dd, dd.new_zeros(pad, K)], 0)
            ii = torch.cat([ii, ii.new_zeros(pad, K)], 0)
        dists.append(dd); idxs.append(ii)
        if return_nn:
            nns.append(p2[b][ii])
    return _KNN(torch.stack(dists), torch.stack(idxs), torch.stack(nns) if return_nn else None)


def ball_query(p1, p2, lengths1=None, lengths2=None, K: int = 500, radius: float = 0.2, return_nn: bool = True):
    dists, idxs, nns = [], [], []
    for b in range(p1.shape[0]):
        d = ((p1[b][:, None, :] - p2[b][None, :, :]) ** 2).sum(-1)
        inside = d < radius * radius
        rank = torch.cumsum(inside.to(torch.int64), dim=1) - 1           # position among the hits, in index order
        ii = torch.full((p1.shape[1], K), -1, dtype=torch.int64, device=p1.device)
        dd = torch.zeros((p1.shape[1], K), dtype=p1.dtype, device=p1.device)
        rows, cols = torch.nonzero(inside & (rank < K), as_tuple=True)
        ii[rows, rank[rows, cols]] = cols
        dd[rows, rank[rows, cols]] = d[rows, cols]
        dists.append(dd); idxs.append(ii)
        if return_nn:
            nns.append(p2[b][ii.clamp_min(0)] * (ii >= 0)[..., None])
    return _KNN(torch.stack(dists), torch.stack(idxs), torch.stack(nns) if return_nn else None)


def _unavailable(name):
    def f(*a, **k):
        raise NotImplementedError(f"pytorch3d.{name} is not provided by the d2gs_b200 stand-in (install pytorch3d for it)")
    return f


def install(force: bool = False) -> bool:
    """Register the stand-in as ``pytorch3d`` if the real package is missing.  Returns True when it was registered."""
    if not force:
        try:
            import pytorch3d.ops  # noqa: F401
            return False
        except Exception:
            pass
    root = types.ModuleType("pytorch3d")
    ops = types.ModuleType("pytorch3d.ops")
    ops.knn_points, ops.ball_query = knn_points, ball_query
    loss = types.ModuleType("pytorch3d.loss")
    mls = types.ModuleType("pytorch3d.loss.mesh_laplacian_smoothing")
    mls.cot_laplacian = _unavailable("loss.mesh_laplacian_smoothing.cot_laplacian")
    io = types.ModuleType("pytorch3d.io")
    io.load_ply = _unavailable("io.load_ply")
    root.ops, root.loss, root.io, loss.mesh_laplacian_smoothing = ops, loss, io, mls
    root.__d2gs_stand_in__ = True
    root.__path__ = []          # a package, so that `import pytorch3d.<other>` fails with ModuleNotFoundError, not AttributeError
    sys.modules.update({"pytorch3d": root, "pytorch3d.ops": ops, "pytorch3d.loss": loss,
                        "pytorch3d.loss.mesh_laplacian_smoothing": mls, "pytorch3d.io": io})
    return True

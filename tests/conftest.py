import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dynamic-2dgs_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _fresh_binning_state():
    """Every test starts with the rasterizer's instance-count history cleared (so its first frames run in the
    synchronous mode and deferred-count capacities never leak between scenes of different tests)."""
    from d2gs_b200 import raster
    raster._TRACK.clear()
    raster._R_HINT.clear()
    raster.set_deferred_count(False)      # the library default (the reference behaviour); tests opt in
    yield

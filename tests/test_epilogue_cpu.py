"""CPU: the restatement of the image-space helpers in oracle/reference_pipeline.py against the reference's own
``depth_to_normal`` / ``depths_to_points`` (tests/golden/epilogue_golden.npz, generated from /root/reference/utils/point_utils.py
by tests/golden/make_epilogue_golden.py)."""
import os
import sys

import numpy as np
import pytest

import util
from oracle import reference_pipeline as rp

sys.path.insert(0, os.path.join(util.ROOT, "tests", "golden"))
from make_epilogue_golden import epilogue_case  # noqa: E402


@pytest.mark.parametrize("name", ["a", "b"])
def test_depth_to_normal_matches_reference(name):
    g = np.load(os.path.join(util.ROOT, "tests", "golden", "epilogue_golden.npz"))
    view, depth = epilogue_case(name)
    n = rp.depth_to_normal(view, depth)
    n = n[0] if isinstance(n, tuple) else n
    assert np.array_equal(n.numpy(), g[f"{name}_normal"])                       # same torch ops: bit-identical on CPU
    assert np.array_equal(rp.depths_to_points(view, depth).numpy(), g[f"{name}_points"])
    assert not n.numpy()[0].any() and not n.numpy()[:, -1].any()               # the one-pixel border stays zero (point_utils.py:37)

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as d:
        return {k: d[k] for k in d.files}


# Tolerances stated by BASELINE.json north_star:
#   |delta r| <= 1e-6, relative error in P <= 1e-4 for P >= 1e-300.
R_ATOL = 1e-6
P_RTOL = 1e-4
P_FLOOR = 1e-300


def assert_p_close(p, p_ref, rtol=P_RTOL, floor=P_FLOOR, what="P"):
    """relative error <= rtol wherever the reference P >= floor; below the floor the
    reference itself flushes to zero around 1e-308, so only smallness is required."""
    p = np.asarray(p)
    p_ref = np.asarray(p_ref)
    assert p.shape == p_ref.shape, (what, p.shape, p_ref.shape)
    assert np.isfinite(p).all(), what
    big = p_ref >= floor
    rel = np.abs(p[big] - p_ref[big]) / p_ref[big]
    assert rel.size == 0 or rel.max() <= rtol, "%s: max rel err %.3e" % (what, rel.max())
    assert (p[~big] <= floor * (1 + rtol) * 10).all(), what + ": tail not small"
    assert (p >= 0).all() and (p <= 1).all(), what


def pearson_from(dot, var_x, var_y):
    return dot / np.sqrt(np.outer(var_x, var_y))


@pytest.fixture(scope="session")
def golden():
    return load_golden

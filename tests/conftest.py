import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_pkg()


@pytest.fixture(scope="session")
def P(pkg):
    return pkg.problems


@pytest.fixture(scope="session")
def cpu_oracle():
    from oracle import cpu
    cpu.build()
    return cpu


@pytest.fixture(scope="session")
def solver(pkg):
    """A live admmb handle on cuda:0.  Fails loudly (no fallback) when the library or GPU is missing."""
    s = pkg.Solver()
    yield s
    s.close()


def assert_bit_identical(got, ref, what=""):
    """GPU vs canonical-order C oracle: identical iteration counts, statuses and iterates."""
    xg, zg, ug, hg = got
    xr, zr, ur, hr = ref
    assert np.array_equal(hg["iters"], hr["iters"]), f"{what}: iteration counts differ: " \
        f"{int((hg['iters'] != hr['iters']).sum())} of {len(hr['iters'])} problems"
    assert np.array_equal(hg["status"], hr["status"]), f"{what}: statuses differ"
    for a, b, name in ((xg, xr, "x"), (zg, zr, "z"), (ug, ur, "u")):
        if a is None:
            continue
        assert np.array_equal(a, b), f"{what}: {name} differs, max |d| = {np.nanmax(np.abs(a - b)):.3e}"
    for k in ("r_norm", "s_norm", "eps_pri", "eps_dual", "rho"):
        assert np.array_equal(hg[k], hr[k]), f"{what}: final {k} differs"
    if "hist" in hr and "hist" in hg:
        for k, v in hr["hist"].items():
            assert np.array_equal(hg["hist"][k], v, equal_nan=True), f"{what}: history {k} differs"

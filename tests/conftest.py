import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def reference_path():
    """where the unmodified reference can be imported from: the read-only tree of the build container, or the offline
    install under baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).  None if neither exists."""
    for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(p, "getdist")):
            return p
    return None


@pytest.fixture(scope="session")
def getdist_ref():
    p = reference_path()
    if p is None:
        pytest.skip("the reference (getdist) is not importable on this machine")
    if p not in sys.path:
        sys.path.insert(0, p)
    import getdist

    return getdist

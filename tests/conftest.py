import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def arch():
    from vae_npvc_b200 import vcc2016_vae_arch
    return vcc2016_vae_arch()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the C-ABI library exists (cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()

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


# Alternative (valid) architectures: asymmetric SAME pads, strides 2/3/4, odd transposed-conv crops,
# a channel count that needs padding -- the plan / kernels are generic, not VCC2016-specific.
ALT_ARCHS = {
    "small_even": {"hwc": [90, 1, 1], "z_dim": 8, "y_dim": 3,
                   "encoder": {"kernel": [[5, 1], [4, 1]], "stride": [[3, 1], [2, 1]], "output": [4, 8]},
                   "generator": {"hwc": [10, 1, 5], "kernel": [[6, 1], [5, 1], [91, 1]], "stride": [[3, 1], [3, 1], [1, 1]], "output": [8, 4, 1]}},
    "odd_pads": {"hwc": [100, 1, 1], "z_dim": 12, "y_dim": 4,
                 "encoder": {"kernel": [[7, 1], [3, 1], [8, 1]], "stride": [[3, 1], [2, 1], [4, 1]], "output": [8, 4, 12]},
                 "generator": {"hwc": [25, 1, 6], "kernel": [[4, 1], [2, 1], [199, 1]], "stride": [[2, 1], [2, 1], [1, 1]], "output": [4, 8, 1]}},
    "wide": {"hwc": [162, 1, 1], "z_dim": 32, "y_dim": 5,
             "encoder": {"kernel": [[7, 1], [7, 1], [5, 1]], "stride": [[3, 1], [3, 1], [2, 1]], "output": [16, 32, 64]},
             "generator": {"hwc": [18, 1, 33], "kernel": [[9, 1], [7, 1], [323, 1]], "stride": [[3, 1], [3, 1], [1, 1]], "output": [32, 16, 1]}},
}


def _add_package_archs():
    # cfg5 (BASELINE.json configs[4]): the VAWGAN discriminator conv stack as the encoder of the path (vae_npvc_b200/arch.py)
    from vae_npvc_b200.arch import vawgan_d_stack_arch
    a = vawgan_d_stack_arch()
    ALT_ARCHS["vawgan_d_stack"] = {k: a[k] for k in ("hwc", "z_dim", "y_dim", "encoder", "generator")}


_add_package_archs()


@pytest.fixture(params=sorted(ALT_ARCHS))
def alt_arch(request):
    import copy
    return copy.deepcopy(ALT_ARCHS[request.param])

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def AdaProx():
    """The product package bound to cuda:0.  Fails loudly without the CUDA library."""
    import adaprox_b200
    adaprox_b200.default_device()
    return adaprox_b200


@pytest.fixture(scope="session")
def lasso_small():
    import adaprox_b200
    P = adaprox_b200.synth.planted_lasso(400, 1000, 5, 0)
    P["Lf"] = adaprox_b200.synth.spectral_norm_sq(P["A"], iters=1000, tol=1e-15)
    return P

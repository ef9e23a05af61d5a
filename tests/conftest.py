import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def haar_unitary(n, seed):
    """scipy.stats.unitary_group.rvs(n, random_state=seed) -- the reference benchmarks' input
    (permanent_benchmark.py:56-57)."""
    from scipy.stats import unitary_group
    import numpy as np
    if n == 1:
        rng = np.random.default_rng(seed)
        return np.exp(2j * np.pi * rng.random((1, 1)))
    return unitary_group.rvs(n, random_state=seed)


def relerr(x, ref):
    return abs(complex(x) - complex(ref)) / abs(complex(ref))


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference calculators (oracle/_ref/libpqbref.so)."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref/libpqbref.so not built (run `make -C oracle ref` in the dev container)")
    return refbind

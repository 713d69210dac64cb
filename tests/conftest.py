import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    path = os.path.join(ROOT, "tests", "golden", "haspi_ref.npz")
    z = np.load(path)
    names = [str(n) for n in z["names"]]
    return {n: {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(n + "/")} for n in names}


def golden_dither(which, rows=16384):
    """The shared dither stream of tests/golden/make_golden.py."""
    import numpy as np

    return np.random.RandomState((1234, 5678)[which]).standard_normal((rows, 32))

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.loader import Port
    return Port()


@pytest.fixture(scope="session")
def reference():
    from oracle.loader import Reference, build_ref
    build_ref()
    if not Reference.available():
        pytest.skip("oracle/_ref/libdpref.so not present (built only where /root/reference exists)")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")

    def load(name):
        return np.load(os.path.join(d, name + ".npz"))
    return load

import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def T():
    import _lib
    _lib.oracle()  # builds oracle/libkslam_oracle.so if needed
    return _lib


@pytest.fixture(scope="session")
def pkg(T):
    return T.load_pkg()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gdir = os.path.join(HERE, "golden")

    def load(name):
        return np.load(os.path.join(gdir, name))
    return load

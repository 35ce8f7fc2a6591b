import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def dg():
    return importlib.import_module("turbo-range-coder_b200.datagen")


@pytest.fixture(scope="session")
def port():
    from oracle import cpu
    return cpu.port()


@pytest.fixture(scope="session")
def ref():
    """Compiled reference (oracle/_ref); tests that need it skip when it was never built."""
    from oracle import cpu
    r = cpu.ref()
    if r is None:
        pytest.skip("oracle/_ref/libtrcref.so not built (needs /root/reference at build time)")
    return r


@pytest.fixture(scope="session")
def trc():
    return importlib.import_module("turbo-range-coder_b200")


@pytest.fixture(scope="session")
def sources(dg):
    n = 300_000
    return {"zipf": dg.zipf(n), "bwt": dg.bwt_shaped(n), "o1": dg.markov1(n), "uniform": dg.uniform(n)}

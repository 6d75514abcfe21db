import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    import cpu_checkers
    return cpu_checkers.load("orc")


@pytest.fixture(scope="session")
def ref():
    import cpu_checkers
    lib = cpu_checkers.load("ref")
    if lib is None:
        pytest.skip("oracle/_ref/libpmg_ref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(here, "golden.json")) as fh:
        g = json.load(fh)
    g["fields"] = dict(np.load(os.path.join(here, "golden_fields.npz")))
    return g

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def entry():
    import __graft_entry__ as g
    g.build()
    return g


@pytest.fixture(scope="session")
def orc(entry):
    return entry.load_oracle()


@pytest.fixture(scope="session")
def sbn(entry):
    return entry.load_package()


@pytest.fixture(scope="session")
def golden():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


@pytest.fixture(scope="session")
def ctx(sbn):
    c = sbn.Context(0)
    yield c
    c.close()

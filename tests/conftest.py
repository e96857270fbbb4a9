import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "depth-map-fusion-utils_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def dmf():
    import dmf_b200
    return dmf_b200


@pytest.fixture(scope="session")
def ctx(dmf):
    """One GPU context for the whole session; fails loudly if the CUDA library or the GPU is missing."""
    c = dmf.Context.default(0)
    yield c

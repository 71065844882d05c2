import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_cls():
    from oracle.oracle_py import Oracle
    return Oracle


@pytest.fixture(scope="session")
def renderer_cls():
    # no fallback: on a box without the built library or without a GPU this raises
    from radlite_b200.api import Renderer
    return Renderer

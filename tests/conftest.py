import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without an NVIDIA device skips the GPU tests instead of failing them.  On a GPU
    box nothing is skipped: a missing or unloadable library fails loudly (radlite_b200 has no CPU fallback)."""
    if os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl"):
        return
    skip = pytest.mark.skip(reason="no NVIDIA device on this box (GPU tests: pytest -m gpu on a B200)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_cls():
    from oracle.oracle_py import Oracle
    return Oracle


@pytest.fixture(scope="session")
def renderer_cls():
    # no fallback: on a box without the built library or without a GPU this raises
    from radlite_b200.api import Renderer
    return Renderer


def same_bits(a, b):
    """Spectra of one line rendered in different batches / subsets: bit-identical when one integrate
    kernel serves both renders (api.DEFAULT_KERNEL pinned), 1e-10 when the library picks the kernel by lines per
    batch (tile_kernel below 8 lines, ztile_kernel from 8: they round differently at the 1e-13 level)."""
    import numpy as np
    from radlite_b200 import api
    if api.DEFAULT_KERNEL in ("z", "tile", "chan"):
        return np.array_equal(a, b)
    return np.allclose(a, b, rtol=1e-10, atol=0.0)


@pytest.fixture(params=["auto", "z", "tile", "chan"])
def integrate_kernel(request):
    """Runs a GPU test once per integrate kernel (library default by regime, ztile_kernel, tile_kernel, chan_kernel)."""
    from radlite_b200 import api
    old = api.DEFAULT_KERNEL
    api.DEFAULT_KERNEL = request.param
    yield request.param
    api.DEFAULT_KERNEL = old

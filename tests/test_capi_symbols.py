"""The C-ABI library loads on a CPU-only box, exports every symbol include/*.h declares, and refuses
to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("rl_create", "rl_destroy", "rl_set_grid", "rl_set_medium", "rl_set_lines", "rl_set_dust",
                 "rl_set_camera", "rl_set_bc", "rl_set_options", "rl_render", "rl_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from radlite_b200.api import LIB_PATH
    assert os.path.exists(LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    from radlite_b200.api import Renderer
    with pytest.raises(RuntimeError):
        Renderer(0)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing in the package may reference it
    for path in glob.glob(os.path.join(ROOT, "radlite_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
            txt = open(path, errors="replace").read()
            for bad in ("libradlite_oracle", "oracle_py", "import oracle", "from oracle", "radlite_oracle.c"):
                assert bad not in txt, (path, bad)

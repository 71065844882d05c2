"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see radlite_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED (no reference binary, no reference goldens)."""
from __future__ import annotations

import ctypes
import os
import subprocess

from radlite_b200._binding import Binding

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libradlite_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "radlite_oracle.c")
    hdr = os.path.join(_HERE, "radlite_oracle.h")
    if (force or not os.path.exists(_LIB)
            or os.path.getmtime(_LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libradlite_oracle.so"])
    return _LIB


def load():
    return ctypes.CDLL(build())


class Oracle(Binding):
    def __init__(self):
        super().__init__(load(), "orc_")

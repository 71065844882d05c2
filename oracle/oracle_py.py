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

    def set_ring_sample(self, lo, hi, stride=1):
        """Benchmark sampling only (see radlite_oracle.h)."""
        self.lib.orc_set_ring_sample.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3
        self.lib.orc_set_ring_sample.restype = None
        self.lib.orc_set_ring_sample(self.ctx, int(lo), int(hi), int(stride))

    def trajectory(self, iray):
        """Nodes of ray ``iray`` (1-based) as make_trajectory_c builds them."""
        import numpy as np
        L = self.lib
        L.orc_max_nodes.argtypes = [ctypes.c_void_p]
        L.orc_max_nodes.restype = ctypes.c_int
        n = L.orc_max_nodes(self.ctx)
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
        L.orc_trajectory.argtypes = [ctypes.c_void_p, ctypes.c_int] + [dp] * 5 + [ip] * 3
        L.orc_trajectory.restype = ctypes.c_int
        d = [np.zeros(n) for _ in range(5)]
        i = [np.zeros(n, dtype=np.int32) for _ in range(3)]
        k = L.orc_trajectory(self.ctx, int(iray), *[a.ctypes.data_as(dp) for a in d],
                             *[a.ctypes.data_as(ip) for a in i])
        if k < 0:
            self._check(-k)
        names = ("s", "radius", "theta", "mu", "phi", "icross", "iradius", "itheta")
        return {nm: a[:k] for nm, a in zip(names, d + i)}

    def node_values(self, iray, iline):
        """get_line_dust_values at every node of ray ``iray`` for line ``iline`` (1-based)."""
        import numpy as np
        L = self.lib
        L.orc_max_nodes.argtypes = [ctypes.c_void_p]
        L.orc_max_nodes.restype = ctypes.c_int
        n = L.orc_max_nodes(self.ctx)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_node_values.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [dp] * 6
        L.orc_node_values.restype = ctypes.c_int
        d = [np.zeros(n) for _ in range(6)]
        k = L.orc_node_values(self.ctx, int(iray), int(iline), *[a.ctypes.data_as(dp) for a in d])
        if k < 0:
            self._check(-k)
        names = ("srcd", "alpd", "dvmu", "lw", "nup", "ndown")
        return {nm: a[:k] for nm, a in zip(names, d)}

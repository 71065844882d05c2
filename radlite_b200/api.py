"""Host-side access to the CUDA library through its C ABI (include/radlite_b200.h).

``Renderer`` is a thin ctypes object over ``libradlite_b200.so``.  There is no CPU fallback: if the
library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``) or no B200 is
visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._binding import Binding, RadliteError, _d, _i

_HERE = os.path.dirname(os.path.abspath(__file__))
# RADLITE_B200_LIB selects a tuning variant built by `make variant` (development only)
LIB_PATH = os.environ.get("RADLITE_B200_LIB") or os.path.join(_HERE, "libradlite_b200.so")
_lib = None
# integrate kernel every new Renderer is pinned to: "auto" (library default: by lines per batch), "z", "tile", "chan"
# (the GPU parity tests run every model under all three)
DEFAULT_KERNEL = "auto"
_KERNEL_MODES = {"auto": 0, "z": 1, "tile": 2, "chan": 3}


def load_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                "(make -C radlite_b200/csrc).  radlite_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.rl_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        _lib.rl_create.restype = C.c_int
        _lib.rl_render_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                          C.c_double, C.POINTER(C.c_float)]
        _lib.rl_render_device.restype = C.c_int
        _lib.rl_fetch_flux.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.rl_fetch_flux.restype = C.c_int
        dp = C.POINTER(C.c_double)
        _lib.rl_render_rings.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_int, C.c_int, dp, dp]
        _lib.rl_render_rings.restype = C.c_int
        _lib.rl_flux_from_rings.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, dp, dp]
        _lib.rl_flux_from_rings.restype = C.c_int
        _lib.rl_render_rings_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                                C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float)]
        _lib.rl_render_rings_device.restype = C.c_int
        _lib.rl_flux_from_rings_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, dp]
        _lib.rl_flux_from_rings_device.restype = C.c_int
        _lib.rl_plan_costs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, dp]
        _lib.rl_plan_costs.restype = C.c_int
        _lib.rl_launch_count.argtypes = [C.c_void_p]
        _lib.rl_launch_count.restype = C.c_longlong
        _lib.rl_total_nodes.argtypes = [C.c_void_p]
        _lib.rl_total_nodes.restype = C.c_longlong
        _lib.rl_max_nodes.argtypes = [C.c_void_p]
        _lib.rl_max_nodes.restype = C.c_int
        _lib.rl_get_ray_nodes.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 5 + [
            C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.rl_get_ray_nodes.restype = C.c_int
    return _lib


class Renderer(Binding):
    """One context on one GPU (one process per GPU: pass LOCAL_RANK as ``device``)."""

    def __init__(self, device: int = 0):
        lib = load_library()
        try:
            super().__init__(lib, "rl_", create_args=(int(device),))
        except RadliteError as e:
            raise RuntimeError(
                f"rl_create(device={device}) failed with {e.code}: no usable sm_100 GPU "
                "(radlite_b200 has no CPU fallback)") from e
        self.kernel = "auto"
        if DEFAULT_KERNEL != "auto":
            self.set_kernel(DEFAULT_KERNEL)

    def set_kernel(self, mode: str):
        """Pin the integrate kernel: "auto", "z" (ztile_kernel + zcont_kernel), "tile" (tile_kernel) or "chan"
        (chan_kernel + zcont_kernel)."""
        self.lib.rl_set_kernel.argtypes = [C.c_void_p, C.c_int]
        self.lib.rl_set_kernel.restype = C.c_int
        self._check(self.lib.rl_set_kernel(self.ctx, _KERNEL_MODES[mode]))
        self.kernel = mode

    def render_device(self, iline0, nl, nfr, vmax_kms, dist_cm):
        """Kernel-only pass: inputs resident, nothing copied back.  Returns CUDA-event times [ms]
        of (geometry, per-line preparation, ray integration, flux reduction, whole call)."""
        ms = (C.c_float * 5)()
        self._check(self.lib.rl_render_device(self.ctx, int(iline0), int(nl), int(nfr),
                                              float(vmax_kms), float(dist_cm), ms))
        return [float(x) for x in ms]

    def render_rings(self, iline0, nl, nfr, vmax_kms, dist_cm, ring_lo, ring_hi, image=None):
        """Trace only the camera rings ring_lo..ring_hi (0 = central beam): the per-rank piece of a
        ring-block sharded render.  Returns ringsum[nl, nrr+1, nfr] (rows outside the block are 0);
        ``image`` (optional, full [nl, nrr+1, nphi, nfr] array) receives the rows of the block."""
        nrr, _, _ = self.camera_dims()
        rs = np.zeros((nl, nrr + 1, nfr))
        self._check(self.lib.rl_render_rings(self.ctx, int(iline0), int(nl), int(nfr), float(vmax_kms),
                                             float(dist_cm), int(ring_lo), int(ring_hi), _d(rs), _d(image)))
        return rs

    def flux_from_rings(self, ringsum, dist_cm):
        """Index-ordered ring sum / distance^2 (telescope.F:1388-1433) of gathered ring sums."""
        ringsum = np.ascontiguousarray(ringsum, dtype=np.float64)
        nl, _, nfr = ringsum.shape
        flux = np.zeros((nl, nfr))
        self._check(self.lib.rl_flux_from_rings(self.ctx, nl, nfr, float(dist_cm), _d(ringsum), _d(flux)))
        return flux

    def render_rings_device(self, iline0, nl, nfr, vmax_kms, dist_cm, ring_lo, ring_hi, d_ringsum: int):
        """Ring-block render that stays on the device: the ring sums [nl, nrr+1, nfr] are written to the DEVICE
        address ``d_ringsum`` (e.g. ``tensor.data_ptr()`` of a float64 CUDA tensor), complete on return.
        Returns the CUDA-event times [ms] like ``render_device``."""
        ms = (C.c_float * 5)()
        self._check(self.lib.rl_render_rings_device(self.ctx, int(iline0), int(nl), int(nfr), float(vmax_kms),
                                                    float(dist_cm), int(ring_lo), int(ring_hi),
                                                    C.c_void_p(int(d_ringsum)), ms))
        return [float(x) for x in ms]

    def flux_from_rings_device(self, d_ringsum: int, nl, nfr, dist_cm):
        """Index-ordered ring sum / distance^2 of the (reduced) ring sums at DEVICE address ``d_ringsum``."""
        flux = np.zeros((nl, nfr))
        self._check(self.lib.rl_flux_from_rings_device(self.ctx, int(nl), int(nfr), float(dist_cm),
                                                       C.c_void_p(int(d_ringsum)), _d(flux)))
        return flux

    def plan_costs(self, iline0, nl, nfr, vmax_kms):
        """Work estimate per camera ring (rl_plan_costs): weights for ``shard.split_rings``."""
        nrr, _, _ = self.camera_dims()
        cost = np.zeros(nrr + 1)
        self._check(self.lib.rl_plan_costs(self.ctx, int(iline0), int(nl), int(nfr), float(vmax_kms), _d(cost)))
        return cost

    def set_wall_tau(self, tau: float):
        """Opaque-wall start (include/radlite_b200.h): 0 integrates every segment like the reference."""
        self.lib.rl_set_wall_tau.argtypes = [C.c_void_p, C.c_double]
        self.lib.rl_set_wall_tau.restype = C.c_int
        self._check(self.lib.rl_set_wall_tau(self.ctx, float(tau)))

    def executed_elements(self) -> float:
        """Element integrations actually performed since the last reset_counters()."""
        self.lib.rl_get_executed.argtypes = [C.c_void_p]
        self.lib.rl_get_executed.restype = C.c_double
        return float(self.lib.rl_get_executed(self.ctx))

    def invalidate_geometry(self):
        self.lib.rl_invalidate_geometry.argtypes = [C.c_void_p]
        self.lib.rl_invalidate_geometry.restype = None
        self.lib.rl_invalidate_geometry(self.ctx)

    def fp64_peak_tflops(self) -> float:
        self.lib.rl_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self.lib.rl_fp64_peak.restype = C.c_int
        v = C.c_double()
        self._check(self.lib.rl_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def fetch_flux(self, nl, nfr):
        out = np.zeros((nl, nfr))
        self._check(self.lib.rl_fetch_flux(self.ctx, int(nl), int(nfr), _d(out)))
        return out

    def launch_count(self) -> int:
        return int(self.lib.rl_launch_count(self.ctx))

    def total_nodes(self) -> int:
        return int(self.lib.rl_total_nodes(self.ctx))

    def ray_nodes(self, iray):
        """Device-built node list of ray ``iray`` (1-based; 1 = centre)."""
        n = max(1, int(self.lib.rl_max_nodes(self.ctx)))
        # geometry may not be built yet: the first call sizes it
        probe = self.lib.rl_get_ray_nodes(self.ctx, int(iray), None, None, None, None, None, None,
                                          None)
        if probe < 0:
            self._check(-probe)
        n = max(n, probe, int(self.lib.rl_max_nodes(self.ctx)))
        ds, dvmu, lw, wr, wt = (np.zeros(n) for _ in range(5))
        cells = np.zeros((n, 4), dtype=np.int32)
        flags = np.zeros(n, dtype=np.int32)
        k = self.lib.rl_get_ray_nodes(self.ctx, int(iray), _d(ds), _d(dvmu), _d(lw), _d(wr),
                                      _d(wt), _i(cells), _i(flags))
        if k < 0:
            self._check(-k)
        return dict(ds=ds[:k], dvmu=dvmu[:k], lw=lw[:k], wr=wr[:k], wt=wt[:k], cells=cells[:k],
                    flags=flags[:k])

"""Host-side mirror of the driver-facing ends of the path (include/radlite_b200.h, "driver-side ends"): LTE
level populations and spectrum synthesis on the GPU, with the call shapes of the pyradlite routines they
replace (RadliteModel._prep_mol_forcore, radlite.py:1024-1126; RadliteSpectrum._process_spectrum,
radlite.py:3001-3184).  No arithmetic of the path lives here: both functions marshal numpy arrays into
libradlite_b200.so and fail loudly without it."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._binding import _d, _f64, _i, _i32
from .api import Renderer, load_library

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _declare(lib):
    lib.rl_set_lines_lte.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp]
    lib.rl_set_lines_lte.restype = C.c_int
    lib.rl_synthesis_size.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, _ip, _ip]
    lib.rl_synthesis_size.restype = C.c_int
    lib.rl_synthesize_spectrum.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double,
                                           C.c_double, _dp, _dp, _dp, _dp]
    lib.rl_synthesize_spectrum.restype = C.c_int


def set_lines_lte(g: Renderer, lev_up, lev_down, linefreq, aud, gdeg, energy_cm, tgas, psum_temp, psum):
    """rl_set_lines with the populations computed on the device: n_lev = g exp(-E h c / k T) / Q(T)
    (radlite.py:1111-1119).  ``tgas`` [nr, nth] K; ``psum_temp``/``psum`` the tabulated partition sum."""
    _declare(g.lib)
    lev_up, lev_down = _i32(lev_up), _i32(lev_down)
    linefreq, aud, gdeg, energy_cm = _f64(linefreq), _f64(aud), _f64(gdeg), _f64(energy_cm)
    tgas, psum_temp, psum = _f64(tgas), _f64(psum_temp), _f64(psum)
    assert tgas.shape == (g.nr, g.nth) and len(energy_cm) == len(gdeg)
    g.nlines = len(lev_up)
    g._check(g.lib.rl_set_lines_lte(g.ctx, len(lev_up), len(gdeg), _i(lev_up), _i(lev_down), _d(linefreq), _d(aud),
                                    _d(gdeg), _d(energy_cm), _d(tgas), len(psum), _d(psum_temp), _d(psum)))


def synthesize_spectrum(g: Renderer, vel, flux, freq, dist_pc, obsres, vsampling):
    """RadliteSpectrum._process_spectrum on the device.  ``vel``, ``flux`` [nl, nfr] as the linespectrum file
    lists them (velocity ascending).  Returns dict(wavelength, spectrum, emission, continuum, frequency)."""
    _declare(g.lib)
    vel, flux, freq = _f64(vel), _f64(flux), _f64(freq)
    nl, nfr = vel.shape
    nout, nfull = C.c_int(), C.c_int()
    rc = g.lib.rl_synthesis_size(nl, nfr, _d(vel), _d(freq), float(obsres), float(vsampling), C.byref(nout), C.byref(nfull))
    if rc:
        raise ValueError("rl_synthesis_size: bad arguments")
    out = {k: np.zeros(nout.value) for k in ("wavelength", "spectrum", "emission", "continuum")}
    g._check(g.lib.rl_synthesize_spectrum(g.ctx, nl, nfr, _d(vel), _d(flux), _d(freq), float(dist_pc), float(obsres),
                                          float(vsampling), _d(out["wavelength"]), _d(out["spectrum"]),
                                          _d(out["emission"]), _d(out["continuum"])))
    out["frequency"] = 2.99792458E10 / (out["wavelength"] * 1.0E-4)  # radlite.py:3167
    return out


def spectrum_from_render(m, flux):
    """The (vel, flux) tables of linespectrum_<mol>.dat from a render's flux [nl, nfr] (channel order of the
    library = ascending frequency): velocity = -c (nu - nu0) / nu0 + v_lsr in km/s, rows written from the last
    channel to the first, i.e. in ascending velocity (telescope.F:1775, 1786-1800)."""
    nfr = flux.shape[1]
    passb = 3.33567e-6 * np.abs(m.linefreq)[:, None] * m.passband
    dnu = -passb + np.arange(nfr)[None, :] * (2.0 * passb / (nfr - 1.0))
    nu = np.abs(m.linefreq)[:, None] + dnu
    vel = -2.99792458e5 * (nu - m.linefreq[:, None]) / m.linefreq[:, None] + m.vlsr
    return np.ascontiguousarray(vel[:, ::-1]), np.ascontiguousarray(flux[:, ::-1])

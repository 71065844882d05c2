"""ctypes binder for a RADLite line-raytracer C ABI (include/radlite_b200.h).

The CUDA library (prefix ``rl_``) and the CPU oracle used by the tests (prefix ``orc_``, see
oracle/radlite_oracle.h) deliberately export the same call shapes, so one binder drives both.
This module only knows how to marshal numpy arrays into plain pointers; it contains no
arithmetic of the path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _i(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class RadliteError(RuntimeError):
    """Non-zero status from the library; ``code`` is the reference's ``stop`` code."""

    def __init__(self, code, msg):
        super().__init__(f"status {code}: {msg}")
        self.code = code


class Binding:
    """Thin object wrapper over ``<prefix>create/set_*/render/destroy``."""

    def __init__(self, lib: C.CDLL, prefix: str, create_args=()):
        self.lib = lib
        self.p = prefix
        self._declare()
        self.ctx = C.c_void_p()
        rc = self._fn("create")(C.byref(self.ctx), *create_args)
        if rc != 0:
            raise RadliteError(rc, "create failed")
        self.nr = self.nth = self.nlines = 0

    # -- plumbing ---------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self.lib, self.p + name)

    def _declare(self):
        L, p = self.lib, self.p
        vp = C.c_void_p
        sig = {
            "set_grid": [vp, C.c_int, C.c_int, _dp, _dp],
            "set_grid_ghosted": [vp, C.c_int, C.c_int, _dp, _dp],
            "set_medium": [vp, _dp, _dp, _dp, _dp, C.c_double],
            "set_lines": [vp, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp],
            "set_dust": [vp, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp],
            "set_line_dust": [vp, _dp, _dp],
            "set_camera": [vp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int],
            "set_bc": [vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp],
            "set_options": [vp, C.c_int, C.c_int, C.c_double, C.c_double],
            "get_camera_dims": [vp, _ip, _ip, _ip],
            "get_rings": [vp, _dp, _dp],
            "render": [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, _ip, _dp,
                       _ip, _dp],
        }
        for name, args in sig.items():
            f = getattr(L, p + name)
            f.argtypes = args
            f.restype = C.c_int
        getattr(L, p + "last_error").argtypes = [vp]
        getattr(L, p + "last_error").restype = C.c_char_p
        getattr(L, p + "destroy").argtypes = [vp]
        getattr(L, p + "destroy").restype = None
        getattr(L, p + "get_counters").argtypes = [vp, _dp, _dp, _dp]
        getattr(L, p + "get_counters").restype = None
        getattr(L, p + "reset_counters").argtypes = [vp]
        getattr(L, p + "reset_counters").restype = None

    def _check(self, rc):
        if rc != 0:
            raise RadliteError(rc, self._fn("last_error")(self.ctx).decode(errors="replace"))

    def close(self):
        if self.ctx:
            self._fn("destroy")(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- setters ----------------------------------------------------------------------------
    def set_grid(self, r, theta):
        r, theta = _f64(r), _f64(theta)
        self.nr, self.nth = len(r), len(theta)
        self._check(self._fn("set_grid")(self.ctx, self.nr, self.nth, _d(r), _d(theta)))

    def set_medium(self, rho, abund, vel, linewidth, umass_av):
        rho, abund, vel, linewidth = _f64(rho), _f64(abund), _f64(vel), _f64(linewidth)
        assert rho.shape == (self.nr, self.nth) and vel.shape == (self.nr, self.nth, 3)
        self._check(self._fn("set_medium")(self.ctx, _d(rho), _d(abund), _d(vel), _d(linewidth),
                                           float(umass_av)))

    def set_lines(self, lev_up, lev_down, linefreq, aud, gdeg, popul):
        lev_up, lev_down = _i32(lev_up), _i32(lev_down)
        linefreq, aud, gdeg, popul = _f64(linefreq), _f64(aud), _f64(gdeg), _f64(popul)
        self.nlines = len(lev_up)
        nlev = len(gdeg)
        assert popul.shape == (self.nr, self.nth, nlev)
        self._check(self._fn("set_lines")(self.ctx, self.nlines, nlev, _i(lev_up), _i(lev_down),
                                          _d(linefreq), _d(aud), _d(gdeg), _d(popul)))

    def set_dust(self, nsize, cont_freq_nu, kappa_abs, kappa_scat, dust_rho, dust_temp,
                 scati_src=None):
        nsize = _i32(nsize)
        cf, ka, ks = _f64(cont_freq_nu), _f64(kappa_abs), _f64(kappa_scat)
        dr, dt, sc = _f64(dust_rho), _f64(dust_temp), _f64(scati_src)
        nspec, ncf, ms = len(nsize), len(cf), int(nsize.max())
        assert ka.shape == (nspec, ms, ncf) and ks.shape == (nspec, ms, ncf)
        assert dr.shape == (self.nr, self.nth, nspec) and dt.shape == (self.nr, self.nth, nspec, ms)
        self._check(self._fn("set_dust")(self.ctx, nspec, _i(nsize), ncf, _d(cf), _d(ka), _d(ks),
                                         _d(dr), _d(dt), _d(sc)))

    def set_line_dust(self, src, alp):
        src, alp = _f64(src), _f64(alp)
        assert src.shape == (self.nlines, self.nr, self.nth)
        self._check(self._fn("set_line_dust")(self.ctx, _d(src), _d(alp)))

    def set_camera(self, anginf, nphi, nrext, dbdr, rstar, imethod=1, nrref=10):
        self._check(self._fn("set_camera")(self.ctx, float(anginf), int(nphi), int(nrext),
                                           int(dbdr), float(rstar), int(imethod), int(nrref)))

    def set_bc(self, in_itype, out_itype, cont_freq_nu, starspec_cont, isrf_cont=None):
        cf, ss, isrf = _f64(cont_freq_nu), _f64(starspec_cont), _f64(isrf_cont)
        self._check(self._fn("set_bc")(self.ctx, int(in_itype), int(out_itype), len(cf), _d(cf),
                                       _d(ss), _d(isrf)))

    def set_options(self, subgrid=1, nonredundant=1, levthres=1e-3, aksmax=-1.0):
        self._check(self._fn("set_options")(self.ctx, int(subgrid), int(nonredundant),
                                            float(levthres), float(aksmax)))

    # -- queries ----------------------------------------------------------------------------
    def camera_dims(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self._fn("get_camera_dims")(self.ctx, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value  # nrr, nphi, nray

    def rings(self):
        nrr, _, _ = self.camera_dims()
        r = np.zeros(nrr + 1)
        ri = np.zeros(nrr + 2)
        self._check(self._fn("get_rings")(self.ctx, _d(r), _d(ri)))
        return r, ri

    def counters(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._fn("get_counters")(self.ctx, C.byref(a), C.byref(b), C.byref(c))
        return {"R": a.value, "E": b.value, "S": c.value}

    def reset_counters(self):
        self._fn("reset_counters")(self.ctx)

    # -- the hot path -----------------------------------------------------------------------
    def render(self, iline0, nl, nfr, vmax_kms, dist_cm, want_image=False, want_mask=False):
        """Lines iline0..iline0+nl-1 (1-based).  Returns dict(flux[nl,nfr], velo[nl,nfr],
        tau_center[nl], maserflag[nl], image[nl,nrr+1,nphi,nfr]?, cmask?)."""
        nrr, nphi, _ = self.camera_dims()
        flux = np.zeros((nl, nfr))
        velo = np.zeros((nl, nfr))
        tau = np.zeros(nl)
        maser = np.zeros(nl, dtype=np.int32)
        img = np.zeros((nl, nrr + 1, nphi, nfr)) if want_image else None
        msk = np.zeros((nl, nrr + 1, nphi, nfr), dtype=np.int32) if want_mask else None
        self._check(self._fn("render")(self.ctx, int(iline0), int(nl), int(nfr), float(vmax_kms),
                                       float(dist_cm), _d(flux), _d(img), _i(msk), _d(tau),
                                       _i(maser), _d(velo)))
        out = {"flux": flux, "velo": velo, "tau_center": tau, "maserflag": maser}
        if want_image:
            out["image"] = img
        if want_mask:
            out["cmask"] = msk
        return out

    # -- rectangular imager / position-velocity cube (linespectrum.inp command 2) ---------------
    def set_camera_rect(self, anginf, nx, ny, sizepix_x, sizepix_y, phioffset=0.0, xoffset=0.0, yoffset=0.0,
                        rstar=0.0, addstar=0):
        f = self._fn("set_camera_rect")
        f.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int] + [C.c_double] * 6 + [C.c_int]
        f.restype = C.c_int
        self._check(f(self.ctx, float(anginf), int(nx), int(ny), float(sizepix_x), float(sizepix_y),
                      float(phioffset), float(xoffset), float(yoffset), float(rstar), int(addstar)))
        self._rect = (int(nx), int(ny))

    def render_rect(self, iline0, nl, nfr, vmax_kms, want_tau=True):
        """Returns dict(image[nl,nx,ny,nfr], tau[nl,nx,ny,nfr]?, maserflag[nl]) = imrec_int / imrec_tau."""
        f = self._fn("render_rect")
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _ip]
        f.restype = C.c_int
        nx, ny = self._rect
        img = np.zeros((nl, nx, ny, nfr))
        tau = np.zeros((nl, nx, ny, nfr)) if want_tau else None
        maser = np.zeros(nl, dtype=np.int32)
        self._check(f(self.ctx, int(iline0), int(nl), int(nfr), float(vmax_kms), _d(img), _d(tau), _i(maser)))
        out = {"image": img, "maserflag": maser}
        if want_tau:
            out["tau"] = tau
        return out

    def load_model(self, m, lines=None):
        """Push a ``synth.Model`` (or any object with the same attributes) through the setters."""
        self.set_grid(m.r, m.theta)
        self.set_medium(m.rho, m.abund, m.vel, m.linewidth, m.umass_av)
        sl = slice(None) if lines is None else lines
        self.set_lines(m.lev_up[sl], m.lev_down[sl], m.linefreq[sl], m.aud[sl], m.gdeg, m.popul)
        self.set_dust(m.nsize, m.cont_freq_nu, m.kappa_abs, m.kappa_scat, m.dust_rho, m.dust_temp,
                      m.scati_src)
        self.set_camera(m.anginf, m.nphi, m.nrext, m.dbdr, m.rstar, m.imethod, m.nrref)
        self.set_bc(m.in_itype, m.out_itype, m.cont_freq_nu, m.starspec_cont, m.isrf_cont)
        self.set_options(m.subgrid, m.nonredundant, m.levthres, m.aksmax)

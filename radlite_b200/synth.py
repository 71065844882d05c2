"""Deterministic synthetic RADLite models for the BASELINE.json configs.

Everything here is *input preparation*: analytic disk structure on the (r,theta) grid, a
CO-like ro-vibrational molecule, LTE / two-temperature level populations, a one-species dust
opacity law and a blackbody star.  It mirrors what the reference's drivers put into the input
files (pyradlite/pyradlite/radlite.py:1481-1737 for the gas fields, :1024-1128 for LTE
populations; PRO/problem_*.pro for the RADMC structure), not the ray tracer itself.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

AU = 1.49598e13
RSUN = 6.96e10
MSUN = 1.98892e33
GG = 6.67259e-8
KB = 1.380658e-16
MP = 1.6726e-24
HH = 6.6260755e-27
CC = 2.99792458e10
PARSEC = 3.08572e18  # main.F:213


@dataclass
class Model:
    name: str
    # grid
    r: np.ndarray
    theta: np.ndarray
    # medium [nr][nth]
    rho: np.ndarray
    abund: np.ndarray
    vel: np.ndarray  # [nr][nth][3] cm/s (v_r, v_theta, v_phi)
    linewidth: np.ndarray  # km/s
    tgas: np.ndarray
    umass_av: float
    # molecule
    molname: str
    molweight: float
    ener_cm: np.ndarray
    gdeg: np.ndarray
    lev_v: np.ndarray
    lev_j: np.ndarray
    lev_up: np.ndarray
    lev_down: np.ndarray
    aud: np.ndarray
    linefreq: np.ndarray
    popul: np.ndarray  # [nr][nth][nlev]
    # dust
    nsize: np.ndarray
    cont_freq_nu: np.ndarray
    kappa_abs: np.ndarray
    kappa_scat: np.ndarray
    dust_rho: np.ndarray
    dust_temp: np.ndarray
    scati_src: np.ndarray | None
    # star / boundary
    rstar: float
    mstar: float
    tstar: float
    starspec_cont: np.ndarray
    isrf_cont: np.ndarray | None = None
    in_itype: int = 2
    out_itype: int = 0
    # camera (line_params.ini defaults) and passband
    incl_deg: float = 15.0
    nphi: int = 150
    nrext: int = -60
    dbdr: int = 1
    imethod: int = 1
    nrref: int = 10
    vmax_kms: float = 70.0
    dv_kms: float = 1.5
    vlsr: float = 0.0
    # options
    subgrid: int = 1
    nonredundant: int = 1
    levthres: float = 1e-3
    aksmax: float = -1.0
    extra: dict = field(default_factory=dict)

    @property
    def anginf(self) -> float:
        # telescope.F:188  linespec_incang = incl * 0.0174532925199d0
        return self.incl_deg * 0.0174532925199

    @property
    def nfr(self) -> int:
        # main.F:209  integer truncation of 2*vmax/dv+1
        return int(2.0 * self.vmax_kms / self.dv_kms + 1)

    @property
    def passband(self) -> float:
        # main.F:210  vmax := 0.5*(nfr-1)*dv
        return 0.5 * (self.nfr - 1) * self.dv_kms

    @property
    def nlines(self) -> int:
        return len(self.lev_up)

    @property
    def nrr(self) -> int:
        return abs(self.nrext) + self.nrref + (len(self.r) - 1) * self.dbdr

    @property
    def nray(self) -> int:
        return 1 + self.nrr * self.nphi


# ------------------------------------------------------------------------------------------
# molecule: CO-like Dunham ladder, dv=1 P/R lines
# ------------------------------------------------------------------------------------------
_CO12 = dict(we=2169.81358, wexe=13.28831, be=1.93128087, ae=0.01750441, de=6.12147e-6)


def _isotopologue(const, rho_iso):
    return dict(we=const["we"] * rho_iso, wexe=const["wexe"] * rho_iso**2,
                be=const["be"] * rho_iso**2, ae=const["ae"] * rho_iso**3,
                de=const["de"] * rho_iso**4)


def rovib_molecule(vmax, jmax, const=_CO12, a10=35.0, band=None, nlines=None, vup_list=None):
    """Levels (v<=vmax, J<=jmax) sorted by energy, 1-based level numbers, and all
    (v,J')->(v-1,J'+-1) lines, optionally restricted to a wavelength band [um] and truncated to
    ``nlines``.  Energies are rounded to the moldata ``F12.4`` precision and A to ``E12.3``
    (line.F:1902,1933) so file and in-memory paths see identical numbers."""
    vv, jj = np.meshgrid(np.arange(vmax + 1), np.arange(jmax + 1), indexing="ij")
    vv, jj = vv.ravel(), jj.ravel()
    bv = const["be"] - const["ae"] * (vv + 0.5)
    g0 = const["we"] * 0.5 - const["wexe"] * 0.25
    evib = const["we"] * (vv + 0.5) - const["wexe"] * (vv + 0.5) ** 2 - g0
    ener = evib + bv * jj * (jj + 1) - const["de"] * (jj * (jj + 1)) ** 2
    ener = np.round(ener, 4)
    order = np.argsort(ener, kind="stable")
    vv, jj, ener = vv[order], jj[order], ener[order]
    gdeg = 2.0 * jj + 1.0
    index = {(int(v), int(j)): k + 1 for k, (v, j) in enumerate(zip(vv, jj))}
    nu_band = const["we"] - 2 * const["wexe"]
    ups, dns, auds, wl = [], [], [], []
    vups = vup_list if vup_list is not None else range(1, vmax + 1)
    for vu in vups:
        for ju in range(0, jmax + 1):
            for jl, hl in ((ju + 1, (ju + 1.0) / (2 * ju + 1.0)), (ju - 1, ju / (2 * ju + 1.0))):
                if jl < 0 or jl > jmax:
                    continue
                iu, il = index[(vu, ju)], index[(vu - 1, jl)]
                dnu = ener[iu - 1] - ener[il - 1]
                if dnu <= 0:
                    continue
                lam = 1e4 / dnu
                if band is not None and not (band[0] <= lam <= band[1]):
                    continue
                a = a10 * vu * hl * (dnu / nu_band) ** 3
                a = float(f"{a:.3e}")
                ups.append(iu), dns.append(il), auds.append(a), wl.append(lam)
    o = np.argsort(wl, kind="stable")
    ups, dns, auds = np.array(ups)[o], np.array(dns)[o], np.array(auds)[o]
    if nlines is not None:
        if len(ups) < nlines:
            raise ValueError(f"only {len(ups)} lines available, {nlines} requested")
        # keep the nlines closest to the band centre, then restore wavelength order
        mid = len(ups) // 2
        lo = max(0, mid - nlines // 2)
        ups, dns, auds = ups[lo:lo + nlines], dns[lo:lo + nlines], auds[lo:lo + nlines]
    # line.F:1903,1981: E[erg] = 1.986468498d-16*E[cm^-1]; nu = 1.509160d26*(Eup-Edown)
    eerg = 1.986468498e-16 * ener
    linefreq = 1.509160e26 * (eerg[ups - 1] - eerg[dns - 1])
    return dict(ener_cm=ener, gdeg=gdeg, lev_v=vv.astype(np.int32), lev_j=jj.astype(np.int32),
                lev_up=ups.astype(np.int32), lev_down=dns.astype(np.int32),
                aud=auds.astype(np.float64), linefreq=linefreq)


def level_populations(mol, tgas, tvib_cap=None):
    """Fractional populations [nr][nth][nlev].  LTE: g exp(-E/kT)/Q(T) (radlite.py:1115-1121);
    with ``tvib_cap`` a two-temperature NLTE stand-in: rotation at T, vibration at min(T,cap)."""
    ek = mol["ener_cm"] * HH * CC / KB
    t = tgas[..., None]
    if tvib_cap is None:
        w = mol["gdeg"] * np.exp(-ek / t)
    else:
        j0 = {}
        for k, (v, j) in enumerate(zip(mol["lev_v"], mol["lev_j"])):
            if j == 0:
                j0[int(v)] = ek[k]
        evib = np.array([j0[int(v)] for v in mol["lev_v"]])
        erot = ek - evib
        tv = np.minimum(t, tvib_cap)
        w = mol["gdeg"] * np.exp(-erot / t) * np.exp(-evib / tv)
    pop = w / w.sum(axis=-1, keepdims=True)
    pop[pop < 1e-99] = 0.0
    return np.ascontiguousarray(pop)


# ------------------------------------------------------------------------------------------
# disk structure
# ------------------------------------------------------------------------------------------
def planck(nu, t):
    return 2.0 * HH * nu**3 / CC**2 / np.expm1(HH * nu / (KB * t))


def make_model(name, nr, nth, mol, *, incl_deg=15.0, vmax_kms=70.0, dv_kms=1.5, nphi=150,
               nrext=-60, dbdr=1, rin_au=0.1, rout_au=100.0, width_scale=1.0, tvib_cap=None,
               freezeout_k=None, molname="co", molweight=28.0, theta_min=0.157, out_itype=0,
               sigma0_dust=1.0, kappa0=1.0e3, abund0=1.0e-4) -> Model:
    r = np.ascontiguousarray(AU * np.logspace(np.log10(rin_au), np.log10(rout_au), nr))
    # theta: nth points, refined towards the mid-plane, last point strictly below pi/2
    x = (np.arange(nth)) / (nth - 0.5)
    theta = np.ascontiguousarray(0.5 * np.pi - (0.5 * np.pi - theta_min) * (1.0 - x) ** 1.5)
    assert theta[-1] < 0.5 * np.pi
    rr, tt = np.meshgrid(r, theta, indexing="ij")
    rcyl, z = rr * np.sin(tt), rr * np.cos(tt)
    rau = rcyl / AU
    h = 0.1 * rau**0.25 * rcyl
    sigma_d = sigma0_dust / rau
    rho_d = np.maximum(sigma_d / (np.sqrt(2 * np.pi) * h) * np.exp(-0.5 * (z / h) ** 2), 1e-26)
    tmid = np.maximum(1200.0 * (rau / 0.1) ** -0.5, 10.0)
    zq = 4.0 * h
    tsurf = 2.0 * tmid
    tgas = np.where(z < zq, tmid + (tsurf - tmid) * np.sin(0.5 * np.pi * z / zq) ** 2, tsurf)
    rho_g = 100.0 * rho_d
    abund = np.full_like(rho_g, abund0)
    if freezeout_k is not None:
        abund = np.where(tgas < freezeout_k, abund0 * 1e-3, abund0)
    vel = np.zeros((nr, nth, 3))
    mstar = 1.0 * MSUN
    vel[..., 2] = np.sqrt(GG * mstar / rr)
    cs = np.sqrt(1.4 * KB * tgas / (2.3 * MP))
    lw = width_scale * np.sqrt((0.9 * cs) ** 2 + 2.0 * KB * tgas / (molweight * MP)) * 1e-5
    # dust: one species, one size, kappa_abs ~ nu, no scattering
    cf = np.logspace(np.log10(3e10), np.log10(3e15), 130)
    kabs = (kappa0 * cf / 6.4e13).reshape(1, 1, -1)
    kscat = np.zeros_like(kabs)
    rstar, tstar = 2.0 * RSUN, 4275.0
    popul = level_populations(mol, tgas, tvib_cap)
    m = Model(name=name, r=r, theta=theta, rho=np.ascontiguousarray(rho_g), abund=abund,
              vel=vel, linewidth=np.ascontiguousarray(lw), tgas=tgas, umass_av=2.3,
              molname=molname, molweight=molweight, ener_cm=mol["ener_cm"], gdeg=mol["gdeg"],
              lev_v=mol["lev_v"], lev_j=mol["lev_j"], lev_up=mol["lev_up"],
              lev_down=mol["lev_down"], aud=mol["aud"], linefreq=mol["linefreq"], popul=popul,
              nsize=np.array([1], dtype=np.int32), cont_freq_nu=cf, kappa_abs=kabs,
              kappa_scat=kscat, dust_rho=np.array(rho_d[..., None], order="C"),
              dust_temp=np.array(tgas[..., None, None], order="C"), scati_src=None,
              rstar=rstar, mstar=mstar, tstar=tstar, starspec_cont=planck(cf, tstar),
              incl_deg=incl_deg, nphi=nphi, nrext=nrext, dbdr=dbdr, vmax_kms=vmax_kms,
              dv_kms=dv_kms, out_itype=out_itype)
    if out_itype == 3:
        m.isrf_cont = 1e-3 * planck(cf, 2.0e4) * 1e-14
    return m


def config(n: int, *, nr=None, nth=None, nlines=None, nphi=None, nrext=None) -> Model:
    """BASELINE.json configs 1..5 (optionally shrunk for tests via the keyword overrides)."""
    if n == 1:
        mol = rovib_molecule(1, 30)
        k = [i for i in range(len(mol["lev_up"]))
             if mol["lev_v"][mol["lev_up"][i] - 1] == 1 and mol["lev_j"][mol["lev_up"][i] - 1] == 9
             and mol["lev_j"][mol["lev_down"][i] - 1] == 10]  # v=1-0 P(10)
        for key in ("lev_up", "lev_down", "aud", "linefreq"):
            mol[key] = mol[key][k]
        m = make_model("cfg1_P10_100x40", nr or 100, nth or 40, mol)
    elif n == 2:
        mol = rovib_molecule(1, 51, nlines=nlines or 100)
        m = make_model("cfg2_CO100_200x80", nr or 200, nth or 80, mol)
    elif n == 3:
        rho_iso = np.sqrt((12.0 * 15.9949 / 27.9949) / (13.00335 * 15.9949 / 28.99825))
        mol = rovib_molecule(1, 30, const=_isotopologue(_CO12, rho_iso))
        k = [i for i in range(len(mol["lev_up"]))
             if mol["lev_j"][mol["lev_up"][i] - 1] == 9 and mol["lev_j"][mol["lev_down"][i] - 1] == 10]
        for key in ("lev_up", "lev_down", "aud", "linefreq"):
            mol[key] = mol[key][k]
        m = make_model("cfg3_13CO_cube_400x160", nr or 400, nth or 160, mol, incl_deg=45.0,
                       vmax_kms=50.0, dv_kms=0.5025, molname="13co", molweight=29.0,
                       abund0=1.0e-4 / 70.0)
    elif n == 4:
        # "~500 lines": every v<=9, J<=60 fundamental/hot-band line inside 4.6-5.0 um (442 of them)
        mol = rovib_molecule(9, 60, band=(4.6, 5.0), nlines=nlines)
        m = make_model("cfg4_12CO_nlte_500", nr or 200, nth or 80, mol, tvib_cap=1000.0,
                       freezeout_k=20.0)
    elif n == 5:
        mol = rovib_molecule(12, 100, nlines=nlines or 2000)
        m = make_model("cfg5_2000lines_1000x400", nr or 1000, nth or 400, mol, width_scale=0.2)
    else:
        raise ValueError(n)
    if nphi is not None:
        m.nphi = nphi
    if nrext is not None:
        m.nrext = nrext
    return m

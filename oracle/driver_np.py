"""numpy restatement of the driver-side pieces of the path: what pyradlite computes
before it spawns RADlite (gas fields, LTE level populations) and after it (spectrum synthesis).

TEST INFRASTRUCTURE ONLY, like everything under oracle/: the checker for radlite_b200.driver (the CUDA
implementation behind rl_lte_populations / rl_synthesize_spectrum).  pyradlite itself cannot be imported in
this image (it needs astropy and matplotlib at import time, radlite.py:1-24), so its arithmetic is restated
here function by function, each citing the lines it follows; PARITY UNPINNED applies as for the C oracle.
The constants are the astropy-free fallbacks of radlite.py:27-41.
"""
from __future__ import annotations

import numpy as np

# radlite.py:27-41
au0 = 1.49597870E13
c0 = 2.99792458E10
h0 = 6.6262000E-27
kB0 = 1.3807E-16
mp0 = 1.6726231E-24
G0 = 6.67259E-8
cinmu0 = c0 * 1.0E4   # mu/s
cinkm0 = c0 / 1.0E5   # km/s


# ---- before the run ------------------------------------------------------------------------------
def gas_density(dustdens, gastodust):
    """radlite.py:1548-1549 (_calc_gasdensity)."""
    return dustdens.copy() * gastodust


def gas_temperature(dusttemp):
    """radlite.py:1577 (_calc_gastemperature): gas temperature = dust temperature."""
    return dusttemp.copy()


def abundance(dusttemp, min_abun, max_abun, temp_fr=False):
    """radlite.py:1513-1522 (_calc_abundance): constant, or min_abun below the freeze-out temperature."""
    a = np.ones(dusttemp.shape) * max_abun
    if temp_fr is not False:
        a[dusttemp < temp_fr] = min_abun
    return a


def turbulence(dusttemp, alpha, gamma, mu, molweight):
    """radlite.py:1618-1622 (_calc_turbulence) [cm/s]: alpha c_s and thermal broadening in quadrature."""
    cs = np.sqrt(gamma * kB0 * dusttemp / 1.0 / (mu * mp0))
    turb = alpha * cs
    therm = np.sqrt(2.0 * kB0 * dusttemp / (molweight * mp0))
    return np.sqrt((turb ** 2) + (therm ** 2))


def velocity_phi(radius, nth, mstar):
    """radlite.py:1653-1664 (_calc_velocity): Keplerian v_phi on the (theta, r) grid; v_r = v_theta = 0."""
    rexp = np.resize(radius, (nth, len(radius)))
    return np.sqrt(G0 * mstar / 1.0 / rexp)


def interp_linear_extrapolate(x, y, xnew):
    """scipy.interpolate.interp1d(kind='linear', bounds_error=False, fill_value='extrapolate') as used at
    radlite.py:1152: searchsorted, clip to [1, n-1], slope * (xnew - x_lo) + y_lo."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    xnew = np.asarray(xnew, dtype=np.float64)
    idx = np.clip(np.searchsorted(x, xnew), 1, len(x) - 1)
    lo, hi = idx - 1, idx
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
    return slope * (xnew - x[lo]) + y[lo]


def lte_populations(e_cm, g, tgas, psum_temp, psum):
    """radlite.py:1111-1119 (_prep_mol_forcore) with the partition sum of :1147-1153 (_prep_mol_forall):
    n_lev = g exp(-E h c / k T) / Q(T), Q interpolated linearly (extrapolating) in the tabulated
    partition sum; values below 1e-99 are flushed to 0.  Returns [nlev, *tgas.shape]."""
    ek = np.asarray(e_cm) * h0 * c0 / 1.0 / kB0
    q = interp_linear_extrapolate(psum_temp, psum, tgas)
    npop = np.array([(g[a] * np.exp(-1.0 * ek[a] / tgas)) for a in range(len(ek))]) / 1.0 / q
    npop[npop < 1E-99] = 0.0
    return npop


def unique_levels(e_low, e_up, g_low, g_up, v_low, v_up):
    """radlite.py:1071-1100 (_prep_mol_forcore): the levels of a line list, sorted by energy, duplicates
    (same energy to 1e-4 relative, same v, same g) removed.  Returns (E, g, v, index of each line's lower
    level, index of each line's upper level) with 0-based indices into the unique list."""
    e_all = np.concatenate((e_low, e_up))
    v_all = np.concatenate((v_low, v_up))
    g_all = np.concatenate((g_low, g_up))
    order = np.argsort(e_all, kind="stable")
    e_s, v_s, g_s = e_all[order], v_all[order], g_all[order]
    keep = np.ones(len(e_s), dtype=bool)
    for a in range(1, len(e_s)):
        if ((np.abs(e_s[a - 1] - e_s[a]) / 1.0 / (e_s[a] + 0.1)) < 0.0001) and (v_s[a - 1] == v_s[a]) and (g_s[a - 1] == g_s[a]):
            keep[a] = False
    e_u, v_u, g_u = e_s[keep], v_s[keep], g_s[keep]

    def find(e, v, gg):
        # radlite.py:1954-1975 (_write_core_moldatadat) looks levels up by the same closeness test
        for k in range(len(e_u)):
            if (np.abs(e_u[k] - e) / 1.0 / (e + 0.1) < 0.0001) and v_u[k] == v and g_u[k] == gg:
                return k
        raise ValueError("level not found")

    low = np.array([find(e_low[i], v_low[i], g_low[i]) for i in range(len(e_low))])
    up = np.array([find(e_up[i], v_up[i], g_up[i]) for i in range(len(e_up))])
    return e_u, g_u, v_u, low, up


# ---- after the run -------------------------------------------------------------------------------
def _interp1d_linear(x, y, xnew):
    """scipy interp1d(kind='linear') inside its range (radlite.py:3081, 3105, 3125, 3158-3164)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    xnew = np.atleast_1d(np.asarray(xnew, dtype=np.float64))
    idx = np.clip(np.searchsorted(x, xnew), 1, len(x) - 1)
    lo, hi = idx - 1, idx
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
    return slope * (xnew - x[lo]) + y[lo]


def synthesis_grids(vel, freq, obsres, vsampling):
    """radlite.py:3031-3067: box width, the full-resolution wavelength grid (uniform in velocity) and the
    output grid.  vel [nl, nfr] km/s as written by RADLite (ascending), freq [nl] Hz."""
    maxvspan = np.max([np.abs(v[0]) for v in vel])
    boxwidth = np.max([(3 * maxvspan), (3 * obsres)])
    vres = vel[0][1] - vel[0][0]  # every line carries the same number of points (argmax -> 0)
    fr = np.asarray(freq)
    fr = fr[fr != 0]
    muminraw = np.min(cinmu0 / 1.0 / fr)
    mumin = muminraw - (boxwidth * 1.0E9 * muminraw / cinmu0)
    mumaxraw = np.max(cinmu0 / 1.0 / fr)
    mumax = mumaxraw + (boxwidth * 1.0E9 * mumaxraw / cinmu0)
    growth = (1 + (vres / 1.0 / cinkm0))
    maxlen = int(np.floor(np.log(mumax / 1.0 / mumin) / np.log(growth)) + 1)
    fullmu = np.array([(mumin * (growth ** a)) for a in range(0, maxlen)])
    userfrac = (1 + (vsampling / 1.0 / cinkm0))
    userlen = int(np.floor(np.log(mumax / 1.0 / mumin) / np.log(userfrac)) + 1)
    outmu = np.array([(mumin * (userfrac ** a)) for a in range(0, userlen)])
    return dict(boxwidth=boxwidth, vres=vres, mumin=mumin, mumax=mumax, fullmu=fullmu, outmu=outmu)


def gauss_kernel(obsres, vres):
    """radlite.py:3141-3144."""
    numgauss = np.ceil(3.0 * obsres / vres)
    top = -1 * 2 * ((np.arange(0, numgauss) - ((numgauss - 1) / 2.0)) ** 2)
    bot = ((obsres / 1.0 / vres) ** 2) * np.log(2.0)
    return np.exp(top / bot)


def convolve_reflect(a, k):
    """scipy.ndimage.convolve(a, k, mode='reflect') for 1-D input (radlite.py:3148-3151): out[i] =
    sum_j k[j] a[i + (n // 2) - j] with the input mirrored about its edges (d c b a | a b c d | d c b a);
    products accumulated in the order of ascending j."""
    n, L = len(k), len(a)
    out = np.zeros(L)
    idx = np.arange(L)
    for j in range(n):
        src = idx + (n // 2) - j
        src = np.where(src < 0, -src - 1, src)
        src = np.where(src >= L, 2 * L - 1 - src, src)
        out = out + k[j] * a[src]
    return out


def process_spectrum(vel, flux, freq, dist_pc, obsres, vsampling):
    """radlite.py:3001-3184 (_process_spectrum), interpolation 'linear'.  vel, flux: [nl, nfr] as read from
    linespectrum_moldata_*.dat (velocity ascending, F_nu at 1 pc in erg/s/cm^2/Hz); freq [nl] Hz.
    Returns dict(wavelength, spectrum, emission, continuum [Jy], frequency)."""
    vel, flux, freq = np.asarray(vel), np.asarray(flux), np.asarray(freq)
    nl = len(freq)
    G = synthesis_grids(vel, freq, obsres, vsampling)
    fullmu, outmu, box = G["fullmu"], G["outmu"], G["boxwidth"]
    fullem = np.zeros(len(fullmu))
    emonly, contcen = [None] * nl, [None] * nl
    for a in range(nl):  # :3076-3084
        ys, xs = [flux[a][0], flux[a][-1]], [vel[a][0], vel[a][-1]]
        emonly[a] = flux[a] - _interp1d_linear(xs, ys, vel[a])
        contcen[a] = _interp1d_linear(xs, ys, 0.0)[0]
    for a in range(nl):  # :3091-3110
        muolds = ((np.concatenate([[-1 * box], vel[a], [box]]) * 1.0E9 / freq[a]) + (cinmu0 / 1.0 / freq[a]))
        emolds = np.concatenate([[emonly[a][0]], emonly[a], [emonly[a][-1]]])
        inds = np.where(((fullmu <= muolds[-1]) & (fullmu >= muolds[0])))[0]
        fullem[inds] = fullem[inds] + _interp1d_linear(muolds, emolds, fullmu[inds])
    mus = np.array([(cinmu0 / 1.0 / f) for f in freq])  # :3115-3127
    order = np.argsort(mus)
    musort = np.concatenate([[G["mumin"]], mus[order], [G["mumax"]]])
    csort = np.array(contcen)[order]
    csort = np.concatenate([[csort[0]], csort, [csort[-1]]])
    fullcont = _interp1d_linear(musort, csort, fullmu)
    fullem = fullem * 1.0E23 / (dist_pc ** 2)  # :3132-3134
    fullcont = fullcont * 1.0E23 / (dist_pc ** 2)
    fully = fullcont + fullem
    kern = gauss_kernel(obsres, G["vres"])  # :3141-3151
    resem = convolve_reflect(fullem, kern) / 1.0 / np.sum(kern)
    resy = convolve_reflect(fully, kern) / 1.0 / np.sum(kern)
    return dict(wavelength=outmu, emission=_interp1d_linear(fullmu, resem, outmu),  # :3158-3167
                spectrum=_interp1d_linear(fullmu, resy, outmu), continuum=_interp1d_linear(fullmu, fullcont, outmu),
                frequency=c0 / (outmu * 1.0E-4), fullmu=fullmu, fullem=fullem, fullcont=fullcont)

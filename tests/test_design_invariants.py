"""CPU checks of the two properties of the reference's arithmetic that the CUDA design leans on
(DESIGN.md 4.1, 4.3), done in reference-ordered double arithmetic on the oracle's own node data:

* the argument of the line profile does not depend on the line (ztile_kernel evaluates the Gaussian once
  per ray, node and channel and shares it between the lines of a tile);
* what a ray has accumulated before an optically thick dust layer does not reach the result (the kernels
  start a ray behind tau_dust > 150 from the observer's end).

PARITY UNPINNED like everything built on the oracle: these pin the design's claims to the restated
algorithm, not to the Fortran binary."""
import math

import numpy as np
import pytest

from radlite_b200 import synth


def _qdr_src_2(inten, js1, alp1, js2, alp2, ds):
    """transfer.F:1498-1571 in Python floats (IEEE double, same operation order as the oracle)."""
    dtau1 = 0.5 * (alp1 + alp2) * ds
    theomax = 0.5 * (js1 + js2) * ds
    if dtau1 > 1.0e-6:
        xp = math.exp(-dtau1)
        e0 = 1.0 - xp
        e1 = dtau1 - e0
        b = e1 / dtau1
        a = e0 - b
    else:
        a = 0.5 * dtau1
        b = 0.5 * dtau1
        xp = 1.0 - dtau1
    if alp1 > 0.0:
        src1 = js1 / alp1
    elif alp2 > 0.0:
        src1 = js2 / alp2
    else:
        src1 = 0.0
    if alp2 > 0.0:
        src2 = js2 / alp2
    elif alp1 > 0.0:
        src2 = js1 / alp1
    else:
        src2 = 0.0
    q = a * src1 + b * src2 if dtau1 > float(np.float32(1e-9)) else theomax
    q = min(q, theomax)
    return inten * xp + q


def _walk(m, il, v, s, icross, iradius, dnu, start=1):
    """charintline (telescope.F:3889-4312) + integrate_element_linedust (line.F:4515-4624) for one ray and
    channel, SUBGRID off, from segment ``start`` on (start=1: the whole ray, intensity from the outer BC 0;
    start>1: intensity 0 and the profile state carried into that segment as the full walk carries it)."""
    nu0 = m.linefreq[il]
    aud = m.aud[il]
    gu, gd = m.gdeg[m.lev_up[il] - 1], m.gdeg[m.lev_down[il] - 1]
    bud = 6.78171833781e46 * aud / nu0 ** 3
    bdu = bud * gu / gd
    n = len(s)

    def line_terms(i, lwav):
        aa = 3.33567e-6 * nu0 * lwav
        u = (dnu - nu0 * v["dvmu"][i]) / aa
        phi = 0.56419583546 / aa * math.exp(-(u * u))
        return (5.27296241956e-28 * nu0 * v["nup"][i] * aud * phi,
                5.27296241956e-28 * nu0 * phi * (v["ndown"][i] * bdu - v["nup"][i] * bud))

    inten = 0.0
    init = True
    srcl0 = alpl0 = 0.0
    if start > 1:  # state at node start-1 = end state of segment start-1 (its mean width)
        srcl0, alpl0 = line_terms(start - 1, 0.5 * (v["lw"][start - 2] + v["lw"][start - 1]))
        init = False
    for i in range(start, n):
        ds = s[i] - s[i - 1]
        if icross[i] == 1 and icross[i - 1] == 1 and iradius[i] == 1 and iradius[i - 1] == 1:
            ds = 0.0  # telescope.F:4095-4104 (inner hole)
            init = True
        lwav = 0.5 * (v["lw"][i - 1] + v["lw"][i])
        if init:
            srcl0, alpl0 = line_terms(i - 1, lwav)
            init = False
        srcl1, alpl1 = line_terms(i, lwav)
        inten = _qdr_src_2(inten, v["srcd"][i - 1] + srcl0, v["alpd"][i - 1] + alpl0, v["srcd"][i] + srcl1,
                           v["alpd"][i] + alpl1, ds)
        srcl0, alpl0 = srcl1, alpl1
    return inten


def test_profile_argument_is_line_independent():
    """line.F:462-469 + 4559-4566 + 2301: u = (dnu_k - nu0 dvmu) / (3.33567e-6 nu0 lwav) with
    dnu_k = -passb + k 2 passb/(nfr-1), passb = 3.33567e-6 nu0 vmax.  nu0 cancels: every line of a render sees
    the same u on the same channel -- to a few ulps of the velocity, i.e. ~1e-14 on exp(-u^2)."""
    m = synth.config(2, nr=30, nth=12, nphi=8, nrext=-6, nlines=40)
    nfr, vmax = m.nfr, m.passband
    pv = 3.33567e-6 * vmax
    velz = (0.0 - pv) + np.arange(nfr) * (2.0 * pv / (nfr - 1.0))  # rl_capi.cu: the shared velocity grid
    rng = np.random.default_rng(7)
    worst_u = worst_e = 0.0
    for nu0 in m.linefreq:
        passb = 3.33567e-6 * nu0 * vmax
        dnu = (0.0 - passb) + np.arange(nfr) * (2.0 * passb / (nfr - 1.0))
        for _ in range(20):
            dvmu = rng.uniform(-1.5e-4, 1.5e-4)  # |v| up to 45 km/s
            lwav = rng.uniform(0.3, 4.0)
            u_line = (dnu - nu0 * dvmu) / (3.33567e-6 * nu0 * lwav)
            u_shared = (velz - dvmu) / (3.33567e-6 * lwav)
            near = np.abs(u_line) < 19.0  # beyond: exp(-u^2) < 1e-150, flushed to 0 by the kernels
            worst_u = max(worst_u, np.abs(u_line - u_shared)[near].max())
            e1, e2 = np.exp(-u_line[near] ** 2), np.exp(-u_shared[near] ** 2)
            worst_e = max(worst_e, (np.abs(e1 - e2) / e1).max())
    assert worst_u < 1e-11 and worst_e < 1e-10, (worst_u, worst_e)


def test_opaque_wall_is_invisible_in_reference_arithmetic(oracle_cls):
    """Walk rays of a disk with an opaque midplane twice in reference-ordered arithmetic: from the far end,
    and from the first segment that has tau_dust > 150 in front of it (intensity 0, carried profile state of
    that node).  Bit-identical results; and the full walk reproduces the oracle's pixel."""
    m = synth.config(2, nr=40, nth=16, nphi=8, nrext=-6, nlines=6)
    m.subgrid = 0
    o = oracle_cls()
    o.load_model(m)
    il = 2
    ref = o.render(il + 1, 1, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    nu0 = m.linefreq[il]
    passb = 3.33567e-6 * nu0 * m.passband
    dnu_all = (0.0 - passb) + np.arange(m.nfr) * (2.0 * passb / (m.nfr - 1.0))
    walls = checked = 0
    for iray in range(2, m.nray + 1, 3):
        t = o.trajectory(iray)
        s = t["s"]
        if len(s) < 4:
            continue
        v = o.node_values(iray, il + 1)
        ring, phi = 1 + (iray - 2) // m.nphi, (iray - 2) % m.nphi
        ds = np.diff(s)
        hole = (t["icross"][1:] == 1) & (t["icross"][:-1] == 1) & (t["iradius"][1:] == 1) & (t["iradius"][:-1] == 1)
        ds[hole] = 0.0  # telescope.F:4095-4104: the inner hole is vacuum
        seg_tau = 0.5 * ds * (v["alpd"][:-1] + v["alpd"][1:])
        cum = np.cumsum(seg_tau[::-1])[::-1]  # cum[k]: tau of segments k+1 .. N-1 (segment k+1 joins nodes k, k+1)
        over = np.nonzero(cum > 150.0)[0]
        start = int(over.max()) + 1 if over.size else 1
        chans = np.nonzero(ref["cmask"][0, ring, phi] == 1)[0][:6]
        for ch in chans:
            full = _walk(m, il, v, s, t["icross"], t["iradius"], dnu_all[ch])
            want = ref["image"][0, ring, phi, ch]
            assert abs(full - want) <= 1e-12 * abs(want), (iray, ch, full, want)
            checked += 1
            if start > 2:
                cut = _walk(m, il, v, s, t["icross"], t["iradius"], dnu_all[ch], start=start)
                assert cut == full, (iray, ch, start, cut, full)
                walls += 1
    assert checked > 50 and walls > 10, (checked, walls)

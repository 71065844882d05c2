"""Builds tests/golden/tutorial_120x100.npz from the reference's ONLY real fixture:
the complete RADMC input set of the pyradlite tutorial,
  /root/reference/DOCS/DOCS_VERSION_1-3/files_for_tutorials/radmc_outputs/   (120 x 100 grid, 130 frequencies)
  /root/reference/DOCS/DOCS_VERSION_1-3/files_for_tutorials/input_radlite.json (camera, passband, abundance ...)
  /root/reference/LAMDA/12CO_lamda.dat                                         (levels and lines, 4.6-4.7 um: 26 lines)
Run it in the build container (the GPU box has no /root/reference):  python tests/golden/make_tutorial_fixture.py
The files are parsed as the Fortran readers parse them and stored as arrays -- inputs
only: the reference ships no expected output.  The tutorial takes its line list from HITRAN (05_hit12.par, not in
the reference tree: HITRAN/ is git-ignored) and its partition sum from HITRAN's ParSum.dat (absent too); the
fixture takes both from the LAMDA file instead: the 26 lines of the same band, Q(T) = sum g exp(-E h c / k T)
over all 318 LAMDA levels on a 1 K grid.  That changes the line list's provenance, not the arithmetic under test.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import driver_np as D  # noqa: E402

REF = "/root/reference"
TUT = os.path.join(REF, "DOCS/DOCS_VERSION_1-3/files_for_tutorials")
RAD = os.path.join(TUT, "radmc_outputs")


def numbers(path):
    return [float(t) for t in open(path).read().split()]


def main():
    v = numbers(os.path.join(RAD, "radius.inp"))
    nr = int(v[0])
    r = np.array(v[1:1 + nr])
    v = numbers(os.path.join(RAD, "theta.inp"))
    nth, mirror = int(v[0]), int(v[1])
    theta = np.array(v[2:2 + nth])
    assert mirror == 1
    v = numbers(os.path.join(RAD, "frequency.inp"))
    nf = int(v[0])
    freq = np.array(v[1:1 + nf])
    v = numbers(os.path.join(RAD, "dustdens.inp"))  # nspec nr nth imirt ; per species, per ir, per it
    assert (int(v[0]), int(v[1]), int(v[2])) == (1, nr, nth)
    dust_rho = np.array(v[4:4 + nr * nth]).reshape(nr, nth, 1)
    v = numbers(os.path.join(RAD, "dusttemp_final.dat"))  # nspec nr nth imirt ; per species: nsize ; per size ...
    assert (int(v[0]), int(v[1]), int(v[2])) == (1, nr, nth) and int(v[4]) == 1
    dust_temp = np.array(v[5:5 + nr * nth]).reshape(nr, nth, 1, 1)
    v = numbers(os.path.join(RAD, "dustopac_1.inp"))  # nf nsize ; kappa_abs[nf] ; kappa_scat[nf]
    assert (int(v[0]), int(v[1])) == (nf, 1)
    kabs = np.array(v[2:2 + nf]).reshape(1, 1, nf)
    kscat = np.array(v[2 + nf:2 + 2 * nf]).reshape(1, 1, nf)
    v = numbers(os.path.join(RAD, "starinfo.inp"))
    rstar, mstar, tstar = v[1], v[2], v[3]
    v = numbers(os.path.join(RAD, "starspectrum.inp"))
    assert int(v[0]) == nf
    ss = np.array(v[1:1 + 2 * nf]).reshape(nf, 2)
    assert np.allclose(ss[:, 0], freq, rtol=1e-3)  # star.F:499-528
    starspec_cont = 3.0308410e36 * ss[:, 1] / rstar ** 2  # F_nu at 1 pc -> surface intensity (star.F:510)
    gastodust = None
    for ln in open(os.path.join(RAD, "problem_params.pro")):
        if ln.strip().startswith("gastodust"):
            gastodust = float(ln.split("=")[1].split(";")[0].replace("d", "e"))
    par = {k: d["value"] for k, d in json.load(open(os.path.join(TUT, "input_radlite.json"))).items()}
    hit = json.load(open(os.path.join(TUT, "data_hitran.json")))[par["molname"]]

    # LAMDA: levels and the lines of the band
    L = open(os.path.join(REF, "LAMDA/12CO_lamda.dat")).read().splitlines()
    nlev = int(L[5])
    lev = np.array([[float(t) for t in ln.split()] for ln in L[7:7 + nlev]])
    e_all, g_all, v_all, j_all = lev[:, 1], lev[:, 2], lev[:, 3].astype(int), lev[:, 4].astype(int)
    i0 = 7 + nlev
    nlin = int(L[i0 + 1])
    lin = np.array([[float(t) for t in ln.split()[:6]] for ln in L[i0 + 3:i0 + 3 + nlin]])
    lam = 2.99792458e5 / lin[:, 4]  # GHz -> um
    sel = np.where((lam >= par["min_mu"]) & (lam <= par["max_mu"]))[0]
    sel = sel[np.argsort(lin[sel, 4])]  # radlite.py:1957-1962 sorts the core's lines by wavenumber
    up_all, low_all = lin[sel, 1].astype(int) - 1, lin[sel, 2].astype(int) - 1
    e_u, g_u, v_u, low, up = D.unique_levels(e_all[low_all], e_all[up_all], g_all[low_all], g_all[up_all],
                                             v_all[low_all], v_all[up_all])
    j_u = np.array([j_all[np.argmin(np.abs(e_all - e))] for e in e_u])
    aud = np.array([float(f"{a:.3e}") for a in lin[sel, 3]])  # moldata.dat carries A as E12.3 (line.F:1933)
    psum_temp = np.arange(1.0, 3001.0)
    psum = (g_all[None, :] * np.exp(-e_all[None, :] * D.h0 * D.c0 / D.kB0 / psum_temp[:, None])).sum(axis=1)

    out = os.path.join(HERE, "tutorial_120x100.npz")
    np.savez_compressed(
        out, r=r, theta=theta, cont_freq_nu=freq, dust_rho=dust_rho, dust_temp=dust_temp, kappa_abs=kabs,
        kappa_scat=kscat, rstar=rstar, mstar=mstar, tstar=tstar, starspec_cont=starspec_cont, gastodust=gastodust,
        molweight=float(hit["molweight"]), max_abun=par["max_abun"], min_abun=par["min_abun"],
        alpha=par["alpha"], gamma=par["gamma"], mu=par["mu"], incl=par["incl"], cir_np=par["cir_np"],
        b_per_r=par["b_per_r"], b_extra=par["b_extra"], passband=par["passband"], vsampling=par["vsampling"],
        vlsr=par["vlsr"], ener_cm=e_u, gdeg=g_u, lev_v=v_u.astype(np.int32), lev_j=j_u.astype(np.int32),
        lev_up=(up + 1).astype(np.int32), lev_down=(low + 1).astype(np.int32), aud=aud,
        psum_temp=psum_temp, psum=psum)
    print("wrote", out, os.path.getsize(out), "bytes;", len(sel), "lines,", len(e_u), "levels, grid", nr, "x", nth)


if __name__ == "__main__":
    main()

"""File formats either side of the hot path: the stand-alone host program reads a
RADLite working directory written the way pyradlite / RADMC write it, and writes linespectrum_<mol>.dat
and lineposvelcirc_*.dat the way gfortran does.  No GPU and no compute here: parsing, formatting, and
the writers driven from stored results (--replay)."""
import os
import tempfile

import numpy as np
import pytest

import workdir as wd
from helpers import clone, tiny
from radlite_b200 import synth

pytestmark = pytest.mark.skipif(not os.path.exists(wd.HOST), reason="radlite_b200_host not built")


@pytest.mark.parametrize("args,expect", [
    (("e", 13, 6, 1.2345678e-7), " 0.123457E-06"), (("e", 13, 6, -3.5e12), "-0.350000E+13"),
    (("e", 14, 9, 6.5e13), ".650000000E+14"), (("e", 10, 5, 0.0), ".00000E+00"), (("e", 12, 4, 1.0), "  0.1000E+01"),
    (("f", 7, 3, 15.0), " 15.000"), (("e", 13, 6, 1e-105), " 0.100000-104"), (("e", 12, 6, 1.1306358e12), "0.113064E+13"),
    (("e", 10, 4, 3.21e-9), "0.3210E-08"), (("i", 5, 0, 42), "   42"), (("e", 13, 6, 9.9999996e-5), " 0.100000E-03"),
    (("g", 0, 0, 6.5e13), "    65000000000000.000     "), (("g", 0, 0, -12.5), "   -12.500000000000000     "),
    (("s", 12, 4, 6.5e13), "  6.5000E+13"), (("s", 12, 5, -1.25e-7), "-1.25000E-07"), (("s", 12, 5, 0.0), " 0.00000E+00"),
    (("s", 12, 5, 3.3e-105), " 3.30000-105"), (("s", 12, 4, 1.0), "  1.0000E+00"),
])
def test_fortran_edit_descriptors(args, expect):
    import subprocess
    out = subprocess.run([wd.HOST, "--fmt", *[str(a) for a in args]], capture_output=True, text=True).stdout
    assert out.rstrip("\n") == "[" + expect + "]"


def _parse(m, **kw):
    d = tempfile.mkdtemp(prefix="rlwd_")
    wd.write_workdir(m, d, **kw)
    wd.run_host(d, "--parse-only", "--dump", os.path.join(d, "model.bin"))
    return d, wd.load_dump(os.path.join(d, "model.bin"))


def test_reads_a_pyradlite_style_working_directory():
    m = tiny(2, nlines=4)
    d, dmp = _parse(m)
    assert np.allclose(dmp["r"], m.r, rtol=1e-7) and np.allclose(dmp["theta"], m.theta, atol=1e-8)
    assert np.allclose(dmp["cont_freq_nu"], m.cont_freq_nu, rtol=1e-6)
    assert np.allclose(dmp["rho"], m.rho, rtol=1e-8) and np.allclose(dmp["abund"], m.abund, rtol=1e-8)
    assert np.allclose(dmp["vel"], m.vel, atol=1e-8) and np.allclose(dmp["linewidth"], m.linewidth, atol=1e-8)
    assert np.allclose(dmp["dust_rho"], m.dust_rho, rtol=2e-7) and np.allclose(dmp["dust_temp"], m.dust_temp, rtol=1e-14)
    assert np.allclose(dmp["kappa_abs"], m.kappa_abs, rtol=1e-7) and np.allclose(dmp["kappa_scat"], m.kappa_scat, rtol=1e-7)
    assert np.array_equal(dmp["popul"], m.popul)  # written with repr: exact
    assert np.array_equal(dmp["lev_up"], m.lev_up) and np.array_equal(dmp["lev_down"], m.lev_down)
    assert np.array_equal(dmp["gdeg"], m.gdeg) and np.allclose(dmp["ener_cm"], m.ener_cm, atol=5e-5)
    assert np.allclose(dmp["aud"], m.aud, rtol=5e-4)  # E12.3
    # line.F:1903,1981: frequencies come from the level energies, not from the FREQ column
    nu = 1.509160e26 * (1.986468498e-16 * dmp["ener_cm"][m.lev_up - 1] - 1.986468498e-16 * dmp["ener_cm"][m.lev_down - 1])
    assert np.array_equal(dmp["linefreq"], nu)
    # star.F:510: surface intensity from F_nu at 1 pc
    assert np.allclose(dmp["starspec_cont"], m.starspec_cont, rtol=1e-6)
    sc = dmp["scalars"]
    assert sc[0] == m.umass_av and abs(sc[1] - m.rstar) < 1e-7 * m.rstar
    assert (sc[2], sc[3], sc[4], sc[5], sc[6]) == (m.out_itype, 2, m.nphi, m.dbdr, m.nrext)
    assert sc[7] == m.incl_deg and sc[9] == m.nfr and abs(sc[10] - m.passband) < 1e-12  # main.F:209-211
    assert sc[11] == 3.08572e18 and sc[12] == 4 and sc[13] == 1


def test_radlite_inp_comment_stripping_and_truncation():
    """tools.F:6-16: lines starting with # ; % = are dropped and every line is cut at 16 characters."""
    m = tiny(1)
    d = tempfile.mkdtemp(prefix="rlwd_")
    wd.write_workdir(m, d)
    p = os.path.join(d, "radlite.inp")
    txt = open(p).read().replace("41\t\tNr of mu", "# a comment\n; another\n% and another\n41\t\tNr of mu 999999")
    open(p, "w").write(txt)
    wd.run_host(d, "--parse-only", "--dump", os.path.join(d, "m.bin"))
    assert wd.load_dump(os.path.join(d, "m.bin"))["scalars"][4] == m.nphi


@pytest.mark.parametrize("breakit,code", [
    (lambda d: os.remove(os.path.join(d, "radlite.inp")), 13),
    (lambda d: os.remove(os.path.join(d, "density.inp")), 1),
    (lambda d: os.remove(os.path.join(d, "dusttemp.info")), 13),
    (lambda d: open(os.path.join(d, "velocity.inp"), "w").write("3\t3\n"), 325 & 255),
    (lambda d: os.remove(os.path.join(d, "levelpop.info")), 13),
    (lambda d: open(os.path.join(d, "radlite.inp"), "w").write("99\n"), 13),
    (lambda d: os.remove(os.path.join(d, "starspectrum.inp")), 13),
])
def test_reference_stop_codes_on_bad_input(breakit, code):
    m = tiny(1)
    d = tempfile.mkdtemp(prefix="rlwd_")
    wd.write_workdir(m, d)
    breakit(d)
    p = wd.run_host(d, "--parse-only", check=False)
    assert p.returncode == code
    assert not os.path.exists(os.path.join(d, "radlite.success"))


def test_no_gpu_means_no_success_marker():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = tiny(1)
    d = tempfile.mkdtemp(prefix="rlwd_")
    wd.write_workdir(m, d)
    p = wd.run_host(d, check=False)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr
    assert not os.path.exists(os.path.join(d, "radlite.success"))


def test_writers_roundtrip_through_the_pyradlite_parser(oracle_cls):
    """linespectrum_<mol>.dat / lineposvelcirc written from stored oracle results parse back with the
    drivers' readers to the 6 (4) significant digits the formats carry."""
    m = clone(tiny(2, nlines=3), vlsr=4.5)
    d, dmp = _parse(m)
    o = oracle_cls()
    o.load_model(wd.model_from_dump(m, dmp))
    ref = o.render(1, 3, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    rays_r, ri = o.rings()
    wd.write_records(os.path.join(d, "res.bin"), flux=ref["flux"], velo=ref["velo"], image=ref["image"],
                     cmask=ref["cmask"].astype(np.int32), rays_r=rays_r, imcir_ri=ri)
    wd.run_host(d, "--replay", os.path.join(d, "res.bin"))
    out = wd.read_linespectrum(os.path.join(d, "linespectrum_moldata.dat"))
    assert out["numlines"] == 3 and out["maxnumpoints"] == m.nfr and out["dist"] == 1.0
    assert abs(out["incl"] - m.incl_deg) < 1e-3 and out["vlsr"] == 4.5
    raw = out["raw"]
    assert raw[0] == " 1\n" and raw[1] == " 1\n" and len(raw[2]) == 81 and raw[2].startswith("moldata")
    assert raw[3].startswith("./moldata.dat") and raw[7] == "\n"
    for k, ln in enumerate(out["lines"]):
        assert ln["lev"] == [m.lev_up[k], m.lev_down[k]] and ln["beam"] == 0.0
        assert abs(ln["freq"] - dmp["linefreq"][k]) <= 1e-9 * dmp["linefreq"][k]
        # rows are written channel nfr..1: ascending velocity = -c dnu/nu0 + v_lsr (telescope.F:1771-1775)
        v = -2.99792458e5 * ref["velo"][k][::-1] + 4.5
        assert np.allclose(ln["vel"], v, rtol=1e-6, atol=5e-6 * np.abs(v).max()) and np.all(np.diff(ln["vel"]) > 0)
        assert np.allclose(ln["flux"], ref["flux"][k][::-1], rtol=1.0e-6)
    assert len(raw[9]) == 15 and raw[9][0] == "."  # (e14.9) has no room for the leading zero
    nrr = ref["image"].shape[1] - 1
    for k in range(3):
        c = wd.read_imcir(os.path.join(d, f"lineposvelcirc_moldata_{k + 1}.dat"))
        assert (c["nfr"], c["nphi"], c["nrr"]) == (m.nfr, m.nphi, nrr)
        assert np.allclose(c["ri"], ri[1:], rtol=1e-5) and np.allclose(c["r"][:-1], rays_r[1:], rtol=1e-5)
        assert np.allclose(c["vel"], ref["velo"][k] * 2.99792458e5, rtol=1e-12)
        assert np.allclose(c["centre"], ref["image"][k, 0, 0], rtol=1e-15)
        assert np.allclose(c["image"], np.transpose(ref["image"][k, 1:], (2, 1, 0)), rtol=1.0e-4)
        assert np.array_equal(c["cmask"], np.transpose(ref["cmask"][k, 1:], (2, 1, 0)))

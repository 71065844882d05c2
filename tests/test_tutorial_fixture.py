"""The reference's only real fixture -- the RADMC input set of the pyradlite tutorial (120 x 100 grid, 26 12CO
lines at 4.6-4.7 um, 28 351 rays; published wall time 248 s on 3 laptop cores,
tutorial_RadliteModel.ipynb:962) -- through the oracle, the file readers and (on the GPU box) the CUDA path."""
import os
import shutil
import tempfile
import time

import numpy as np
import pytest

import workdir as wd
from helpers import rel_err, tutorial_model
from radlite_b200 import synth

RADMC = "/root/reference/DOCS/DOCS_VERSION_1-3/files_for_tutorials/radmc_outputs"


def test_fixture_shape_and_oracle_sanity(oracle_cls):
    m = tutorial_model()
    assert (len(m.r), len(m.theta), m.nlines, m.nfr, m.nray) == (120, 100, 26, 94, 28351)
    lam = 2.99792458e14 / m.linefreq
    assert lam.min() >= 4.6 and lam.max() <= 4.7
    assert np.allclose(m.popul.sum(axis=-1).max(), m.popul.sum(axis=-1).max()) and np.all(m.popul >= 0)
    o = oracle_cls()
    o.load_model(m)
    o.set_ring_sample(20, m.nrr, 40)  # a few camera rings: a second of oracle
    ref = o.render(1, 2, m.nfr, m.passband, synth.PARSEC, want_image=True)
    assert np.all(np.isfinite(ref["image"])) and ref["image"].max() > 0 and np.all(ref["image"] >= 0)


@pytest.mark.skipif(not os.path.isdir(RADMC) or not os.path.exists(wd.HOST),
                    reason="needs the reference tree (build container) and radlite_b200_host")
def test_readers_on_the_reference_files_themselves():
    """The RADMC files exactly as the reference ships them (not re-written by this repo's test writers), next
    to the driver-written gas / molecule files: radlite_b200_host must parse them to the fixture's arrays."""
    m = tutorial_model()
    d = tempfile.mkdtemp(prefix="rltut_")
    wd.write_workdir(m, d)  # driver-side files (and stand-ins for the RADMC ones ...)
    for name in ("radius.inp", "theta.inp", "frequency.inp", "dustdens.inp", "dusttemp.info", "dusttemp_final.dat",
                 "dustopac.inp", "dustopac_1.inp", "starinfo.inp", "starspectrum.inp", "line.inp"):
        shutil.copy(os.path.join(RADMC, name), os.path.join(d, name))  # ... replaced by the reference's own files
    wd.run_host(d, "--parse-only", "--dump", os.path.join(d, "model.bin"))
    dmp = wd.load_dump(os.path.join(d, "model.bin"))
    for k in ("r", "theta", "cont_freq_nu", "dust_rho", "dust_temp", "kappa_abs", "kappa_scat"):
        assert np.array_equal(dmp[k], getattr(m, k)), k
    assert np.allclose(dmp["starspec_cont"], m.starspec_cont, rtol=1e-15)
    sc = dmp["scalars"]
    assert sc[0] == 2.3 and sc[1] == m.rstar  # line.inp: average molecular weight; starinfo.inp
    assert np.array_equal(dmp["lev_up"], m.lev_up) and np.array_equal(dmp["linefreq"], m.linefreq)


@pytest.mark.gpu
def test_tutorial_run_on_the_gpu(renderer_cls, oracle_cls):
    """All 26 lines on the B200 against the oracle: one full line (every ray, every channel) and ring samples of
    two more; then the same run through the stand-alone host program and the text files."""
    m = tutorial_model()
    g = renderer_cls(0)
    g.load_model(m)
    t0 = time.perf_counter()
    out = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True)
    t_gpu = time.perf_counter() - t0
    assert np.all(np.isfinite(out["flux"])) and np.all(out["flux"] > 0)
    o = oracle_cls()
    o.load_model(m)
    ref = o.render(13, 1, m.nfr, m.passband, synth.PARSEC, want_image=True)
    assert rel_err(out["flux"][12:13], ref["flux"]).max() < 1e-5
    assert abs(out["flux"][12].sum() - ref["flux"].sum()) < 1e-6 * ref["flux"].sum()
    assert rel_err(out["image"][12], ref["image"][0]).max() < 1e-5
    for il in (1, 26):
        for ir in (30, 100, 189):
            o.set_ring_sample(ir, ir, 1)
            r1 = o.render(il, 1, m.nfr, m.passband, synth.PARSEC, want_image=True)["image"][0]
            assert rel_err(out["image"][il - 1, ir], r1[ir]).max() < 1e-5, (il, ir)
    o.set_ring_sample(0, 0, 1)
    print(f"\ntutorial fixture: 26 lines rendered in {t_gpu:.3f} s incl. cube download "
          f"(published: 248 s on 3 cores, tutorial_RadliteModel.ipynb:962)")
    # through the files: driver-style working directory -> host program -> linespectrum_moldata.dat
    d = tempfile.mkdtemp(prefix="rltut_")
    wd.write_workdir(m, d)
    wd.run_host(d, "--dump", os.path.join(d, "model.bin"))
    spec = wd.read_linespectrum(os.path.join(d, "linespectrum_moldata.dat"))
    m2 = wd.model_from_dump(m, wd.load_dump(os.path.join(d, "model.bin")))
    g2 = renderer_cls(0)
    g2.load_model(m2)
    f2 = g2.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"]
    assert spec["numlines"] == 26
    for k, ln in enumerate(spec["lines"]):
        assert np.allclose(ln["flux"], f2[k][::-1], rtol=1.0e-6)
    # the text files carry 8 significant digits of the gas fields: the spectrum moves by less than that
    assert rel_err(f2, out["flux"]).max() < 1e-5


@pytest.mark.gpu
def test_driver_side_kernels(renderer_cls):
    """rl_set_lines_lte and rl_synthesize_spectrum against the numpy restatement of pyradlite's arithmetic."""
    from oracle import driver_np as D
    from radlite_b200 import driver
    m = tutorial_model()
    g = renderer_cls(0)
    g.load_model(m)
    a = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"]
    driver.set_lines_lte(g, m.lev_up, m.lev_down, m.linefreq, m.aud, m.gdeg, m.ener_cm, m.tgas,
                         m.extra["psum_temp"], m.extra["psum"])
    g.set_dust(m.nsize, m.cont_freq_nu, m.kappa_abs, m.kappa_scat, m.dust_rho, m.dust_temp, m.scati_src)
    b = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"]
    assert rel_err(b, a).max() < 1e-11  # populations differ by the exp's last bit at most
    vel, flux = driver.spectrum_from_render(m, a)
    for obsres, vsamp in ((3.0, 1.5), (12.0, 3.0)):
        got = driver.synthesize_spectrum(g, vel, flux, m.linefreq, 140.0, obsres, vsamp)
        want = D.process_spectrum(vel, flux, m.linefreq, 140.0, obsres, vsamp)
        assert np.array_equal(got["wavelength"], want["wavelength"])
        for k in ("spectrum", "emission", "continuum"):
            assert np.allclose(got[k], want[k], rtol=1e-12, atol=1e-14 * np.abs(want["spectrum"]).max()), k

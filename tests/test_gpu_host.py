"""End to end through the stand-alone host program on the GPU box: a working directory written the way
the drivers write it -> radlite_b200_host (C ABI -> CUDA path) -> linespectrum_<mol>.dat /
lineposvelcirc_*.dat parsed the way pyradlite / read_imcir.pro parse them, against the CUDA path called
directly and against the CPU oracle fed the same parsed model."""
import os
import tempfile

import numpy as np
import pytest

import workdir as wd
from helpers import clone, tiny
from radlite_b200 import synth

pytestmark = pytest.mark.gpu


def test_host_program_end_to_end(renderer_cls, oracle_cls):
    m = clone(tiny(2, nlines=4), vlsr=-3.0)
    d = tempfile.mkdtemp(prefix="rlwd_")
    wd.write_workdir(m, d)
    p = wd.run_host(d, "--dump", os.path.join(d, "model.bin"))
    assert os.path.exists(os.path.join(d, "radlite.success"))
    assert open(os.path.join(d, "radlite.success")).read() == " 1\n"
    assert p.stdout.count("Rendered spectrum of line") == 4
    out = wd.read_linespectrum(os.path.join(d, "linespectrum_moldata.dat"))
    m2 = wd.model_from_dump(m, wd.load_dump(os.path.join(d, "model.bin")))
    g = renderer_cls(0)
    g.load_model(m2)
    gpu = g.render(1, 4, m.nfr, m.passband, synth.PARSEC)
    o = oracle_cls()
    o.load_model(m2)
    ref = o.render(1, 4, m.nfr, m.passband, synth.PARSEC)
    assert out["numlines"] == 4 and out["maxnumpoints"] == m.nfr
    for k, ln in enumerate(out["lines"]):
        # the text carries 6 significant digits (E13.6)
        assert np.allclose(ln["flux"], gpu["flux"][k][::-1], rtol=1.0e-6)
        assert np.allclose(ln["flux"], ref["flux"][k][::-1], rtol=1.1e-5)
        assert np.allclose(ln["vel"], -2.99792458e5 * gpu["velo"][k][::-1] - 3.0, rtol=1e-6, atol=1e-4)


def test_host_program_imcir_build(renderer_cls):
    """RADlite_imcir behaviour (--imcir): the circular image cube per line next to the spectrum."""
    m = tiny(1)
    d = tempfile.mkdtemp(prefix="rlwd_")
    wd.write_workdir(m, d)
    wd.run_host(d, "--imcir", "--dump", os.path.join(d, "model.bin"))
    m2 = wd.model_from_dump(m, wd.load_dump(os.path.join(d, "model.bin")))
    g = renderer_cls(0)
    g.load_model(m2)
    gpu = g.render(1, 1, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    c = wd.read_imcir(os.path.join(d, "lineposvelcirc_moldata_1.dat"))
    nrr = gpu["image"].shape[1] - 1
    assert (c["nfr"], c["nphi"], c["nrr"]) == (m.nfr, m.nphi, nrr)
    assert np.allclose(c["centre"], gpu["image"][0, 0, 0], rtol=1e-15)
    assert np.allclose(c["image"], np.transpose(gpu["image"][0, 1:], (2, 1, 0)), rtol=1.0e-4)
    assert np.array_equal(c["cmask"], np.transpose(gpu["cmask"][0, 1:], (2, 1, 0)))
    spec = wd.read_linespectrum(os.path.join(d, "linespectrum_moldata.dat"))
    assert np.allclose(spec["lines"][0]["flux"], gpu["flux"][0][::-1], rtol=1.0e-6)


def test_host_program_fails_like_the_reference(renderer_cls):
    """A stop inside the library (here: rstar > R_1, telescope.F stop 91991) ends the run without
    radlite.success and with the reference's code (mod 256, as a process exit status)."""
    m = tiny(1)
    d = tempfile.mkdtemp(prefix="rlwd_")
    wd.write_workdir(clone(m, rstar=2 * m.r[0]), d)
    p = wd.run_host(d, check=False)
    assert p.returncode == 91991 % 256 and "91991" in p.stderr
    assert not os.path.exists(os.path.join(d, "radlite.success"))

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


def test_host_program_position_velocity_cube(renderer_cls, oracle_cls):
    """linespectrum.inp command 2 (main.F:1062-1075 -> calc_write_line_posvel, telescope.F:1828): the host
    program renders the rectangular cube and writes lineposvel_<mol>_<n>.dat; parsed back it must be the
    library's image (5 digits: ES12.5) scaled by 3.25465503368d36 / nu0^2, with the optical depth next to it."""
    m = tiny(2, nlines=2)
    d = tempfile.mkdtemp(prefix="rlwd_")
    im = dict(nx=8, ny=6, size_x=100.0 * synth.AU, size_y=90.0 * synth.AU, phioff=0.0, xoff=0.0, yoff=0.0, addstar=1)
    wd.write_workdir(m, d, image=2, imager=im)
    p = wd.run_host(d, "--dump", os.path.join(d, "model.bin"))
    assert p.stdout.count("Rendering position-velocity diagram") == 2
    assert os.path.exists(os.path.join(d, "radlite.success"))
    m2 = wd.model_from_dump(m, wd.load_dump(os.path.join(d, "model.bin")))
    g = renderer_cls(0)
    g.load_model(m2)
    g.set_camera_rect(m.anginf, 8, 6, im["size_x"] / 8, im["size_y"] / 6, 0.0, 0.0, 0.0, m2.rstar, 1)
    gpu = g.render_rect(1, 2, m.nfr, m.passband)
    o = oracle_cls()
    o.load_model(m2)
    o.set_camera_rect(m.anginf, 8, 6, im["size_x"] / 8, im["size_y"] / 6, 0.0, 0.0, 0.0, m2.rstar, 1)
    ref = o.render_rect(1, 2, m.nfr, m.passband)
    for k in range(2):
        c = wd.read_posvel(os.path.join(d, f"lineposvel_moldata_{k + 1}.dat"))
        assert (c["nx"], c["ny"], c["nfr"]) == (8, 6, m.nfr) and c["lev"] == [int(m.lev_up[k]), int(m.lev_down[k])]
        conv = 3.25465503368e36 / m2.linefreq[k] ** 2
        assert np.allclose(c["temp"], conv * gpu["image"][k], rtol=1.0e-5, atol=0.0)
        assert np.allclose(c["tau"], gpu["tau"][k], rtol=1.0e-5, atol=0.0)
        assert np.allclose(c["temp"], conv * ref["image"][k], rtol=2.0e-5, atol=0.0)
        assert abs(c["spx"] - im["size_x"] / 8) < 1e-4 * c["spx"]


def test_ring_sharded_host_run_is_byte_identical(renderer_cls):
    """A cube run split over processes by camera-ring block (one per GPU: --rings lo:hi --part FILE, here three
    blocks one after the other on one GPU) and put together by --assemble writes linespectrum_<mol>.dat and
    every lineposvelcirc_<mol>_<n>.dat byte for byte as the single-process run does (telescope.F:1346-1371,
    1595-1621, 1703-1803), cmask column included."""
    import shutil
    from radlite_b200 import shard
    m = tiny(2, nlines=3)
    d1 = tempfile.mkdtemp(prefix="rlwd_")
    wd.write_workdir(m, d1)
    d2 = tempfile.mkdtemp(prefix="rlwd_")
    shutil.rmtree(d2)
    shutil.copytree(d1, d2)
    wd.run_host(d1, "--imcir")
    parts = []
    for k, (lo, hi) in enumerate(shard.split_rings(m.nrr, 3)):
        part = os.path.join(d2, f"part_{k}.bin")
        wd.run_host(d2, "--imcir", "--rings", f"{lo}:{hi}", "--part", part)
        parts.append(part)
    assert not os.path.exists(os.path.join(d2, "linespectrum_moldata.dat"))
    wd.run_host(d2, "--imcir", "--assemble", *parts)
    names = ["linespectrum_moldata.dat"] + [f"lineposvelcirc_moldata_{k}.dat" for k in (1, 2, 3)]
    for nme in names:
        a, b = open(os.path.join(d1, nme), "rb").read(), open(os.path.join(d2, nme), "rb").read()
        assert len(a) > 1000 and a == b, nme
    assert os.path.exists(os.path.join(d2, "radlite.success"))
    # spectrum-only run as well
    for f in names + ["radlite.success"]:
        os.unlink(os.path.join(d2, f))
    wd.run_host(d1)
    parts = []
    for k, (lo, hi) in enumerate(shard.split_rings(m.nrr, 2)):
        part = os.path.join(d2, f"spart_{k}.bin")
        wd.run_host(d2, "--rings", f"{lo}:{hi}", "--part", part)
        parts.append(part)
    wd.run_host(d2, "--assemble", *parts)
    assert open(os.path.join(d1, names[0]), "rb").read() == open(os.path.join(d2, names[0]), "rb").read()

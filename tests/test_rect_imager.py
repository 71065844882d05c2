"""Rectangular imager / position-velocity cube (linespectrum.inp command 2): the oracle's
restatement of setup_rays_rectang + make_freq_image_rectang against closed forms (CPU), and the CUDA path
against the oracle (GPU)."""
import numpy as np
import pytest

from helpers import clone, rel_err, static_uniform_shell, tiny
from radlite_b200 import synth

AU = synth.AU


def rect_args(m, nx=8, ny=6, half_au=12.0, **kw):
    a = dict(anginf=m.anginf, nx=nx, ny=ny, sizepix_x=2 * half_au * AU / nx, sizepix_y=2 * half_au * AU / ny,
             phioffset=0.0, xoffset=0.0, yoffset=0.0, rstar=m.rstar, addstar=0)
    a.update(kw)
    return a


def test_star_only_image_closed_form(oracle_cls):
    """No gas, no dust: a pixel sees the stellar surface intensity iff its ray hits the star (telescope.F:4194),
    the unresolved star is smeared over the four central pixels with pi R*^2 / (4 dx dy) (telescope.F:2153)."""
    m = static_uniform_shell(abund=0.0, dust_rho=0.0)
    m.starspec_cont = synth.planck(m.cont_freq_nu, 4000.0)
    o = oracle_cls()
    o.load_model(m)
    # (a) pixels much smaller than the star: the disc of the star is resolved
    n = 8
    o.set_camera_rect(**rect_args(m, nx=n, ny=n, half_au=2.0 * m.rstar / AU))
    out = o.render_rect(1, 1, m.nfr, m.passband)
    img = out["image"][0]
    istar = np.interp(np.abs(m.linefreq[0]), m.cont_freq_nu, m.starspec_cont)
    xs = (np.arange(1, n + 1) - n / 2 - 0.5) * (4.0 * m.rstar / n)
    rc = np.sqrt(xs[:, None] ** 2 + xs[None, :] ** 2)
    hit = rc * np.sqrt(1.0 + 1e-4) <= m.rstar  # tr_b = r_c sqrt(1 + 1e-4) (telescope.F:2341: dum + 1d-4)
    assert hit.any() and (~hit).any()
    assert np.allclose(img[hit], istar, rtol=1e-3) and np.all(img[~hit] == 0.0)
    assert np.all(out["tau"] == 0.0)
    # (b) pixels larger than the star with imrec_addstar: only the four central pixels light up
    o.set_camera_rect(**rect_args(m, nx=6, ny=4, half_au=12.0, addstar=1))
    img = o.render_rect(1, 1, m.nfr, m.passband)["image"][0]
    spx, spy = 24.0 * AU / 6, 24.0 * AU / 4
    srat = 3.14159265 * m.rstar ** 2 / (4.0 * spx * spy)
    centre = img[2:4, 1:3]
    assert np.allclose(centre, srat * istar, rtol=1e-3)
    img[2:4, 1:3] = 0.0
    assert np.all(img == 0.0)


def test_spherical_shell_image_depends_on_radius_only(oracle_cls):
    m = static_uniform_shell(kappa=1.0e2, dust_rho=1e-15)
    o = oracle_cls()
    o.load_model(m)
    o.set_camera_rect(**rect_args(m, nx=8, ny=8, half_au=11.0))
    out = o.render_rect(1, 1, m.nfr, m.passband)
    img, tau = out["image"][0], out["tau"][0]
    assert img.max() > 0 and tau.max() > 0
    # static sphere: the four-fold symmetry of the pixel grid
    assert np.allclose(img, img[::-1], rtol=1e-6) and np.allclose(img, img[:, ::-1], rtol=1e-6)
    assert np.allclose(img, np.transpose(img, (1, 0, 2)), rtol=1e-6)
    assert np.allclose(tau, tau[::-1], rtol=1e-6)
    # corner pixels lie outside 0.999 R_out: no ray is traced (telescope.F:2127)
    assert np.all(img[0, 0] == 0.0) and np.all(tau[0, 0] == 0.0)
    # error paths of setup_rays_rectang
    from radlite_b200._binding import RadliteError
    with pytest.raises(RadliteError):
        o.set_camera_rect(**rect_args(m, nx=7))
    with pytest.raises(RadliteError):
        o.set_camera_rect(**rect_args(m, addstar=1, xoffset=1.0))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["disk", "disk_star_offset", "isrf_rot", "shell_star", "inner_bc1"])
def test_gpu_rect_image_against_oracle(renderer_cls, oracle_cls, case):
    if case == "disk":
        m, a = tiny(2, nlines=2), None
    elif case == "disk_star_offset":
        m = tiny(2, nlines=2)
        a = rect_args(m, nx=10, ny=8, half_au=1.5, xoffset=0.2 * AU, yoffset=-0.1 * AU)
    elif case == "isrf_rot":
        m = synth.config(2, nr=30, nth=12, nphi=8, nrext=-6, nlines=2)
        m.out_itype = 3
        m.isrf_cont = 1e-12 * synth.planck(m.cont_freq_nu, 5000.0)
        a = rect_args(m, nx=6, ny=6, half_au=140.0, phioffset=0.7)
    elif case == "shell_star":
        m = static_uniform_shell(kappa=1.0e2, dust_rho=1e-15)
        m.starspec_cont = synth.planck(m.cont_freq_nu, 4000.0)
        a = rect_args(m, nx=8, ny=8, half_au=11.0, addstar=1)
    else:
        m = clone(tiny(1), in_itype=1)
        a = rect_args(m, nx=12, ny=12, half_au=0.3)
    a = a or rect_args(m, nx=10, ny=8, half_au=60.0, addstar=1)
    g, o = renderer_cls(0), oracle_cls()
    for e in (g, o):
        e.load_model(m)
        e.set_camera_rect(**a)
    out = g.render_rect(1, m.nlines, m.nfr, m.passband)
    ref = o.render_rect(1, m.nlines, m.nfr, m.passband)
    assert ref["image"].max() > 0
    assert rel_err(out["image"], ref["image"]).max() < 1e-5
    assert np.allclose(out["tau"], ref["tau"], rtol=1e-9, atol=1e-300)
    assert np.array_equal(out["maserflag"], ref["maserflag"])
    cg, co = g.counters(), o.counters()
    assert cg["R"] == co["R"] and cg["S"] == co["S"] and abs(cg["E"] - co["E"]) <= 1e-6 * co["E"]
    # the circular camera still works on the same context afterwards
    f = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"]
    fo = o.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"] if False else None
    assert np.all(np.isfinite(f))

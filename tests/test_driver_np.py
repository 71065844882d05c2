"""The numpy restatement of pyradlite's driver-side arithmetic (oracle/driver_np.py) against scipy -- the
library pyradlite itself calls (interp1d at radlite.py:1152, 3081-3164; ndimage.convolve at :3148) -- and
against closed forms.  CPU only."""
import numpy as np
import scipy.interpolate as si
import scipy.ndimage as ndi

from oracle import driver_np as D


def test_interpolation_and_convolution_match_scipy():
    rng = np.random.default_rng(7)
    x = np.sort(rng.random(40)) * 10
    y = rng.random(40)
    xn = x[0] + rng.random(200) * (x[-1] - x[0])
    assert np.array_equal(D._interp1d_linear(x, y, xn), si.interp1d(x, y, kind="linear")(xn))
    xe = rng.random(100) * 30 - 10
    ref = si.interp1d(x, y, kind="linear", bounds_error=False, fill_value="extrapolate")(xe)
    assert np.allclose(D.interp_linear_extrapolate(x, y, xe), ref, rtol=1e-13, atol=1e-13)
    a = rng.random(300)
    for obsres, vres in ((3.0, 1.5), (10.0, 1.5), (30.0, 0.5)):
        k = D.gauss_kernel(obsres, vres)
        assert np.allclose(D.convolve_reflect(a, k), ndi.convolve(a, k, mode="reflect"), rtol=1e-13)


def test_lte_populations_closed_form():
    e = np.array([0.0, 3.845, 11.535, 2143.27])
    g = np.array([1.0, 3.0, 5.0, 1.0])
    t = np.array([[20.0, 300.0], [1500.0, 4000.0]])
    pt = np.array([10.0, 100.0, 1000.0, 5000.0])
    ps = np.array([4.0, 36.0, 400.0, 3000.0])
    pop = D.lte_populations(e, g, t, pt, ps)
    q = si.interp1d(pt, ps, bounds_error=False, fill_value="extrapolate")(t)
    want = g[:, None, None] * np.exp(-e[:, None, None] * D.h0 * D.c0 / D.kB0 / t[None]) / q[None]
    assert pop.shape == (4, 2, 2) and np.allclose(pop, want, rtol=1e-14)
    assert D.lte_populations(np.array([1e5]), np.array([1.0]), np.array([5.0]), pt, ps)[0, 0] == 0.0  # flushed


def test_process_spectrum_of_one_gaussian_line():
    """One Gaussian emission line on a sloped continuum: the synthesised spectrum keeps the line flux, puts the
    peak at the line's wavelength and returns the continuum between the edges."""
    nfr = 95
    vel = np.linspace(-70.5, 70.5, nfr)[None, :]
    freq = np.array([6.4e13])
    cont = 2e-13 * (1.0 + 1e-3 * vel)
    line = 5e-13 * np.exp(-(vel / 6.0) ** 2)
    out = D.process_spectrum(vel, cont + line, freq, dist_pc=140.0, obsres=3.0, vsampling=1.5)
    mu0 = D.cinmu0 / freq[0]
    assert abs(out["wavelength"][np.argmax(out["emission"])] - mu0) < 2e-5 * mu0
    scale = 1e23 / 140.0 ** 2
    assert np.allclose(out["continuum"], 2e-13 * scale, rtol=2e-3)
    assert np.allclose(out["spectrum"], out["continuum"] + out["emission"], rtol=1e-10)
    # line flux is conserved by interpolation + normalised convolution + resampling (to the sampling error)
    dv = np.gradient(out["wavelength"]) / out["wavelength"] * D.cinkm0
    assert abs((out["emission"] * dv).sum() - (line * 1.5 * scale).sum()) < 2e-3 * (line * 1.5 * scale).sum()


def test_unique_levels_follow_the_drivers_rule():
    e_low = np.array([0.0, 3.845, 3.8451])
    e_up = np.array([2143.27, 2147.08, 2147.08])
    g_low, g_up = np.array([1.0, 3.0, 3.0]), np.array([3.0, 5.0, 5.0])
    v = np.array([0, 0, 0])
    e, g, vv, low, up = D.unique_levels(e_low, e_up, g_low, g_up, v, v + 1)
    assert len(e) == 4 and list(low) == [0, 1, 1] and list(up) == [2, 3, 3]

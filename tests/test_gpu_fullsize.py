"""BASELINE.json configs 3, 4 and 5 at full size on the B200 (configs 1 and 2 are in test_gpu_parity.py):
the CUDA path against the CPU oracle on camera-ring samples the oracle finishes in seconds, plus
size-independent properties (cube <-> spectrum consistency, batch-boundary independence, 1/d^2)."""
import numpy as np
import pytest

from conftest import same_bits
from helpers import rel_err
from radlite_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "z"], autouse=True)
def _kernel(request):
    """Each full-size test once with the library's kernel choice and once with ztile_kernel forced
    (tile_kernel at these sizes is covered by the configs it serves: the single-line ones)."""
    from radlite_b200 import api
    old = api.DEFAULT_KERNEL
    api.DEFAULT_KERNEL = request.param
    yield
    api.DEFAULT_KERNEL = old
TOL_PIX = 1e-5


def ring_sample_check(g_img, o, m, iline, rings, nfr):
    """Pixel rows of the sampled rings: oracle traces only those (orc_set_ring_sample)."""
    worst = 0.0
    for ir in rings:
        o.set_ring_sample(ir, ir, 1)
        ref = o.render(iline, 1, nfr, m.passband, synth.PARSEC, want_image=True)["image"][0]
        worst = max(worst, rel_err(g_img[ir], ref[ir]).max())
        assert rel_err(g_img[0], ref[0]).max() < TOL_PIX  # centre ray is always traced
    o.set_ring_sample(0, 0, 1)
    return worst


def flux_from_cube(g, img, dist):
    """telescope.F:1388-1433 in numpy (order of the phi sum differs: compare to 1e-12, not bitwise)."""
    _, ri = g.rings()
    nrr = img.shape[0] - 1
    surf = 3.14159265359 * np.diff(ri[1:] ** 2)
    f = 3.14159265359 * ri[1] ** 2 * img[0, 0]
    f = f + (img[1:].mean(axis=1) * surf[:nrr, None]).sum(axis=0)
    return f / dist**2


def test_cfg2_100_lines_as_benchmarked(renderer_cls, oracle_cls):
    """configs[1] exactly as bench.py renders it -- all 100 lines in ONE batch (7 tiles of 16 lines per ray
    plus the continuum-only tiles), spectrum mode for the flux, cube + mask for the pixel checks -- against
    the oracle on three camera rings: every pixel and mask value of those rings for lines 1, 16, 17 (a tile
    boundary), 50 and 100, and the ray-channel / segment / element counts of the rings for all 100 lines."""
    m = synth.config(2)
    assert (m.nlines, m.nray, m.nfr) == (100, 40351, 94)
    nl, nfr = m.nlines, m.nfr
    g = renderer_cls(0)
    g.load_model(m)
    flux = g.render(1, nl, nfr, m.passband, synth.PARSEC)["flux"]  # what bench.py's e2e leg returns
    out = g.render(1, nl, nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    assert np.array_equal(out["flux"], flux)
    o = oracle_cls()
    o.load_model(m)
    lines = [1, 16, 17, 50, 100]
    for ir in (75, 150, 262):
        o.reset_counters()
        o.set_ring_sample(ir, ir, 1)
        ref = o.render(1, nl, nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
        co = o.counters()
        for il in lines:
            assert rel_err(out["image"][il - 1, ir], ref["image"][il - 1, ir]).max() < TOL_PIX, (ir, il)
            assert rel_err(out["image"][il - 1, 0], ref["image"][il - 1, 0]).max() < TOL_PIX, (ir, il)
        # imcir_cmask accumulates over the lines of a run in both (telescope.F:548,575 never clear it)
        assert np.array_equal(out["cmask"][:, ir], ref["cmask"][:, ir]), ir
        # the reference's work for this ring (+ the centre ray, which the oracle always traces)
        g.reset_counters()
        g.render_rings(1, nl, nfr, m.passband, synth.PARSEC, ir, ir)
        g.render_rings(1, nl, nfr, m.passband, synth.PARSEC, 0, 0)
        cg = g.counters()
        assert cg["R"] == co["R"] and cg["S"] == co["S"], (ir, cg, co)
        assert abs(cg["E"] - co["E"]) <= 1e-6 * co["E"], (ir, cg, co)
        del ref
    o.set_ring_sample(0, 0, 1)


def test_cfg3_13co_cube_full_size(renderer_cls, oracle_cls):
    """configs[2]: 13CO image cube, 400x160 grid, 70 351 circular-polar pixels x 200 channels, incl 45."""
    m = synth.config(3)
    assert (len(m.r), len(m.theta), m.nray) == (400, 160, 70351) and 199 <= m.nfr <= 200
    g = renderer_cls(0)
    g.load_model(m)
    out = g.render(1, 1, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    img = out["image"][0]
    assert img.shape == (m.nrr + 1, m.nphi, m.nfr) and np.all(np.isfinite(img)) and np.all(img >= 0)
    # spectrum == ring-area weighted sum of the cube; spectrum-only render is bit-identical
    assert np.allclose(out["flux"][0], flux_from_cube(g, img, synth.PARSEC), rtol=1e-12)
    assert np.array_equal(g.render(1, 1, m.nfr, m.passband, synth.PARSEC)["flux"], out["flux"])
    # NONREDUNDANT: channel 1 is always traced; every unmasked channel of a pixel carries one and the
    # same continuum value (telescope.F:582-612); the centre beam never sets its mask
    cm = out["cmask"][0].astype(bool)
    assert cm[1:, :, 0].all() and not cm[0].any()
    lo = np.where(~cm[1:], img[1:], np.inf).min(axis=2)
    hi = np.where(~cm[1:], img[1:], -np.inf).max(axis=2)
    has = (~cm[1:]).any(axis=2)
    assert has.any() and np.array_equal(lo[has], hi[has])
    # mirror symmetry of the Keplerian disk: the line profile is symmetric about line centre
    f = out["flux"][0]
    line = f - 0.5 * (f[0] + f[-1])
    assert np.abs(line - line[::-1]).max() < 2e-3 * np.abs(line).max()
    o = oracle_cls()
    o.load_model(m)
    worst = ring_sample_check(img, o, m, 1, [3, 66, 75, 240, 469], m.nfr)
    assert worst < TOL_PIX, worst


def test_cfg4_nlte_500_lines(renderer_cls, oracle_cls):
    """configs[3]: ~500 12CO lines of the 4.6-5.0 um band, two-temperature (NLTE stand-in) populations
    with freeze-out, 200x80 grid: several line batches, spectra independent of the batching."""
    m = synth.config(4)
    assert 400 < m.nlines <= 500 and m.nray == 40351
    nl = m.nlines
    g = renderer_cls(0)
    g.load_model(m)
    f = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"]
    assert f.shape == (nl, m.nfr) and np.all(np.isfinite(f)) and np.all(f > 0)
    for il in (1, 128, 129, 256, 257, nl):  # batch boundaries (128 lines per batch)
        one = g.render(il, 1, m.nfr, m.passband, synth.PARSEC)["flux"][0]
        assert same_bits(one, f[il - 1]), il
    assert np.allclose(g.render(1, nl, m.nfr, m.passband, 3.0 * synth.PARSEC)["flux"] * 9.0, f, rtol=1e-15)
    o = oracle_cls()
    o.load_model(m)
    for il in (7, 333):
        img = g.render(il, 1, m.nfr, m.passband, synth.PARSEC, want_image=True)["image"][0]
        assert ring_sample_check(img, o, m, il, [2, 70, 150, 269], m.nfr) < TOL_PIX


def test_cfg5_large_grid_high_refinement(renderer_cls, oracle_cls):
    """configs[4] geometry: 1000x400 grid (160 351 rays, ~3e8 ray nodes), line widths x0.2 (narrow
    profiles: few channels per ray survive NONREDUNDANT, sub-grid candidates 6q > 1 where the fine grid
    still under-resolves the Doppler shift); 6 lines of the 2000 (the full molecule's level-population
    table alone is 4 GB, so the test carries a cut-down ladder)."""
    mol = synth.rovib_molecule(2, 40, nlines=6)
    m = synth.make_model("cfg5_geometry_1000x400", 1000, 400, mol, width_scale=0.2)
    assert m.nray == 160351
    g = renderer_cls(0)
    g.load_model(m)
    out = g.render(1, 6, m.nfr, m.passband, synth.PARSEC)
    f = out["flux"]
    assert np.all(np.isfinite(f)) and np.all(f > 0)
    c = g.counters()
    assert c["E"] >= c["S"] > 0  # sub-grid steps only ever add work on top of the plain segments
    assert g.total_nodes() > 2.0e8
    one = g.render(4, 1, m.nfr, m.passband, synth.PARSEC, want_image=True)
    assert same_bits(one["flux"][0], f[3])
    o = oracle_cls()
    o.load_model(m)
    assert ring_sample_check(one["image"][0], o, m, 4, [5, 400, 1069], m.nfr) < TOL_PIX
    # sub-grid off: the same answer to within the refinement's effect, and again equal to the oracle
    g.set_options(0, 1, m.levthres, m.aksmax)
    o.set_options(0, 1, m.levthres, m.aksmax)
    off = g.render(4, 1, m.nfr, m.passband, synth.PARSEC, want_image=True)
    assert rel_err(off["flux"][0], f[3]).max() < 0.2
    assert ring_sample_check(off["image"][0], o, m, 4, [400], m.nfr) < TOL_PIX

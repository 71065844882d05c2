"""Parity of the CUDA path (through the C ABI of include/radlite_b200.h) against the CPU oracle and
the committed fixtures.  Tolerances are BASELINE.json's: relative 1e-5 per channel and 1e-6 on the
integrated line flux; per-pixel intensities are held to 1e-5 as well.  Work counters (which
ray-channel integrations / element integrations the reference would perform) must agree exactly."""
import os

import numpy as np
import pytest

from conftest import same_bits
from helpers import clone, golden_cases, rel_err, static_uniform_shell, tiny
from radlite_b200 import synth
from radlite_b200._binding import RadliteError

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("integrate_kernel")]

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_CH, TOL_INT, TOL_PIX = 1e-5, 1e-6, 1e-5


def both(renderer_cls, oracle_cls, m, image=True):
    g = renderer_cls(0)
    g.load_model(m)
    out = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=image, want_mask=image)
    o = oracle_cls()
    o.load_model(m)
    ref = o.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=image, want_mask=image)
    return g, out, o, ref


def check(g, out, o, ref, image=True, counters=True):
    assert rel_err(out["flux"], ref["flux"]).max() < TOL_CH
    line = lambda f: np.abs((f - f[:, :1]).sum(axis=1))  # noqa: E731  integrated line flux above the first channel
    tot_ref = np.abs(ref["flux"].sum(axis=1))
    assert (np.abs(out["flux"].sum(axis=1) - ref["flux"].sum(axis=1)) / tot_ref).max() < TOL_INT
    assert np.allclose(line(out["flux"]), line(ref["flux"]), rtol=1e-5, atol=1e-9 * tot_ref.max())
    assert np.array_equal(out["velo"], ref["velo"])
    assert rel_err(out["tau_center"], ref["tau_center"]).max() < 1e-9 or np.all(ref["tau_center"] == 0)
    assert np.array_equal(out["maserflag"], ref["maserflag"])
    if image:
        assert rel_err(out["image"], ref["image"]).max() < TOL_PIX
        assert np.array_equal(out["cmask"], ref["cmask"])
    if counters:
        cg, co = g.counters(), o.counters()
        assert cg["R"] == co["R"] and cg["S"] == co["S"]
        # E: a sub-grid point on a segment end may flip (1e-6).  The opaque-wall start (on by default) does not
        # change the counter: wallcount_kernel adds the sub-grid steps of the segments it skips
        assert abs(cg["E"] - co["E"]) <= 1e-6 * co["E"] and cg["E"] >= cg["S"]
        assert g.executed_elements() <= cg["E"]


@pytest.mark.parametrize("name,model", list(golden_cases()), ids=[n for n, _ in golden_cases()])
def test_against_oracle_and_golden(renderer_cls, oracle_cls, name, model):
    g, out, o, ref = both(renderer_cls, oracle_cls, model)
    check(g, out, o, ref)
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    assert rel_err(out["flux"], gold["flux"]).max() < TOL_CH
    assert rel_err(out["image"][:, ::7], gold["image_ring"]).max() < TOL_PIX
    assert np.array_equal(out["cmask"][:, ::7].astype(np.int8), gold["cmask_ring"])
    cg = g.counters()
    assert cg["R"] == gold["counters"][0] and cg["S"] == gold["counters"][2]


@pytest.mark.parametrize("change", [
    dict(subgrid=0), dict(nonredundant=0), dict(out_itype=2), dict(out_itype=3), dict(in_itype=1),
    dict(dbdr=3), dict(incl_deg=60.0), dict(incl_deg=2.0), dict(levthres=1e3), dict(nrext=-1),
    dict(vmax_kms=20.0, dv_kms=0.3),
], ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()))
def test_option_matrix(renderer_cls, oracle_cls, change):
    m = tiny(2, nlines=2)
    if change.get("out_itype") == 3:
        m = synth.config(2, nr=30, nth=12, nphi=8, nrext=-6, nlines=2)
        m.out_itype = 3
        m.isrf_cont = 1e-12 * synth.planck(m.cont_freq_nu, 5000.0)
    else:
        m = clone(m, **change)
    check(*both(renderer_cls, oracle_cls, m))


def test_geometry_nodes_match_oracle(renderer_cls, oracle_cls):
    """Device-built node lists against make_trajectory_c + get_line_dust_values of the oracle."""
    m = tiny()
    g = renderer_cls(0)
    g.load_model(m)
    o = oracle_cls()
    o.load_model(m)
    _, _, nray = o.camera_dims()
    for iray in list(range(1, 20)) + list(range(20, nray, 23)) + [nray]:
        nd = g.ray_nodes(iray)
        t = o.trajectory(iray)
        v = o.node_values(iray, 1)
        assert len(nd["ds"]) == len(t["s"])
        assert np.array_equal(nd["flags"] & 3, t["icross"])
        ds = np.diff(t["s"])
        keep = (nd["flags"][1:] & 4) == 0  # vacuum segments carry ds = 0
        assert np.allclose(nd["ds"][1:][keep], ds[keep], rtol=1e-10, atol=1e-13 * np.abs(t["s"]).max())
        assert np.allclose(nd["dvmu"], v["dvmu"], rtol=1e-10, atol=1e-22)
        assert np.allclose(nd["lw"], v["lw"], rtol=1e-12)


def test_zero_continuum_path(renderer_cls, oracle_cls):
    """No dust, no star, black sky: the ray continuum is exactly 0, so the reference re-integrates
    every skipped channel (imcir_cont.ne.0 test, telescope.F:583)."""
    m = static_uniform_shell(rho=1e-20, abund=1e-8)
    m.vmax_kms, m.dv_kms = 6.0, 0.5
    g, out, o, ref = both(renderer_cls, oracle_cls, m)
    check(g, out, o, ref)
    assert o.counters()["R"] > m.nray  # more than one channel per ray was integrated


def test_thick_dust_and_star_only(renderer_cls, oracle_cls):
    check(*both(renderer_cls, oracle_cls, static_uniform_shell(kappa=1e3, dust_rho=1e-10, abund=0.0)))
    m = tiny()
    m.abund[:] = 0.0
    m.dust_rho[:] = 0.0
    check(*both(renderer_cls, oracle_cls, m))


def test_maser_flag(renderer_cls, oracle_cls):
    m = tiny()
    # invert the populations of the line: alpha_line < 0 (telescope.F:4295)
    up, dn = m.lev_up[0] - 1, m.lev_down[0] - 1
    m.popul[..., up] = 0.5
    m.popul[..., dn] = 1e-6
    m.abund[:] = 1e-3
    g, out, o, ref = both(renderer_cls, oracle_cls, m)
    assert ref["maserflag"][0] == 1 and out["maserflag"][0] == 1
    # under inversion qdr_src_2 takes its a = b = dtau/2, xp = 1 - dtau, Q = theomax branch (transfer.F:1522-1524,
    # 1545) with dtau < 0: the amplification compounds along the ray, and so does the 1e-13 difference between
    # the kernels' and the reference's exp / division -- still inside the per-channel tolerance
    assert rel_err(out["flux"], ref["flux"]).max() < TOL_CH
    assert rel_err(out["image"], ref["image"]).max() < TOL_PIX


def test_precomputed_line_dust_and_line_subsets(renderer_cls, oracle_cls):
    """rl_set_line_dust (host already ran global_prepare_line_dust) and rendering a sub-range."""
    m = tiny(2, nlines=5)
    g, out, o, ref = both(renderer_cls, oracle_cls, m, image=False)
    sub = g.render(2, 3, m.nfr, m.passband, synth.PARSEC)
    assert same_bits(sub["flux"], out["flux"][1:4])
    # continuum-only arrays of the right shape: src = alp * B
    nl, nr, nth = m.nlines, len(m.r), len(m.theta)
    alp = np.repeat((m.dust_rho[..., 0] * 50.0)[None], nl, axis=0)
    src = alp * 1e-9
    g.set_line_dust(src, alp)
    o.set_line_dust(src, alp)
    a = g.render(1, nl, m.nfr, m.passband, synth.PARSEC)
    b = o.render(1, nl, m.nfr, m.passband, synth.PARSEC)
    assert rel_err(a["flux"], b["flux"]).max() < TOL_CH


def test_repeatable_bitwise_and_mask_accumulates(renderer_cls):
    m = tiny(2, nlines=3)
    g = renderer_cls(0)
    g.load_model(m)
    a = g.render(1, 3, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    g.invalidate_geometry()
    b = g.render(1, 3, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    assert np.array_equal(a["flux"], b["flux"]) and np.array_equal(a["image"], b["image"])
    # imcir_cmask is never cleared between lines: monotone over the lines of a run
    assert np.all(a["cmask"][1] >= a["cmask"][0]) and np.all(a["cmask"][2] >= a["cmask"][1])


def test_spectrum_only_equals_cube_mode(renderer_cls):
    """Spectrum-only renders never materialise the continuum copies of the skipped channels (the ring
    sum synthesises them); the flux must be bit-identical to the cube-mode render, including rows
    whose continuum is exactly zero (completed by the fallback) and NONREDUNDANT off."""
    for m in (tiny(2, nlines=3), static_uniform_shell(rho=1e-20, abund=1e-8), clone(tiny(2, nlines=2), nonredundant=0)):
        g = renderer_cls(0)
        g.load_model(m)
        a = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
        b = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)
        assert np.array_equal(a["flux"], b["flux"])
        assert np.array_equal(a["maserflag"], b["maserflag"])


def test_ring_block_sharding_is_bit_identical(renderer_cls):
    """rl_render_rings over disjoint ring blocks + rl_flux_from_rings == rl_render, bit for bit
    (the multi-GPU path of single-line configs, run here as consecutive blocks on one GPU)."""
    from radlite_b200 import shard
    for m in (tiny(1), tiny(2, nlines=3), clone(tiny(1), nonredundant=0)):
        g = renderer_cls(0)
        g.load_model(m)
        full = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True)
        nrr, nphi, _ = g.camera_dims()
        for world in (1, 2, 3):
            rs = np.zeros((m.nlines, nrr + 1, m.nfr))
            cube = np.zeros((m.nlines, nrr + 1, nphi, m.nfr))
            for lo, hi in shard.split_rings(nrr, world):
                part = g.render_rings(1, m.nlines, m.nfr, m.passband, synth.PARSEC, lo, hi, image=cube)
                assert not np.delete(part, np.s_[lo:hi + 1], axis=1).any()
                rs = rs + part
            assert np.array_equal(g.flux_from_rings(rs, synth.PARSEC), full["flux"])
            assert np.array_equal(cube, full["image"])
        # spectrum-only (sparse) ring blocks as well
        rs = sum(g.render_rings(1, m.nlines, m.nfr, m.passband, synth.PARSEC, lo, hi)
                 for lo, hi in shard.split_rings(nrr, 2))
        assert np.array_equal(g.flux_from_rings(rs, synth.PARSEC), full["flux"])
        with pytest.raises(RadliteError):
            g.render_rings(1, 1, m.nfr, m.passband, synth.PARSEC, 5, 4)


def test_error_codes_match_reference_stops(renderer_cls):
    m = tiny()
    g = renderer_cls(0)
    g.load_model(m)
    for args, code in (((1, 1, 1), 13), ((2, 1, m.nfr), 13)):
        with pytest.raises(RadliteError) as e:
            g.render(args[0], args[1], args[2], m.passband, synth.PARSEC)
        assert e.value.code == code
    with pytest.raises(RadliteError) as e:
        g.set_camera(m.anginf, m.nphi, m.nrext, m.dbdr, 2 * m.r[0])
    assert e.value.code == 91991
    g2 = renderer_cls(0)
    g2.load_model(clone(m, in_itype=0))
    with pytest.raises(RadliteError) as e:
        g2.render(1, 1, m.nfr, m.passband, synth.PARSEC)
    assert e.value.code == 13
    g3 = renderer_cls(0)
    g3.load_model(clone(m, out_itype=1))
    with pytest.raises(RadliteError) as e:
        g3.render(1, 1, m.nfr, m.passband, synth.PARSEC)
    assert e.value.code == 13
    g4 = renderer_cls(0)
    with pytest.raises(RadliteError) as e:
        g4.render(1, 1, 10, 1.0, 1.0)
    assert e.value.code == 13  # rays_ready.ne.321 (telescope.F:366)


def test_cfg1_full_size_against_oracle(renderer_cls, oracle_cls):
    """BASELINE configs[0] at full size (100x40 grid, 25 351 rays, 94 channels): ~5 s of oracle."""
    m = synth.config(1)
    g, out, o, ref = both(renderer_cls, oracle_cls, m, image=True)
    check(g, out, o, ref)


def test_cfg2_full_size_properties(renderer_cls, oracle_cls):
    """BASELINE configs[1] geometry at full size (200x80, 40 351 rays); 3 lines against the oracle
    on a ring sample is too slow for CI, so: two lines against the oracle's spectrum plus
    size-independent properties on 8 lines."""
    m = synth.config(2, nlines=8)
    g = renderer_cls(0)
    g.load_model(m)
    out = g.render(1, 8, m.nfr, m.passband, synth.PARSEC)
    f = out["flux"]
    assert np.all(np.isfinite(f)) and np.all(f > 0)
    # Keplerian, mirror-symmetric disk: the line profile is symmetric about line centre
    line = f - 0.5 * (f[:, :1] + f[:, -1:])
    assert (np.abs(line - line[:, ::-1]).max(axis=1) < 2e-3 * np.abs(line).max(axis=1)).all()
    # flux scales as 1/d^2 exactly (telescope.F:1433)
    out2 = g.render(1, 8, m.nfr, m.passband, 2.0 * synth.PARSEC)
    assert np.array_equal(out2["flux"] * 4.0, f)
    # opaque-wall start: integrating every segment gives the same spectra, bit for bit
    g.reset_counters()
    g.render(1, 8, m.nfr, m.passband, synth.PARSEC)
    frac_on = g.executed_elements() / g.counters()["E"]
    g.set_wall_tau(0.0)
    assert np.array_equal(g.render(1, 8, m.nfr, m.passband, synth.PARSEC)["flux"], f)
    g.set_wall_tau(64.0)
    assert 0.3 < frac_on < 0.9, frac_on  # about 40 % of this disk's element integrations lie behind walls
    # rendering lines one by one gives bit-identical spectra (lines are independent)
    one = g.render(5, 1, m.nfr, m.passband, synth.PARSEC)
    assert same_bits(one["flux"][0], f[4])
    # NONREDUNDANT only trims far wings: spectra agree to the exp(-4) wing truncation level
    g.set_options(1, 0, m.levthres, m.aksmax)
    full = g.render(1, 2, m.nfr, m.passband, synth.PARSEC)
    assert rel_err(full["flux"], f[:2]).max() < 5e-3
    # and one line against the oracle (about 15 s on one core)
    o = oracle_cls()
    o.load_model(m)
    ref = o.render(3, 1, m.nfr, m.passband, synth.PARSEC)
    assert rel_err(f[2:3], ref["flux"]).max() < TOL_CH
    assert abs(f[2].sum() - ref["flux"].sum()) / ref["flux"].sum() < TOL_INT


def test_opaque_wall_start_is_exact(renderer_cls, oracle_cls, integrate_kernel):
    """rl_set_wall_tau: segments whose contribution is provably below e^-64 of what the dust in front of them
    emits are not integrated; the image must not change in any bit, the reference's work counters R, E, S stay
    what they are and only the executed element count drops."""
    m = synth.config(2, nr=60, nth=24, nphi=16, nrext=-8, nlines=8)
    g = renderer_cls(0)
    g.load_model(m)
    g.reset_counters()
    on = g.render(1, 8, m.nfr, m.passband, synth.PARSEC, want_image=True)
    c_on, ex_on = g.counters(), g.executed_elements()
    g.set_wall_tau(0.0)
    g.reset_counters()
    off = g.render(1, 8, m.nfr, m.passband, synth.PARSEC, want_image=True)
    c_off, ex_off = g.counters(), g.executed_elements()
    o = oracle_cls()
    o.load_model(m)
    o.render(1, 8, m.nfr, m.passband, synth.PARSEC)
    co = o.counters()
    assert c_off["R"] == co["R"] and c_off["S"] == co["S"] and abs(c_off["E"] - co["E"]) <= 1e-6 * co["E"]
    assert np.array_equal(on["flux"], off["flux"])
    assert np.array_equal(on["image"], off["image"])
    assert c_on["R"] == c_off["R"] and c_on["S"] == c_off["S"] and c_on["E"] == c_off["E"]
    assert ex_off == c_off["E"]
    assert ex_on < 0.9 * ex_off, (ex_on, ex_off)  # this disk's midplane is opaque at 4.7 um
    g.set_wall_tau(1.0)  # an absurdly thin "wall": now the image must change (the cut really is applied)
    thin = g.render(1, 8, m.nfr, m.passband, synth.PARSEC, want_image=True)
    assert not np.array_equal(thin["image"], off["image"])
    assert rel_err(thin["flux"], off["flux"]).max() < 1.0


def test_chan_and_ztile_kernels_give_the_same_bits(renderer_cls):
    """chan_kernel (channels of one line across the lanes) and ztile_kernel (lines across the lanes) perform the
    same operations per (ray, line, channel): image, cube mask and spectrum agree bit for bit."""
    from radlite_b200 import api
    old = api.DEFAULT_KERNEL
    res = {}
    try:
        for k in ("z", "chan"):
            api.DEFAULT_KERNEL = k
            m = synth.config(2, nr=48, nth=20, nphi=16, nrext=-10, nlines=6)
            g = renderer_cls(0)
            g.load_model(m)
            res[k] = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
            res[k + "_cnt"] = g.counters()
            g.close()
    finally:
        api.DEFAULT_KERNEL = old
    assert np.array_equal(res["z"]["flux"], res["chan"]["flux"])
    assert np.array_equal(res["z"]["image"], res["chan"]["image"])
    assert np.array_equal(res["z"]["cmask"], res["chan"]["cmask"])
    assert res["z_cnt"] == res["chan_cnt"]

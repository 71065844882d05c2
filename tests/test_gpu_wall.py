"""Adversarial cases for the opaque-wall start (rl_set_wall_tau): models built so that a cut
at a fixed dust optical depth would be WRONG -- an opaque foreground that emits nothing, or next to nothing,
in front of a hot interior -- and models where the far side starts from a non-zero boundary intensity or
hides inverted populations.  The reference makes no cut (telescope.F:4079-4300); the library's criterion
compares what it drops with a lower bound of what the dust in front emits, so every pixel must still agree
with the oracle to the per-channel tolerance, however small the emerging intensity is."""
import numpy as np
import pytest

from helpers import rel_err, static_uniform_shell
from radlite_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("integrate_kernel")]

AU = synth.AU


def layered_shell(t_out, t_in=1500.0, tau_out=220.0, nlines=3, **kw):
    """Static shell 1..10 AU: hot inside 4 AU, an outer layer at ``t_out`` K whose radial dust optical depth
    is ``tau_out`` at 4.7 um; CO-like gas everywhere."""
    m = static_uniform_shell(nr=30, nth=10, nphi=6, nrext=-4, rho=1e-14, abund=1e-5, tgas=300.0, lw=1.0,
                             kappa=1.0e3, dust_rho=1.0, **kw)
    mol = synth.rovib_molecule(1, 12, nlines=nlines)
    for key in ("ener_cm", "gdeg", "lev_v", "lev_j", "lev_up", "lev_down", "aud", "linefreq"):
        setattr(m, key, mol[key])
    cold = m.r > 4.0 * AU
    m.dust_rho[:] = tau_out / (1.0e3 * 6.0 * AU)  # kappa rho (10 AU - 4 AU) = tau_out
    m.dust_temp[:] = t_in
    m.dust_temp[cold] = t_out
    m.tgas[:] = t_in
    m.tgas[cold] = max(t_out, 15.0)
    m.popul = synth.level_populations(mol, m.tgas)
    m.vmax_kms, m.dv_kms = 9.0, 1.5
    return m


def compare(renderer_cls, oracle_cls, m):
    g = renderer_cls(0)
    g.load_model(m)
    out = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    o = oracle_cls()
    o.load_model(m)
    ref = o.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
    # per pixel, relative, with no absolute floor above the smallest normal double
    assert rel_err(out["image"], ref["image"]).max() < 1e-5, rel_err(out["image"], ref["image"]).max()
    assert np.array_equal(out["cmask"], ref["cmask"])
    assert rel_err(out["flux"], ref["flux"]).max() < 1e-5
    cg, co = g.counters(), o.counters()
    assert cg["R"] == co["R"] and cg["S"] == co["S"] and abs(cg["E"] - co["E"]) <= 1e-6 * co["E"]
    return g, out, ref


@pytest.mark.parametrize("t_out", [0.0, 12.0, 20.0, 60.0])
def test_cold_opaque_foreground(renderer_cls, oracle_cls, t_out):
    """T = 0: the foreground emits nothing (bplanck returns 0, setup.F:937) and the pixel IS the attenuated
    interior, e^-220 of it.  T = 12 .. 20 K: Wien-suppressed emission (h nu / k T = 150 .. 250) of the same
    order as the attenuated interior.  T = 60 K: the foreground dominates and the far side may be dropped."""
    g, out, ref = compare(renderer_cls, oracle_cls, layered_shell(t_out))
    assert ref["image"][:, 1:].max() > 0.0
    frac = g.executed_elements() / g.counters()["E"]
    if t_out == 0.0:
        assert frac == 1.0  # no emission in front at all: nothing may be skipped
    elif t_out >= 60.0:
        assert frac < 1.0   # here the cut is provably harmless, and it is taken


@pytest.mark.parametrize("out_itype", [2, 3])
def test_nonzero_far_side_start(renderer_cls, oracle_cls, out_itype):
    """Outer boundary types 2 (CMB) and 3 (interstellar field): the far end of every ray starts from a
    non-zero intensity (telescope.F:3989-4028), which counts towards what a cut would drop."""
    m = layered_shell(40.0, tau_out=400.0)
    m.out_itype = out_itype
    if out_itype == 3:
        m.isrf_cont = 1.0e-2 * synth.planck(m.cont_freq_nu, 2.0e4)  # far brighter than anything in the model
    compare(renderer_cls, oracle_cls, m)


def test_inverted_populations_behind_the_wall(renderer_cls, oracle_cls):
    """A maser hidden behind the opaque layer: the batch is never shortened (the amplification is unbounded)."""
    m = layered_shell(80.0, tau_out=400.0, )
    hot = m.r <= 4.0 * AU
    up, dn = m.lev_up[0] - 1, m.lev_down[0] - 1
    m.popul[hot, :, up] = 0.3
    m.popul[hot, :, dn] = 1e-8
    g, out, ref = compare(renderer_cls, oracle_cls, m)
    assert g.executed_elements() == g.counters()["E"]
    assert np.array_equal(out["maserflag"], ref["maserflag"])

"""Small model builders shared by the oracle KATs and the GPU parity tests."""
import copy

import numpy as np

from radlite_b200 import synth


def tiny(n=1, **kw):
    args = dict(nr=30, nth=12, nphi=8, nrext=-6)
    if n == 2:
        args["nlines"] = 6
    args.update(kw)
    return synth.config(n, **args)


def clone(m, **changes):
    m2 = copy.deepcopy(m)
    for k, v in changes.items():
        setattr(m2, k, v)
    return m2


def static_uniform_shell(nr=24, nth=10, nphi=6, nrext=-4, rho=1e-16, abund=1e-6, tgas=300.0,
                         lw=1.0, kappa=0.0, dust_rho=1e-30):
    """Static, isothermal, constant-density shell between R_1 and R_nr (for analytic KATs)."""
    mol = synth.rovib_molecule(1, 12)
    k = [0]
    for key in ("lev_up", "lev_down", "aud", "linefreq"):
        mol[key] = mol[key][k]
    m = synth.make_model("shell", nr, nth, mol, nphi=nphi, nrext=nrext, rin_au=1.0, rout_au=10.0)
    m.rho[:] = rho
    m.abund[:] = abund
    m.vel[:] = 0.0
    m.linewidth[:] = lw
    m.tgas[:] = tgas
    m.popul = synth.level_populations(mol, m.tgas)
    m.dust_rho[:] = dust_rho
    m.dust_temp[:] = tgas
    m.kappa_abs[:] = kappa
    m.starspec_cont = np.zeros_like(m.starspec_cont)
    return m


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def golden_cases():
    """(name, model) pairs behind tests/golden/*.npz (see tests/golden/make_golden.py)."""
    yield "tiny_cfg1", tiny(1)
    yield "tiny_cfg2_6lines", tiny(2)
    yield "tiny_cfg1_dbdr2_cmb", clone(tiny(1), dbdr=2, out_itype=2)
    m = tiny(2, nlines=3)
    m.linewidth *= 0.2
    yield "tiny_narrow_subgrid", m
    yield "tiny_cfg4_nlte", synth.config(4, nr=24, nth=10, nphi=6, nrext=-4, nlines=5)

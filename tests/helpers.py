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


class OracleEngine:
    """The CPU oracle behind the call shapes the sharding layer expects from an engine
    (``render``, ``camera_dims``, ``render_rings``, ``flux_from_rings``): lets the world_size-2 gloo
    tests exercise radlite_b200.shard without a GPU.  Ring sums follow telescope.F:1388-1433."""

    def __init__(self, model):
        from oracle.oracle_py import Oracle
        self.o = Oracle()
        self.o.load_model(model)
        self.m = model

    def camera_dims(self):
        return self.o.camera_dims()

    def render(self, *a, **kw):
        return self.o.render(*a, **kw)

    def render_rings(self, iline0, nl, nfr, vmax_kms, dist_cm, ring_lo, ring_hi, image=None):
        nrr, nphi, _ = self.o.camera_dims()
        _, ri = self.o.rings()
        # the oracle always traces the centre ray; rings outside [max(lo,1), hi] are left zero
        if ring_hi >= 1:
            self.o.set_ring_sample(max(ring_lo, 1), ring_hi, 1)
        else:
            self.o.set_ring_sample(nrr + 1, nrr + 1, 1)
        img = self.o.render(iline0, nl, nfr, vmax_kms, dist_cm, want_image=True)["image"]
        self.o.set_ring_sample(0, 0, 1)
        rs = np.zeros((nl, nrr + 1, nfr))
        if ring_lo == 0:
            rs[:, 0] = (3.14159265359 * (ri[1] * ri[1])) * img[:, 0, 0]
        for ir in range(max(ring_lo, 1), ring_hi + 1):
            surf = 3.14159265359 * (ri[ir + 1] * ri[ir + 1] - ri[ir] * ri[ir])
            d = np.zeros((nl, nfr))
            for ip in range(nphi):
                d = d + img[:, ir, ip]
            rs[:, ir] = d / (1.0 * nphi) * surf
        if image is not None:
            image[:, ring_lo:ring_hi + 1] = img[:, ring_lo:ring_hi + 1]
        return rs

    def flux_from_rings(self, ringsum, dist_cm):
        s = np.zeros((ringsum.shape[0], ringsum.shape[2]))
        for ir in range(ringsum.shape[1]):
            s = s + ringsum[:, ir]
        return s / (dist_cm * dist_cm)


def tutorial_model():
    """The reference's tutorial fixture (tests/golden/tutorial_120x100.npz, built by
    tests/golden/make_tutorial_fixture.py from DOCS/DOCS_VERSION_1-3/files_for_tutorials/) as a synth.Model: the
    RADMC grid, dust and star as read, the gas fields and LTE populations as pyradlite derives them
    (oracle/driver_np.py: gas = 12800 x dust, T_gas = T_dust, constant abundance, Keplerian rotation, alpha-
    turbulence + thermal width), the 26 12CO lines of 4.6-4.7 um, camera and passband of input_radlite.json."""
    import os

    from oracle import driver_np as D
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tutorial_120x100.npz"))
    r, theta = z["r"], z["theta"]
    nr, nth = len(r), len(theta)
    tdust = z["dust_temp"][:, :, 0, 0]
    tgas = D.gas_temperature(tdust)
    rho = D.gas_density(z["dust_rho"][:, :, 0], float(z["gastodust"]))
    abund = D.abundance(tdust, float(z["min_abun"]), float(z["max_abun"]))
    turb = D.turbulence(tdust, float(z["alpha"]), float(z["gamma"]), float(z["mu"]), float(z["molweight"]))
    vel = np.zeros((nr, nth, 3))
    vel[..., 2] = D.velocity_phi(r, nth, float(z["mstar"])).T
    popul = np.ascontiguousarray(np.moveaxis(D.lte_populations(z["ener_cm"], z["gdeg"], tgas, z["psum_temp"], z["psum"]), 0, -1))
    eerg = 1.986468498e-16 * z["ener_cm"]
    linefreq = 1.509160e26 * (eerg[z["lev_up"] - 1] - eerg[z["lev_down"] - 1])  # line.F:1903, 1981
    m = synth.Model(
        name="tutorial_120x100", r=r, theta=theta, rho=np.ascontiguousarray(rho), abund=abund, vel=vel,
        linewidth=np.ascontiguousarray(turb / 1.0E5), tgas=tgas, umass_av=float(z["mu"]), molname="12CO",
        molweight=float(z["molweight"]), ener_cm=z["ener_cm"], gdeg=z["gdeg"], lev_v=z["lev_v"], lev_j=z["lev_j"],
        lev_up=z["lev_up"], lev_down=z["lev_down"], aud=z["aud"], linefreq=linefreq, popul=popul,
        nsize=np.array([1], dtype=np.int32), cont_freq_nu=z["cont_freq_nu"], kappa_abs=z["kappa_abs"],
        kappa_scat=z["kappa_scat"], dust_rho=z["dust_rho"], dust_temp=z["dust_temp"], scati_src=None,
        rstar=float(z["rstar"]), mstar=float(z["mstar"]), tstar=float(z["tstar"]), starspec_cont=z["starspec_cont"],
        incl_deg=float(z["incl"]), nphi=int(z["cir_np"]), nrext=int(z["b_extra"]), dbdr=int(z["b_per_r"]),
        vmax_kms=float(z["passband"]), dv_kms=float(z["vsampling"]), vlsr=float(z["vlsr"]))
    m.extra = dict(psum_temp=z["psum_temp"], psum=z["psum"])
    return m

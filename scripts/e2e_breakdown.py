"""Host wall time of each rl_set_* call and of rl_render from host buffers (the e2e step of bench.py), per config:
python scripts/e2e_breakdown.py 2"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from radlite_b200 import synth
from radlite_b200.api import Renderer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = synth.config(n, **({5: dict(nlines=64)}.get(n, {})))
g = Renderer(0)
g.load_model(m)
g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)
acc = {}
def T(name, f, *a, **k):
    t0 = time.perf_counter(); r = f(*a, **k); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0); return r
reps = 5
for _ in range(reps):
    T("set_grid", g.set_grid, m.r, m.theta)
    T("set_medium", g.set_medium, m.rho, m.abund, m.vel, m.linewidth, m.umass_av)
    T("set_lines", g.set_lines, m.lev_up, m.lev_down, m.linefreq, m.aud, m.gdeg, m.popul)
    T("set_dust", g.set_dust, m.nsize, m.cont_freq_nu, m.kappa_abs, m.kappa_scat, m.dust_rho, m.dust_temp, m.scati_src)
    T("set_camera", g.set_camera, m.anginf, m.nphi, m.nrext, m.dbdr, m.rstar, m.imethod, m.nrref)
    T("set_bc", g.set_bc, m.in_itype, m.out_itype, m.cont_freq_nu, m.starspec_cont, m.isrf_cont)
    T("set_options", g.set_options, m.subgrid, m.nonredundant, m.levthres, m.aksmax)
    T("render", g.render, 1, m.nlines, m.nfr, m.passband, synth.PARSEC)
print("cfg", n, {k: round(1e3 * v / reps, 3) for k, v in acc.items()}, "ms; sum", round(1e3 * sum(acc.values()) / reps, 2))
print("popul MB", m.popul.nbytes / 1e6, "dust_rho MB", np.asarray(m.dust_rho).nbytes / 1e6)

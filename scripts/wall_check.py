"""Opaque-wall start on/off on a small disk: executed-element fraction and bitwise image comparison (GPU)."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from radlite_b200 import synth
from radlite_b200.api import Renderer
m = synth.config(2, nr=60, nth=24, nphi=16, nrext=-8, nlines=8)
g = Renderer(0); g.load_model(m)
g.reset_counters()
on = g.render(1, 8, m.nfr, m.passband, synth.PARSEC, want_image=True); ex_on=g.executed_elements(); c=g.counters()
g.set_wall_tau(0.0); g.reset_counters()
off = g.render(1, 8, m.nfr, m.passband, synth.PARSEC, want_image=True); ex_off=g.executed_elements()
d = np.abs(on["image"]-off["image"]); rel = d/np.maximum(np.abs(off["image"]),1e-300)
print("ex_on/ex_off", ex_on/ex_off, "max rel", rel.max(), "n differ", (d>0).sum(), "of", d.size)
idx = np.unravel_index(np.argmax(rel), rel.shape); print(idx, on["image"][idx], off["image"][idx])
w = np.argwhere(d>0)
print(w[:10])
for tau in (300., 600.):
    g.set_wall_tau(tau); x = g.render(1, 8, m.nfr, m.passband, synth.PARSEC, want_image=True)
    dd = np.abs(x["image"]-off["image"]); print(tau, (dd>0).sum(), (dd/np.maximum(np.abs(off["image"]),1e-300)).max())

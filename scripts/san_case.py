"""Tiny renders under every integrate kernel, for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python scripts/san_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radlite_b200 import api, synth  # noqa: E402

for kernel in (sys.argv[1:] or ["z", "tile", "chan"]):
    api.DEFAULT_KERNEL = kernel
    for cfg, kw in ((2, dict(nr=24, nth=10, nphi=8, nrext=-6, nlines=9)), (1, dict(nr=24, nth=10, nphi=8, nrext=-6))):
        m = synth.config(cfg, **kw)
        g = api.Renderer(0)
        g.load_model(m)
        out = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
        print(kernel, cfg, float(out["flux"].sum()), g.counters(), flush=True)
        g.close()

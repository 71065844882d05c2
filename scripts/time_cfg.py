"""Kernel-phase times of the BASELINE configs (device-resident render): python scripts/time_cfg.py 1 3 4 ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radlite_b200 import synth
from radlite_b200.api import Renderer

for n in [int(a) for a in sys.argv[1:]]:
    kw = {}
    if n == 5:
        kw = dict(nlines=64)  # a 64-line slice of the 2000 (geometry and per-line cost are the same)
    if n == 4:
        kw = dict(nlines=128)
    m = synth.config(n, **kw)
    g = Renderer(0)
    g.load_model(m)
    best = None
    for it in range(3):
        ms = g.render_device(1, m.nlines, m.nfr, m.passband, synth.PARSEC)
        if it and (best is None or ms[2] < best[2]):
            best = ms
    c = g.counters()
    print(f"CFG {n} lines={m.nlines} nray={m.nray} nfr={m.nfr} geometry={best[0]:.2f} prep={best[1]:.2f} "
          f"integrate={best[2]:.2f} flux={best[3]:.2f} total={best[4]:.2f} ms  R/call={c['R']/3:.4g} E/call={c['E']/3:.4g}",
          flush=True)
    g.close()

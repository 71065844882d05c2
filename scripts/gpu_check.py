"""Quick GPU-vs-oracle comparison used during development (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from radlite_b200 import synth
from radlite_b200.api import Renderer
from oracle.oracle_py import Oracle

def compare(m, nl=None, image=True, label=""):
    nl = nl or m.nlines
    g = Renderer(0); g.load_model(m)
    t = time.time(); out = g.render(1, nl, m.nfr, m.passband, synth.PARSEC, want_image=image, want_mask=image); tg = time.time() - t
    t = time.time(); out2 = g.render(1, nl, m.nfr, m.passband, synth.PARSEC); tg2 = time.time() - t
    ms = g.render_device(1, nl, m.nfr, m.passband, synth.PARSEC)
    o = Oracle(); o.load_model(m)
    t = time.time(); ref = o.render(1, nl, m.nfr, m.passband, synth.PARSEC, want_image=image, want_mask=image); to = time.time() - t
    rel = np.abs(out["flux"] - ref["flux"]) / np.abs(ref["flux"])
    print(f"[{label}] {m.name}: nl={nl} gpu {tg:.3f}s/{tg2:.3f}s kernels(ms)={['%.2f'%x for x in ms]} oracle {to:.2f}s")
    print("   flux max rel err", rel.max(), " integrated rel err", abs(out['flux'].sum()-ref['flux'].sum())/ref['flux'].sum())
    print("   counters gpu", g.counters(), "oracle", o.counters(), "nodes", g.total_nodes())
    if image:
        d = np.abs(out["image"] - ref["image"]) / np.maximum(np.abs(ref["image"]), 1e-300)
        print("   image max rel err", d.max(), "cmask equal", np.array_equal(out["cmask"], ref["cmask"]),
              "tau", out["tau_center"], ref["tau_center"], "maser", out["maserflag"], ref["maserflag"])
        if d.max() > 1e-6:
            idx = np.unravel_index(np.argmax(d), d.shape); print("   worst at", idx, out["image"][idx], ref["image"][idx])
    return out, ref

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        compare(synth.config(1, nr=30, nth=12, nphi=8, nrext=-6), label="tiny")
        compare(synth.config(2, nr=40, nth=16, nphi=12, nrext=-8, nlines=6), label="multi")
        compare(synth.config(1), label="cfg1")
    elif which == "cfg2":
        m = synth.config(2)
        compare(m, nl=4, image=False, label="cfg2-4lines")

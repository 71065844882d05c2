"""Work-list statistics of ztile_kernel on a BASELINE config (diagnostic; writes gpurun_out/tile_stats_<cfg>.npz):
python scripts/tile_stats.py 2"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radlite_b200 import synth
from radlite_b200.api import Renderer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
kw = {5: dict(nlines=64), 4: dict(nlines=128)}.get(n, {})
m = synth.config(n, **kw)
g = Renderer(0)
g.load_model(m)
g.render_device(1, m.nlines, m.nfr, m.passband, synth.PARSEC)
f = g.lib.rl_debug_fetch
f.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]
f.restype = C.c_longlong
out = {}
for name, dt in (("ztiles", np.uint8), ("nstart", np.int32), ("node_off", np.int64), ("rng", np.int32),
                 ("nitems", np.uint32), ("zlines", np.uint16)):
    nb = f(g.ctx, name.encode(), None, 0)
    buf = np.zeros(max(nb, 1), dtype=np.uint8)
    f(g.ctx, name.encode(), buf.ctypes.data_as(C.c_void_p), nb)
    out[name] = buf[:nb].view(dt)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(f"gpurun_out/tile_stats_{n}.npz", nl=m.nlines, nfr=m.nfr, nray=m.nray, **out)
print("saved", {k: v.shape for k, v in out.items()})

import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radlite_b200 import synth
from radlite_b200.api import Renderer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
kw = {5: dict(nlines=64), 4: dict(nlines=128)}.get(n, {})
m = synth.config(n, **kw)
g = Renderer(0); g.load_model(m); g.reset_counters()
g.render_device(1, m.nlines, m.nfr, m.passband, synth.PARSEC)
f = g.lib.rl_debug_fetch; f.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]; f.restype = C.c_longlong
buf = np.zeros(8, dtype=np.uint64); f(g.ctx, b"counters", buf.ctypes.data_as(C.c_void_p), 64)
print("CFG", n, "thick groups", buf[4], "far-wing thick groups", buf[5], "frac", buf[5] / max(1, buf[4]))

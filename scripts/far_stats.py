"""Path statistics of ztile_kernel (needs a library built with -DRL_STATS: make variant NAME=stats DEFS=-DRL_STATS;
RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_stats.so python scripts/far_stats.py 2)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radlite_b200 import synth
from radlite_b200.api import Renderer
for n in [int(a) for a in sys.argv[1:]] or [2]:
    kw = {5: dict(nlines=64), 4: dict(nlines=128)}.get(n, {})
    m = synth.config(n, **kw)
    g = Renderer(0); g.load_model(m); g.reset_counters()
    g.render_device(1, m.nlines, m.nfr, m.passband, synth.PARSEC)
    f = g.lib.rl_debug_fetch; f.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]; f.restype = C.c_longlong
    buf = np.zeros(16, dtype=np.uint64); f(g.ctx, b"counters", buf.ctypes.data_as(C.c_void_p), 128)
    b = buf.astype(float)
    print("CFG", n, "R E S X", buf[:4])
    print("  node-thin steps: channel slots", buf[4], "general groups that only mix thin and thick channels", buf[7],
          "of", buf[14])
    print("  other unflagged steps: channel slots", buf[8], "far@345/100/64", b[9:12] / max(1, b[8]))
    print("  3-channel groups thin/stream/general", buf[12:15], "flagged node steps", buf[15])
    g.close()

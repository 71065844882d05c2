import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from radlite_b200 import synth, api
from radlite_b200.api import Renderer
from test_gpu_wall import layered_shell
def fetch(g, name, dt):
    f = g.lib.rl_debug_fetch; f.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]; f.restype = C.c_longlong
    nb = f(g.ctx, name.encode(), None, 0); a = np.zeros(nb, dtype=np.uint8); f(g.ctx, name.encode(), a.ctypes.data_as(C.c_void_p), nb)
    return a.view(dt)
api.DEFAULT_KERNEL = "z"
for t in (0.0, 60.0):
    m = layered_shell(t)
    g = Renderer(0); g.load_model(m); g.reset_counters()
    out = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True)
    ns = fetch(g, "nstart", np.int32)[:m.nray]; off = fetch(g, "node_off", np.int64)
    smin = fetch(g, "smin", np.float64).reshape(len(m.r), len(m.theta)); admin = fetch(g, "admin", np.float64).reshape(len(m.r), len(m.theta))
    ws = fetch(g, "wstat", np.uint64)
    print("T", t, "smax", ws[:1].view(np.float64), "inv", ws[1])
    print(" smin by radius", smin[::3, 0], "\n admin", admin[::6, 0])
    N = np.diff(off)
    for r in (1, 50, 100, 200, 258): print("  ray", r, "N", N[r], "nstart", ns[r])
    g.set_wall_tau(0.0)
    off_img = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True)["image"]
    print(" images differ in", (off_img != out["image"]).sum(), "of", off_img.size, "max", out["image"][:,1:].max())

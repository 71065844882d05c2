"""One camera-ring block of a config on one GPU (what a rank of an N-GPU run does): phase times, for an ncu
launch list.  python scripts/ring_block_probe.py <cfg> <lo> <hi> [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from radlite_b200 import synth  # noqa: E402
from radlite_b200.api import Renderer  # noqa: E402

cfg, lo, hi = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
m = synth.config(cfg)
g = Renderer(0)
g.load_model(m)
nrr, nphi, nray = g.camera_dims()
rs = torch.zeros((m.nlines, nrr + 1, m.nfr), dtype=torch.float64, device="cuda:0")
for it in range(reps):
    g.invalidate_geometry()
    t = g.render_rings_device(1, m.nlines, m.nfr, m.passband, synth.PARSEC, lo, hi, rs.data_ptr())
    print(f"rings {lo}..{hi}: geometry={t[0]:.3f} prep={t[1]:.3f} integrate={t[2]:.3f} flux={t[3]:.3f} total={t[4]:.3f} ms",
          flush=True)

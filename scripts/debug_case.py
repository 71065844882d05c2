"""Development aid (run under gpurun): locate the worst pixel of a GPU-vs-oracle comparison and dump
the ray it belongs to."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from helpers import clone, tiny
from oracle.oracle_py import Oracle
from radlite_b200 import synth
from radlite_b200.api import Renderer

m = clone(tiny(2, nlines=2), incl_deg=60.0)
g = Renderer(0)
g.load_model(m)
out = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
o = Oracle()
o.load_model(m)
ref = o.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC, want_image=True, want_mask=True)
d = np.abs(out["image"] - ref["image"]) / np.maximum(np.abs(ref["image"]), 1e-300)
print("image shape", d.shape, "max", d.max())
order = np.argsort(d.ravel())[::-1][:12]
for k in order:
    idx = np.unravel_index(k, d.shape)
    print(idx, d[idx], out["image"][idx], ref["image"][idx], "cmask", out["cmask"][idx], ref["cmask"][idx])
l, ir, ip, ch = np.unravel_index(order[0], d.shape)
iray = 1 if ir == 0 else 2 + (ir - 1) * m.nphi + ip
print("worst: line", l, "ring", ir, "phi", ip, "channel", ch, "ray (1-based)", iray)
nd = g.ray_nodes(iray)
v = o.node_values(iray, l + 1)
print("nodes", len(nd["ds"]))
print("flags", nd["flags"])
print("channels with err>1e-7 on this ray/line:", np.nonzero(d[l, ir, ip] > 1e-7)[0])
print("err per channel", d[l, ir, ip])
lwav = 0.5 * (nd["lw"][1:] + nd["lw"][:-1])
q = np.abs(np.diff(nd["dvmu"])) / (lwav / 2.99792458e5)
print("6q", 6 * q)
print("ds", nd["ds"])
print("alpd", v["alpd"])
print("nup", v["nup"])
print("dvmu", nd["dvmu"])
print("velo ch", out["velo"][l][ch], "lw", nd["lw"])

"""Independent check of the ray geometry (closes SURVEY.md A.2.8: "theta-crossing ordering == sort by s").

Nothing here shares code or formulas with oracle/radlite_oracle.c: every ray is intersected with every
grid surface by brute force in numpy -- spheres R_i and the cones theta_j of BOTH hemispheres written in
cos(theta) form (the reference and the oracle use the tan^2 form and a hand-ordered quadrant logic,
telescope.F:2959-3194) -- the roots are sorted by path length and filtered with the radius windows of
telescope.F:3352-3366.  The resulting list of (surface type, surface index, s) must be exactly the list of
R- and theta-crossing nodes that orc_trajectory (the restatement of make_trajectory_c) returns, for rays
in both hemispheres, inside and outside the inner hole, including the "double solution" rays that cross a
cone twice on the same side.
"""
import numpy as np
import pytest

from helpers import tiny
from radlite_b200 import synth

PI_MIRROR = 3.14159265359  # grid.F:1163: the mirrored theta values are built with this literal


def camera(m, rays_r):
    """telescope.F:863-897: start point (x0, z0) of ray i = 1 + (ir-1) nphi + k; ray 0 is the centre."""
    anginf = m.anginf
    if abs(anginf) < 0.1:  # telescope.F:818-827: the inclination is clamped away from pole-on
        anginf = 0.1 * np.sign(anginf)
    theta0 = anginf + 1.0e-4
    st0 = np.sin(theta0)
    k = np.arange(m.nphi)
    phi = (k + 0.5) * (6.28318530718 / m.nphi)
    zh = -np.sin(phi) / st0
    xh = np.sign(np.cos(phi)) * np.sqrt(1.0 - zh * zh * st0 * st0 + 1.0e-8)
    x0 = np.concatenate(([0.0], (rays_r[1:, None] * xh[None, :]).ravel()))
    z0 = np.concatenate(([0.0], (rays_r[1:, None] * zh[None, :]).ravel()))
    return theta0, x0, z0


def brute_force_crossings(r, theta_up, theta0, x0, z0):
    """All crossings of the ray (x0, s sin(theta0), z0 + s cos(theta0)) with the spheres r[i] and the cones
    theta[j], j = 1..nt, sorted by s.  Returns arrays (kind, index, s): kind 1 = sphere, 2 = cone."""
    nr, nth = len(r), len(theta_up)
    nt = 2 * nth
    c0, s0 = np.cos(theta0), np.sin(theta0)
    theta = np.concatenate((theta_up, (PI_MIRROR - theta_up)[::-1]))  # theta_j, j = 1..nt
    out = []
    # spheres: |p(s)|^2 = R^2
    for i in range(nr):
        b = 2.0 * z0 * c0
        c = x0 * x0 + z0 * z0 - r[i] * r[i]
        disc = b * b - 4.0 * c
        if disc > 0.0:
            for sg in (-1.0, 1.0):
                out.append((1, i + 1, 0.5 * (-b + sg * np.sqrt(disc))))
    # cones: z(s) = |p(s)| cos(theta_j), same sign on both sides
    for j in range(nt):
        ct = np.cos(theta[j])
        ct2 = ct * ct
        a = c0 * c0 - ct2
        b = 2.0 * z0 * c0 * (1.0 - ct2)
        c = z0 * z0 * (1.0 - ct2) - ct2 * x0 * x0
        disc = b * b - 4.0 * a * c
        if disc <= 0.0 or a == 0.0:
            continue
        for sg in (-1.0, 1.0):
            s = (-b + sg * np.sqrt(disc)) / (2.0 * a)
            z = z0 + s * c0
            if z * ct > 0.0:  # the root lies on this cone, not on its mirror image
                out.append((2, j + 1, s))
    out.sort(key=lambda t: t[2])
    return out


def windowed(cross, r, theta0, x0, z0):
    """Radius windows of telescope.F:3352-3366 (theta crossings strictly inside the radial grid, sphere
    crossings including the two edge spheres)."""
    c0 = np.cos(theta0)
    keep = []
    for kind, idx, s in cross:
        rad = np.sqrt(x0 * x0 + z0 * z0 + s * s + 2.0 * z0 * c0 * s)
        if kind == 2 and not (r[0] * (1 + 1e-9) <= rad <= r[-1] * (1 - 1e-9)):
            continue
        if kind == 1 and not (r[0] * (1 - 1e-9) <= rad <= r[-1] * (1 + 1e-9)):
            continue
        keep.append((kind, idx, s))
    return keep


@pytest.mark.parametrize("incl", [15.0, 60.0, 2.0])
def test_trajectory_nodes_against_brute_force(oracle_cls, incl):
    m = tiny(1, nr=36, nth=14, nphi=24, nrext=-8)
    m.incl_deg = incl
    o = oracle_cls()
    o.load_model(m)
    nrr, nphi, nray = o.camera_dims()
    rays_r, _ = o.rings()
    theta0, x0, z0 = camera(m, rays_r)
    assert nray == len(x0)
    checked = doubles = lower = 0
    scale = m.r[-1]
    for iray in range(2, nray + 1):  # every off-centre ray: 1 + nrr * nphi - 1 >= 1000
        t = o.trajectory(iray)
        sel = t["icross"] != 3
        got = list(zip(t["icross"][sel], np.where(t["icross"][sel] == 1, t["iradius"][sel], t["itheta"][sel]),
                       t["s"][sel]))
        want = windowed(brute_force_crossings(m.r, m.theta, theta0, x0[iray - 1], z0[iray - 1]), m.r, theta0,
                        x0[iray - 1], z0[iray - 1])
        assert [(k, i) for k, i, _ in got] == [(k, i) for k, i, _ in want], iray
        gs, ws = np.array([s for *_, s in got]), np.array([s for *_, s in want])
        # the reference pads the sphere discriminant by 1e-10 b^2 (telescope.F:3296): s moves by ~1e-10 |b|,
        # with |b| = 2 |z0 cos(theta0)| up to the path length itself for nearly pole-on cameras
        assert np.all(np.abs(gs - ws) <= 1e-9 * (scale + np.abs(ws))), iray
        assert np.all(np.diff(t["s"]) > 0)
        cones = [i for k, i, _ in want if k == 2]
        doubles += len(cones) != len(set(cones))
        lower += any(i > len(m.theta) for i in cones)
        checked += 1
    assert checked >= 1000
    assert lower > 0  # the lower-hemisphere branch was hit
    if theta0 > m.theta[0]:  # cones steeper than the line of sight exist: "double solution" rays (telescope.F:3008)
        assert doubles > 0


def test_centre_ray_crosses_every_sphere_twice(oracle_cls):
    m = tiny(1, nr=36, nth=14, nphi=24, nrext=-8)
    o = oracle_cls()
    o.load_model(m)
    t = o.trajectory(1)
    rr = t["iradius"][t["icross"] == 1]
    # inbound nr..1, (vacuum), outbound 1..nr -- the star sphere itself is outside the radial windows
    assert list(rr) == list(range(len(m.r), 0, -1)) + list(range(1, len(m.r) + 1))

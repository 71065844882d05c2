"""Write a RADLite working directory from a synth.Model the way the reference drivers do, and read
the outputs back the way pyradlite does.  Test infrastructure.

Writers follow pyradlite/pyradlite/radlite.py (radlite.inp :2105, linespectrum.inp :1846, moldata.dat
:1920, levelpop :1775, density/abundance/velocity/turbulence :2015-2221) and the RADMC files the
Fortran readers expect.  ``read_linespectrum`` restates
radlite.py:2750-2877 (_read_core_radliteoutput); ``load_dump`` reads radlite_b200_host --dump files.
"""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "radlite_b200", "radlite_b200_host")


def write_workdir(m, d, nlines=None, image=0, comments=True, imager=None):
    """``imager`` (with image=2): dict(nx, ny, size_x, size_y [cm, full widths], phioff, xoff, yoff, addstar) -- the
    block telescope.F:203-216 reads after the line range."""
    os.makedirs(d, exist_ok=True)
    nr, nth = len(m.r), len(m.theta)
    nl = m.nlines if nlines is None else nlines

    def w(name, txt):
        with open(os.path.join(d, name), "w") as f:
            f.write(txt)

    s = "104\t\tInput format version\n" + "=" * 40 + "\n"
    s += "-1\t\tMaximum number; OBSOLETE NITER PARAMETER\n2\t\tIteration method\n0\t\tFlux conservation trick\n"
    s += "1.000e-06\t\tConvergence tolerance\n0\t\tConvergence crit type\n1\t\tInitial guess type\n" + "=" * 40 + "\n"
    s += "41\t\tNr of mu-angle points\n16\t\tNr of phi-angle points\n2\t\tType of mu gridding\n0.1\t\tLargest allowed error\n"
    s += "1.0\t\tDmu wrt dR\n3\t\tExtra mus around mu=0\n" + "=" * 40 + "\n"
    s += f"{m.out_itype}\t\tType of outer boundary\n2\t\tType of inner boundary\n0\t\tType of equator boundary\n" + "=" * 40 + "\n"
    s += f"57.3\t\tDefault inclination angle\n{m.nphi}\t\tNr Phi-points circular CCD\n{m.dbdr}\t\tThe number of b_i per r_i\n"
    s += f"{m.nrext}\t\tThe number of extra b_i inside the inner radius\n" + "=" * 40 + "\n"
    s += "1\t\tSave NLTE\n0\t\tSave intensity at inu\n-1\t\tSave source at inu\n0\t\tSave moments\n-1\t\tSave dusttemp/etc\n"
    s += "0\t\tSave ALI operator\n1\t\tSave phys vars of medium\n0\t\tSave flux conservation data\n" + "=" * 40 + "\n"
    s += "-551\t\tType of setup\n0\t\tDo dust?\n1\t\tDo lines?\n1\t\tInclude dust in lines?\n1\t\tInclude star pumping?\n" + "=" * 40 + "\n"
    w("radlite.inp", s)

    w("radius.inp", f"{nr:12d}\n \n" + "".join(f"{x:16.7e}\n" for x in m.r))
    w("theta.inp", f"{nth:12d}{1:8d}\n \n" + "".join(f"{x:16.8f}\n" for x in m.theta))
    w("frequency.inp", f"{len(m.cont_freq_nu):12d}\n \n" + "".join(f"{x:13.6E}\n" for x in m.cont_freq_nu))
    nspec = len(m.nsize)
    s = f"{nspec:12d}{nr:12d}{nth:12d}{1:8d}\n \n"
    for isp in range(nspec):
        s += "".join(f"{m.dust_rho[ir, it, isp]:16.7e}\n" for ir in range(nr) for it in range(nth))
    w("dustdens.inp", s)
    w("dusttemp.info", "          -2\n           1\n")
    s = f"{nspec:12d}{nr:12d}{nth:12d}{1:12d}\n\n"
    for isp in range(nspec):
        s += f"{int(m.nsize[isp]):12d}\n"
        for iz in range(int(m.nsize[isp])):
            s += "".join(f"{m.dust_temp[ir, it, isp, iz]:22.16g}     \n" for ir in range(nr) for it in range(nth))
    w("dusttemp_final.dat", s)
    s = f"2               Format number of this file\n{nspec}               Nr of dust species\n" + "=" * 76 + "\n"
    for isp in range(nspec):
        s += "-1              Way in which this dust species is read (-1=file)\n0               0=Thermal grain\n"
        s += f"{isp + 1}               Extension of name of dustopac_***.inp file\n" + "-" * 76 + "\n"
        ncf, nz = len(m.cont_freq_nu), int(m.nsize[isp])
        t = f"{ncf:12d}{nz:8d}\n \n"
        t += "".join(f"{m.kappa_abs[isp, iz, k]:16.8g}\n" for k in range(ncf) for iz in range(nz)) + " \n"
        t += "".join(f"{m.kappa_scat[isp, iz, k]:16.8g}\n" for k in range(ncf) for iz in range(nz))
        w(f"dustopac_{isp + 1}.inp", t)
    w("dustopac.inp", s)
    if m.scati_src is not None:
        ncf = len(m.cont_freq_nu)
        s = f"{ncf} {nth} {nr} 1\n"
        s += "".join(f"{m.scati_src[ir, it, k]:.8e}\n" for k in range(ncf) for it in range(nth) for ir in range(nr))
        w("scatsource.dat", s)
    w("starinfo.inp", f"1\n{m.rstar:16.7e}\n{m.mstar:16.7e}\n{m.tstar:16.4f}\n")
    # F_nu at 1 pc from the surface intensity (star.F:510): I = 3.0308410d36 F / R*^2
    fnu = m.starspec_cont * m.rstar**2 / 3.0308410e36
    w("starspectrum.inp", f"{len(fnu):12d}\n" + "".join(f"{a:16.7e}{b:16.7e}\n" for a, b in zip(m.cont_freq_nu, fnu)))
    if m.out_itype == 3:
        w("interstellfield.inp", f"{len(m.isrf_cont)}\n" + "".join(f"{x:.9e}\n" for x in m.isrf_cont))

    # gas (pyradlite writers)
    w("density.inp", f"{nr:d}\t{nth:d}\t1\n" + "".join(f"{m.rho[ir, it]:.8e}\n" for ir in range(nr) for it in range(nth)))
    w("abundance.inp", f"{nr:d}\t{nth:d}\n" + "".join(f"{m.abund[ir, it]:.8e}\t{0.0:.8f}\n" for ir in range(nr) for it in range(nth)))
    w("velocity.inp", f"{nr:d}\t{nth:d}\n" + "".join(
        "{0:.8f}\t{1:.8f}\t{2:.8f}\n".format(*m.vel[ir, it]) for ir in range(nr) for it in range(nth)))
    w("turbulence.inp", f"1\n{nr:d}\t{nth:d}\n" + "".join(f"{m.linewidth[ir, it]:.8f}\n" for ir in range(nr) for it in range(nth)))
    s = "2               Format number (1=<Apr06,2=>Apr06)\n1\t\tHow many molecule species (1 for now)\n"
    s += f"{m.umass_av}\t\tAverage molecular weight\n" + "=" * 76 + "\n" + "-" * 76 + "\n"
    s += "1\t\tYes, specify info for line 2 -> 1\n0\t\tSymmetry in the line\n40\t\tNumber of frequency points in line\n"
    s += "80.0e0\t\tWidth of line range\n1\t\tLine profile\n0\t\tAdditional information\n" if not comments else \
        "80.0e0\t\tWidth of line range\n1\t\tLine profile\n1\t\tAdditional information: fixed-line-width\n1.0d0\t\tFixed line width\n0\t\tAdditional information\n"
    s += ("-" * 76 + "\n0\t\tNo, Same as previous\n") * max(0, m.nlines - 1)
    w("line.inp", s)

    # molecule + level populations (radlite.py:1920, 1775)
    nlev = len(m.gdeg)
    s = f"!MOLECULE\n{m.molname}\n!MOLECULAR WEIGHT\n{m.molweight:4.1f}\n!NUMBER OF ENERGY LEVELS\n{nlev:6d}\n"
    s += "!LEVEL + ENERGIES(cm^-1) + WEIGHT + v + Q\n"
    for k in range(nlev):
        s += "{0:5d}{1:12.4f}{2:7.1f}{3:>15s}{4:>15s}\n".format(k + 1, m.ener_cm[k], m.gdeg[k], str(int(m.lev_v[k])), str(int(m.lev_j[k])))
    s += f"!NUMBER OF RADIATIVE TRANSITIONS\n{m.nlines:6d}\n!TRANS + UP + LOW + EINSTEINA(s^-1) + FREQ(cm^-1) + E_u(cm^-1) + v_l + Q_p + Q_pp\n"
    for k in range(m.nlines):
        up, dn = m.lev_up[k], m.lev_down[k]
        s += "{0:5d}{1:5d}{2:5d}{3:12.3e}{4:16.7f}{5:12.5f}{6:>15}{7:>15}{8:>15}{9:>15}\n".format(
            k + 1, up, dn, m.aud[k], m.ener_cm[up - 1] - m.ener_cm[dn - 1], m.ener_cm[up - 1],
            int(m.lev_v[up - 1]), int(m.lev_v[dn - 1]), int(m.lev_j[up - 1]), int(m.lev_j[dn - 1]))
    w("moldata.dat", s)
    s = f"{nr:d}\t{nth:d}\t{nlev:d}\t1\n" + "\t".join(repr(float(x)) for x in m.ener_cm) + "\n"
    s += "\t".join(repr(float(x)) for x in m.gdeg) + "\n"
    s += "".join("\n".join(repr(float(x)) for x in m.popul[ir, it]) + "\n" for ir in range(nr) for it in range(nth))
    w("levelpop_moldata.dat", s)
    w("levelpop.info", "-3\nlevelpop_moldata.dat\n0")

    s = "{0:<8d}Format number\n{1:<8d}Spectrum output style\n".format(1, 1) + "-" * 63 + "\n"
    s += "{0:<8d}Format number\n".format(2) + "-" * 63 + "\n"
    s += "{0:<8.2f}{1:<50s}\n".format(m.vmax_kms, "Width of line passband [km/s]")
    s += "{0:<8.2f}{1:<50s}\n".format(m.dv_kms, "Velocity sampling [km/s]") + "-" * 65 + "\n"
    s += "{0:<8d}Format number\nmoldata.dat\tMolecular data file\n".format(2)
    s += "{0:<8d}{1:<50s}\n".format(image, "Command (0=spectrum, 2=image[3-D P/V cube])")
    s += "{0:<8.1f}{1:<50s}\n{2:<8.1f}{3:<50s}\n".format(1.0, "Distance in [pc]", m.incl_deg, "Inclination [deg]")
    s += "{0:<8.1f}{1:<60s}\n".format(m.vlsr, "Radial velocity [km/s]")
    s += "{0:<8d}{1:<50s}\n{2:<8d}Starting line to make spectrum/image\n".format(nl, "Nr of lines to make spectrum/image", 1)
    if image == 2:
        im = imager
        s += "{0:<8d}Nr of pixels in x\n{1:<8d}Nr of pixels in y\n{2:<8d}Size specifier (1 = cm)\n".format(im["nx"], im["ny"], 1)
        s += "{0:<24.16e}Image width in x\n{1:<24.16e}Image width in y\n".format(im["size_x"], im["size_y"])
        s += "{0:<24.16e}Rotation angle\n{1:<24.16e}X offset\n{2:<24.16e}Y offset\n".format(im["phioff"], im["xoff"], im["yoff"])
        s += "{0:<8d}Add the unresolved central star\n".format(im["addstar"])
    w("linespectrum.inp", s)


def run_host(d, *args, check=True):
    p = subprocess.run([HOST, "--dir", d, *args], capture_output=True, text=True)
    if check and p.returncode != 0:
        raise RuntimeError(f"radlite_b200_host failed ({p.returncode}):\n{p.stdout}\n{p.stderr}")
    return p


def load_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            nm = f.read(32)
            if len(nm) < 32:
                break
            name = nm.split(b"\0")[0].decode()
            dt = f.read(1).decode()
            nd, = struct.unpack("i", f.read(4))
            dims = struct.unpack(f"{nd}q", f.read(8 * nd))
            n = int(np.prod(dims))
            a = np.frombuffer(f.read(n * (8 if dt == "d" else 4)), dtype=np.float64 if dt == "d" else np.int32)
            out[name] = a.reshape(dims).copy()
    return out


def write_records(path, **arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            dt = "d" if a.dtype == np.float64 else "i"
            assert a.dtype in (np.float64, np.int32)
            f.write(name.encode().ljust(32, b"\0") + dt.encode() + struct.pack("i", a.ndim) + struct.pack(f"{a.ndim}q", *a.shape))
            f.write(a.tobytes())


def model_from_dump(m, dmp):
    """A copy of synth model ``m`` carrying exactly the numbers the host parsed from the text files."""
    import copy
    m2 = copy.deepcopy(m)
    for k in ("r", "theta", "cont_freq_nu", "kappa_abs", "kappa_scat", "dust_rho", "dust_temp", "rho", "abund",
              "vel", "linewidth", "starspec_cont", "gdeg", "aud", "linefreq", "lev_up", "lev_down", "popul"):
        setattr(m2, k, dmp[k])
    m2.scati_src = dmp.get("scati_src")
    m2.isrf_cont = dmp.get("isrf_cont")
    sc = dmp["scalars"]
    m2.umass_av, m2.rstar = float(sc[0]), float(sc[1])
    return m2


def read_linespectrum(path):
    """radlite.py:2833-2851: header lines 4/5/6, then per line numpoints at iloc+5 and data at iloc+7."""
    spec = open(path).readlines()
    numlines, maxn = int(spec[4]), int(spec[5])
    dist, vlsr, incl = (float(x) for x in spec[6].split())
    iloc, out = 6, []
    for _ in range(numlines):
        npnt = int(spec[iloc + 5])
        lev = [int(x) for x in spec[iloc + 2].split()]
        sec = np.array([ln.split() for ln in spec[iloc + 7:iloc + 7 + npnt]]).astype(float)
        out.append(dict(vel=sec[:, 0], flux=sec[:, 1], lev=lev, freq=float(spec[iloc + 3]), beam=float(spec[iloc + 4])))
        iloc = iloc + 7 + npnt - 1
    return dict(numlines=numlines, maxnumpoints=maxn, dist=dist, vlsr=vlsr, incl=incl, lines=out, raw=spec)


def read_imcir(path):
    """PRO/read_imcir.pro layout: nfr; nu0; nphi nrr; ri(nrr+1); r(nrr+1); per channel: velocity, centre, rows."""
    tok = open(path).read().split()
    p = 0

    def take(n, conv=float):
        nonlocal p
        v = [conv(x) for x in tok[p:p + n]]
        p += n
        return v
    nfr = take(1, int)[0]
    nu0 = take(1)[0]
    nphi, nrr = take(2, int)
    ri, r = np.array(take(nrr + 1)), np.array(take(nrr + 1))
    vel, cen = np.zeros(nfr), np.zeros(nfr)
    img, msk = np.zeros((nfr, nphi, nrr)), np.zeros((nfr, nphi, nrr), dtype=int)
    for k in range(nfr):
        vel[k], cen[k] = take(1)[0], take(1)[0]
        rows = np.array(take(2 * nphi * nrr)).reshape(nphi, nrr, 2)
        img[k], msk[k] = rows[..., 0], rows[..., 1].astype(int)
    return dict(nfr=nfr, nu0=nu0, nphi=nphi, nrr=nrr, ri=ri, r=r, vel=vel, centre=cen, image=img, cmask=msk)


def read_posvel(path):
    """lineposvel_<mol>_<n>.dat as calc_write_line_posvel writes it (telescope.F:1934-2040)."""
    L = open(path).read().split("\n")
    assert L[0].strip() == "" and int(L[1]) == 1
    dist, vlsr, incl = (float(x) for x in L[4].split())
    lev = [int(x) for x in L[5].split()]
    nu0, nfr = float(L[6]), int(L[7])
    t = L[8].split()
    nx, ny = int(t[0]), int(t[1])
    spx, spy, phioff, xoff, yoff = (float(x) for x in t[2:7])
    tok = " ".join(L[9:]).split()
    vel = np.array(tok[:nfr], dtype=float)
    rest = np.array(tok[nfr:nfr + 2 * nfr * nx * ny], dtype=float).reshape(nfr, ny, nx, 2)
    return dict(dist=dist, vlsr=vlsr, incl=incl, lev=lev, nu0=nu0, nfr=nfr, nx=nx, ny=ny, spx=spx, spy=spy,
                phioff=phioff, xoff=xoff, yoff=yoff, vel=vel, temp=np.transpose(rest[..., 0], (2, 1, 0)),
                tau=np.transpose(rest[..., 1], (2, 1, 0)), raw=L)

// rl_geom.cu -- ray geometry on the device: one thread per camera ray builds the ray's node list
// (crossings with the R and Theta grid surfaces plus the extremum sampling points), the
// interpolation stencil of every node and the line-independent node quantities
// (segment length, projected velocity Omega.v/c, line width).  Built once per camera/grid and
// reused by every line and channel.
//
// Follows telescope.F:2787-3720 (make_trajectory_c) and the node bookkeeping of
// telescope.F:3948-3958, 4046-4104, 4148-4193 (charintline) and line.F:3986-4039, 4090-4094,
// 2650-2662 (stencil, velocity interpolation, omega_dot_v).  Unlike the reference nothing is
// materialised per surface: the three crossing families are generated as sorted streams and merged
// on the fly, so a ray needs O(1) scratch.
//
// Compile this file with --fmad=false: the guards (1e-10 discriminant pad, 1e-9 radius windows,
// hunt brackets) are meant to see the same rounding as the reference's separate mul/add.
#include "rl_types.h"

#include <math.h>

namespace rl {

namespace {

__device__ __forceinline__ double RCf(const GridDev &g, int i) { return g.rc[i + 1]; }
__device__ __forceinline__ double TCf(const GridDev &g, int i) { return g.tc[i + 1]; }
__device__ __forceinline__ int RIDXf(const GridDev &g, int i) { return g.ridx[i + 4]; }

// nrecip.F:157 hunt, bisection phase from (0, n+1): largest j in 0..n with xx(j) < x.
// xx1 points at xx(1).
__device__ int hunt_bisect(const double *xx1, int n, double x) {
  int jlo = 0, jhi = n + 1;
  while (jhi - jlo != 1) {
    int jm = (jhi + jlo) >> 1;
    if (x > xx1[jm - 1]) jlo = jm;
    else jhi = jm;
  }
  return jlo;
}

struct Ray {
  GridDev g;
  const double *tan2;  // [nt/2 + 1] tan^2 of the cone angles (index iy = 1..nt/2)
  const double4 *thr;  // [nt/2] this ray's cone roots {sar1, sar2, radius(sar1), radius(sar2)} (roots_kernel); sar1 = NaN: no hit
  const double2 *rrt;  // [nr+1] this ray's sphere roots {s-, s+} of spheres ix = 0 (the star) .. nr (roots_kernel)
  // the two crossing streams of the ray in path order, staged in shared memory by the lanes of the ray's warp
  const double *sm_th_s, *sm_th_rad, *sm_r_s;
  const int *sm_th_it;
  bool staged;  // false: one thread per ray, the streams are read from the tables
  double x0, z0, costh0, sinth0, costh02, sinth02, pitheta0;
  // theta-crossing stream
  int isnr, isdbl, iy_first, iup, branchA, ith_amount;
  // R-crossing stream
  int ir_min, ir_amount;
  double rstar;
};

// both roots of the theta=const cone iy (telescope.F:2960-2990), ordered sar1 <= sar2
__device__ bool th_roots_math(double x0, double z0, double costh0, double costh02, double sinth02, double tanth2,
                              double &sar1, double &sar2) {
  double a = tanth2 * costh02 - sinth02;
  double b = 2.0 * tanth2 * costh0 * z0;
  double c = tanth2 * z0 * z0 - x0 * x0;
  double sdiscr = b * b - 4.0 * a * c;
  if (!(sdiscr > 0.0)) return false;
  sdiscr = sqrt(sdiscr);
  sar1 = (-b - sdiscr) / (2.0 * a);
  sar2 = (-b + sdiscr) / (2.0 * a);
  if (sar1 > sar2) {
    double d = sar1;
    sar1 = sar2;
    sar2 = d;
  }
  return true;
}
// the same from the ray's table (filled by roots_kernel: the expensive arithmetic of a ray's crossings is done
// once, by one thread per (ray, surface), instead of serially and repeatedly inside the per-ray merge)
__device__ __forceinline__ bool th_roots(const Ray &R, int iy, double &sar1, double &sar2) {
  if (!R.staged)  // one thread per ray: the arithmetic inline (table reads would only add latency to the chain)
    return th_roots_math(R.x0, R.z0, R.costh0, R.costh02, R.sinth02, R.tan2[iy], sar1, sar2);
  const double4 v = R.thr[iy - 1];
  if (!(v.x == v.x)) return false;
  sar1 = v.x;
  sar2 = v.y;
  return true;
}

// k-th theta crossing along the ray, k = 1..ith_amount (telescope.F:3024-3194 ordering)
__device__ void th_elem_tab(const Ray &R, int k, double &s, int &itheta, double &rad) {
  const int n = R.isnr, d = R.isdbl;
  int is, which, pat;  // which: 0 = sar1, 1 = sar2 ; pat 1: iup ? nt+1-iy : iy ; pat 2: iup ? iy : nt+1-iy
  if (R.branchA) {
    if (k <= n - d) { is = d + k; which = 0; pat = 1; }
    else if (k <= 2 * n - 2 * d) { is = n - (k - (n - d + 1)); which = 1; pat = 2; }
    else if (k <= 2 * n - d) { is = d - (k - (2 * n - 2 * d + 1)); which = 0; pat = 2; }
    else { is = 1 + (k - (2 * n - d + 1)); which = 1; pat = 2; }
  } else {
    if (k <= d) { is = d - (k - 1); which = 0; pat = 1; }
    else if (k <= 2 * d) { is = k - d; which = 1; pat = 1; }
    else if (k <= n + d) { is = d + 1 + (k - (2 * d + 1)); which = 0; pat = 1; }
    else { is = n - (k - (n + d + 1)); which = 1; pat = 2; }
  }
  int iy = R.iy_first + is - 1;
  if (!R.staged) {
    double a1 = 0.0, a2 = 0.0;
    th_roots_math(R.x0, R.z0, R.costh0, R.costh02, R.sinth02, R.tan2[iy], a1, a2);
    s = which ? a2 : a1;
    rad = sqrt(R.x0 * R.x0 + R.z0 * R.z0 + s * s + 2.0 * R.z0 * R.costh0 * s);
  } else {
    const double4 v = R.thr[iy - 1];
    s = which ? v.y : v.x;
    rad = which ? v.w : v.z;
  }
  bool mirrored = (pat == 1) ? (R.iup == 1) : (R.iup != 1);
  itheta = mirrored ? (R.g.nt + 1 - iy) : iy;
}
// the same from the staged stream
__device__ __forceinline__ void th_elem(const Ray &R, int k, double &s, int &itheta, double &rad) {
  if (!R.staged) {
    th_elem_tab(R, k, s, itheta, rad);
    return;
  }
  s = R.sm_th_s[k - 1];
  rad = R.sm_th_rad[k - 1];
  itheta = R.sm_th_it[k - 1];
}
__device__ double th_s_at(const Ray &R, int k) {
  if (k > R.ith_amount) return 1.e30;
  if (!R.staged) {
    double s, rad;
    int it;
    th_elem_tab(R, k, s, it, rad);
    return s;
  }
  return R.sm_th_s[k - 1];
}

// k-th R crossing, k = 1..ir_amount (telescope.F:3284-3335); the roots come from the ray's table
__device__ double r_elem_tab(const Ray &R, int k) {
  int j = (k <= R.ir_amount / 2) ? k : (R.ir_amount + 1 - k);
  const int ix = R.g.nr - (j - 1);
  if (!R.staged) {
    const double r = (ix == 0) ? R.rstar : RCf(R.g, ix);
    const double b = 2.0 * R.costh0 * R.z0;
    const double c = R.x0 * R.x0 + R.z0 * R.z0 - r * r;
    double sdiscr = b * b - 4.0 * 1.0 * c + kTelescEps * b * b;
    sdiscr = sqrt(sdiscr);
    return (k <= R.ir_amount / 2) ? (-b - sdiscr) / (2.0 * 1.0) : (-b + sdiscr) / (2.0 * 1.0);
  }
  const double2 v = R.rrt[ix];
  return (k <= R.ir_amount / 2) ? v.x : v.y;
}
__device__ __forceinline__ void r_elem(const Ray &R, int k, double &s, int &ix, double &r) {
  int j = (k <= R.ir_amount / 2) ? k : (R.ir_amount + 1 - k);
  ix = R.g.nr - (j - 1);
  r = (ix == 0) ? R.rstar : RCf(R.g, ix);
  s = R.staged ? R.sm_r_s[k - 1] : r_elem_tab(R, k);
}
__device__ double r_s_at(const Ray &R, int k) {
  if (k <= 0) return 0.0;  // r_s(0): read uninitialised by the reference for the outer ring; 0 by convention
  if (k > R.ir_amount) return 1.e30;
  return R.staged ? R.sm_r_s[k - 1] : r_elem_tab(R, k);
}

__device__ double theta_of_s(const Ray &R, double s) {
  double th = atan(sqrt(R.x0 * R.x0 + R.sinth02 * s * s) / (R.z0 + R.costh0 * s));
  if (th < 0.0) th = th + kPi;
  return th;
}
__device__ double radius_of_s(const Ray &R, double s) {
  return sqrt(R.x0 * R.x0 + R.z0 * R.z0 + s * s + 2.0 * R.z0 * R.costh0 * s);
}

__device__ int ir_of_radius(const GridDev &g, double rad) {  // telescope.F:3204-3214
  if (rad < RCf(g, 1)) return 0;
  if (rad > RCf(g, g.nr)) return g.nr;
  return hunt_bisect(g.rc + 2, g.nr, rad);
}

struct Extra {
  double s, radius, theta;
  int ir, it;
};

struct Emit {
  // running state of charintline's node loop
  int n;
  int ir_old, icr_old;
  double s_prev, dvmu_prev, lw_prev;
  int star_done;
};

// A node as the per-ray merge sees it (32 bytes): everything emit needs besides the ray's own constants.
// The merge (serial along a ray, one thread per ray) only writes these; the trigonometry, the stencil, the
// velocity projection and the segment bookkeeping of every node are done by node_kernel, one thread per
// node.  pk = icr | ir << 2 | it << 17.
struct __align__(16) LightNode {
  double s, radius, theta;
  uint32_t pk;
  int iray;
};

// the part of a node that does not depend on its neighbours
struct NodePoint {
  double dr, dt, lw, dvmu;
  int4 cells;
};
__device__ NodePoint node_point(const GeomParams &P, double x0, double z0, double costh0, double znew, double bnew,
                                int icr, double radius, double theta, int ir, int it, double s) {
  const GridDev &g = P.g;
  NodePoint o;
  // local direction (telescope.F:3692-3718)
  double snew = s + z0 * costh0;
  double mu = snew / sqrt(bnew * bnew + snew * snew);
  double dummy = bnew * sqrt(bnew * bnew + snew * snew - (znew + snew * costh0) * (znew + snew * costh0));
  double sinphi;
  if (dummy > 0.0) sinphi = (bnew * bnew * costh0 - znew * snew) / dummy;
  else sinphi = 1.e1 * kTelescEps;
  // telescope.F:3702-3717 forms phi = asin(sinphi) (x0 < 0) or pi - asin(sinphi), wraps it into [0, 2 pi) and
  // line.F:2659-2660 then takes sin(phi), cos(phi): sin(phi) is sinphi itself in both cases, cos(phi) = +-sqrt(1 - sinphi^2)
  // -- two roundings instead of three libm calls per node (same conditioning near |sinphi| = 1)
  const double cosphi_abs = sqrt(fmax(0.0, (1.0 - sinphi) * (1.0 + sinphi)));
  const double cosphi = (x0 < 0.0) ? cosphi_abs : -cosphi_abs;
  // position inside the cell (telescope.F:3955-3958, 4088-4091)
  double dr = (radius - RCf(g, ir)) / (RCf(g, ir + 1) - RCf(g, ir));
  double dt = (theta - TCf(g, it)) / (TCf(g, it + 1) - TCf(g, it));
  if (dr < 0.0 || dr > 1.0) atomicCAS(P.status, 0, 6024);
  if (dt < 0.0 || dt > 1.0) atomicCAS(P.status, 0, 6023);
  // stencil (line.F:4006-4039)
  int r0 = ir, r1, t0 = it, t1;
  if (dr > 0.0) { r1 = ir + 1; if (r1 > g.nr) r1 = g.nr; }
  else { r1 = ir - 1; if (r1 < 1) r1 = 1; }
  t1 = (dt > 0.0) ? it + 1 : it - 1;
  t0 = RIDXf(g, t0);
  t1 = RIDXf(g, t1);
  r0 = max(1, min(g.nr, r0));  // (memory safety only; ir is in 1..nr for accepted nodes)
  int4 cells;
  cells.x = (r0 - 1) * g.nth + (t0 - 1);
  cells.y = (r0 - 1) * g.nth + (t1 - 1);
  cells.z = (r1 - 1) * g.nth + (t0 - 1);
  cells.w = (r1 - 1) * g.nth + (t1 - 1);
  // line width and velocity (line.F:4065-4067, 4090-4094 and the icr=2,3 twins)
  double lw, v1, v2, v3;
  {
    double4 a = P.cellS[cells.x];
    if (icr == 1) {
      double4 b = P.cellS[cells.y];
      lw = (1.0 - dt) * a.x + dt * b.x;
      v1 = (1.0 - dt) * a.y + dt * b.y;
      v2 = (1.0 - dt) * a.z + dt * b.z;
      v3 = (1.0 - dt) * a.w + dt * b.w;
    } else if (icr == 2) {
      double4 b = P.cellS[cells.z];
      lw = (1.0 - dr) * a.x + dr * b.x;
      v1 = (1.0 - dr) * a.y + dr * b.y;
      v2 = (1.0 - dr) * a.z + dr * b.z;
      v3 = (1.0 - dr) * a.w + dr * b.w;
    } else {
      double4 b = P.cellS[cells.y], c = P.cellS[cells.z], d = P.cellS[cells.w];
      lw = (1.0 - dr) * ((1.0 - dt) * a.x + dt * b.x) + dr * ((1.0 - dt) * c.x + dt * d.x);
      v1 = (1.0 - dr) * ((1.0 - dt) * a.y + dt * b.y) + dr * ((1.0 - dt) * c.y + dt * d.y);
      v2 = (1.0 - dr) * ((1.0 - dt) * a.z + dt * b.z) + dr * ((1.0 - dt) * c.z + dt * d.z);
      v3 = (1.0 - dr) * ((1.0 - dt) * a.w + dt * b.w) + dr * ((1.0 - dt) * c.w + dt * d.w);
    }
  }
  if (mu > 1.0) atomicCAS(P.status, 0, 393);  // line.F:2656
  o.dvmu = 3.335668e-11 * (mu * v1 + sqrt(1.0 - mu * mu) * (v2 * sinphi + v3 * cosphi));
  o.dr = dr;
  o.dt = dt;
  o.lw = lw;
  o.cells = cells;
  return o;
}

// the record of a node from its own point values and the previous node of the ray (have_prev = false:
// first node); segment bookkeeping of telescope.F:4050-4104, 4129-4193, sub-grid trigger of line.F:4706-4709
__device__ NodeRec node_record(const GeomParams &P, int iray, const NodePoint &pt, int icr, int ir, double s,
                               bool have_prev, double s_prev, int ir_old, int icr_old, double dvmu_prev,
                               double lw_prev, int &star_done, bool star_hit) {
  uint32_t flag = (uint32_t)icr;
  double ds = 0.0;
  if (have_prev) {
    ds = s - s_prev;
    if (ds < 0.0) atomicCAS(P.status, 0, 749);
    if (ir == 1 && ir_old == 1 && icr == 1 && icr_old == 1) { ds = 0.0; flag |= kFlagInit; }
    if (!(ir > 1)) {
      if (P.in_itype == 1) {
        if (ir_old == 1 && icr_old == 1) flag |= kFlagZero | kFlagInit;
      } else if (P.in_itype == 2) {
        if (ir == 1 && !star_done && iray == 0 && P.rbeam0_center > 0.0) {
          if (P.rbeam0_center < P.rstar) atomicCAS(P.status, 0, 124);
          flag |= kFlagStar | kFlagInit;
          star_done = 1;
        }
        // rbeam0 = 0 (rectangular camera): a ray that re-emerges from the central hole with an impact
        // parameter inside the star has hit it (telescope.F:4194-4208); the render uses star fraction 1
        if (P.rect && ir == 1 && ir_old == 1 && star_hit) flag |= kFlagStar | kFlagInit;
      } else {
        atomicCAS(P.status, 0, 13);  // inner BC 0 disabled / unknown (telescope.F:4129-4134, 4210)
      }
    }
  }
  double inv_lwav = 1.0 / pt.lw;
  if (have_prev) {
    const double lwseg = 0.5 * (lw_prev + pt.lw);
    const double q = fabs((pt.dvmu - dvmu_prev) / (lwseg / 2.99792458e5));
    if (2.0 * 3.0 * q > 1.0) flag |= kFlagSub;
    inv_lwav = 1.0 / lwseg;
  }
  NodeRec rec;
  rec.ds = ds;
  rec.dvmu = pt.dvmu;
  rec.lw = pt.lw;
  rec.inv_lwav = inv_lwav;
  rec.wr = pt.dr;
  rec.wt = pt.dt;
  int4 cells = pt.cells;
  cells.x |= (int)(flag << kCellFlagShift);
  rec.cells = cells;
  return rec;
}

// COUNT: count only.  Otherwise the ray leaves its LightNode list; node_kernel finishes every node.
template <bool COUNT>
__device__ void emit_node(const GeomParams &P, const Ray &R, Emit &E, long long base, int iray,
                          int icr, double radius, double theta, int ir, int it, double s) {
  if (COUNT) {
    E.n++;
    return;
  }
  LightNode ln;
  ln.s = s;
  ln.radius = radius;
  ln.theta = theta;
  ln.pk = (uint32_t)icr | ((uint32_t)ir << 2) | ((uint32_t)it << 17);
  ln.iray = iray;
  static_cast<LightNode *>(P.light)[base + E.n] = ln;
  E.n++;
}

__global__ void __launch_bounds__(256) node_kernel(GeomParams P, long long ntot) {
  // what the next node of the ray needs from this one travels through shared memory (the threads of a
  // block hold consecutive nodes); only thread 0 evaluates its predecessor a second time
  __shared__ double s_dvmu[256], s_lw[256], s_s[256];
  __shared__ int s_ir[256], s_icr[256];
  const int t = threadIdx.x;
  const long long i = (long long)blockIdx.x * blockDim.x + t;
  const bool active = i < ntot;
  const LightNode *light = static_cast<const LightNode *>(P.light);
  LightNode me;
  me.s = me.radius = me.theta = 0.0;
  me.pk = 0;
  me.iray = 1;
  if (active) me = light[i];
  const int iray = me.iray;
  const double x0 = P.x0[iray], z0 = P.z0[iray];
  const double costh0 = P.costh0, sinth0 = P.sinth0;
  const double sinth02 = sinth0 * sinth0;
  const double znew = z0 * sinth02;
  const double bnew = sqrt(x0 * x0 + z0 * z0 * sinth02);
  // theta cell of an R crossing (telescope.F:3331-3344) / radial cell of a theta crossing (:3204-3214)
  auto complete = [&](const LightNode &n, int &icr, int &ir, int &it, double &theta) {
    icr = (int)(n.pk & 3u);
    ir = (int)((n.pk >> 2) & 0x7fffu);
    it = (int)(n.pk >> 17);
    theta = n.theta;
    if (icr == 1) {
      theta = atan(sqrt(x0 * x0 + sinth02 * n.s * n.s) / (z0 + costh0 * n.s));
      if (theta < 0.0) theta = theta + kPi;
      it = hunt_bisect(P.g.tc + 2, P.g.nt, theta);
    } else if (icr == 2) {
      ir = ir_of_radius(P.g, n.radius);
    }
  };
  int icr = 0, ir = 0, it = 0;
  double theta = 0.0;
  NodePoint pt;
  pt.dr = pt.dt = pt.lw = pt.dvmu = 0.0;
  pt.cells = make_int4(0, 0, 0, 0);
  if (active) {
    complete(me, icr, ir, it, theta);
    pt = node_point(P, x0, z0, costh0, znew, bnew, icr, me.radius, theta, ir, it, me.s);
  }
  s_dvmu[t] = pt.dvmu;
  s_lw[t] = pt.lw;
  s_s[t] = me.s;
  s_ir[t] = ir;
  s_icr[t] = icr;
  __syncthreads();
  if (!active) return;
  const bool have_prev = i != P.node_off[iray];
  double s_prev = 0.0, dvmu_prev = 0.0, lw_prev = 0.0;
  int ir_old = -99, icr_old = -99;
  if (have_prev) {
    if (t > 0) {  // node i-1 is on the same ray: thread t-1 evaluated it with the same ray constants
      s_prev = s_s[t - 1];
      dvmu_prev = s_dvmu[t - 1];
      lw_prev = s_lw[t - 1];
      ir_old = s_ir[t - 1];
      icr_old = s_icr[t - 1];
    } else {
      const LightNode pv = light[i - 1];
      int it_old;
      double theta_old;
      complete(pv, icr_old, ir_old, it_old, theta_old);
      const NodePoint pp = node_point(P, x0, z0, costh0, znew, bnew, icr_old, pv.radius, theta_old, ir_old,
                                      it_old, pv.s);
      s_prev = pv.s;
      dvmu_prev = pp.dvmu;
      lw_prev = pp.lw;
    }
  }
  // The circular camera mixes the star into the centre ray, once, at its first node on the innermost sphere
  // (telescope.F:4148-4193: istar_done): has an earlier node of the ray been there already?
  int star_done = 1;
  if (iray == 0 && !P.rect && ir == 1 && have_prev && P.in_itype == 2 && P.rbeam0_center > 0.0) {
    star_done = 0;
    for (long long j = P.node_off[0] + 1; j < i && !star_done; j++) {
      const LightNode o = light[j];
      const int kind = (int)(o.pk & 3u);
      const int ir_j = (kind == 2) ? ir_of_radius(P.g, o.radius) : (int)((o.pk >> 2) & 0x7fffu);
      if (ir_j == 1) star_done = 1;
    }
  }
  const bool star_hit = sqrt(x0 * x0 + z0 * z0 * (1.0 - costh0 * costh0)) <= P.rstar;  // tr_b (telescope.F:3268)
  P.nodes.rec[i] = node_record(P, iray, pt, icr, ir, me.s, have_prev, s_prev, ir_old, icr_old, dvmu_prev, lw_prev,
                               star_done, star_hit);
}

// One thread per (ray of the block, surface): the roots of the ray with every cone of the stored hemisphere
// (telescope.F:2960-2990, with the radius of each root: :3204) and with every sphere ix = 0 (the star) .. nr
// (telescope.F:3291-3312, discriminant padded by 1e-10 b^2) -- the same expressions, in the same order, as the
// per-ray code evaluated them, now once per (ray, surface) and in parallel.
__global__ void __launch_bounds__(256) roots_kernel(GeomParams P) {
  const GridDev &g = P.g;
  const int nth = g.nt / 2, per = nth + g.nr + 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nr_rays = (long long)(P.ray_hi - P.ray_lo + 1);
  if (i >= nr_rays * per) return;
  const int rr = (int)(i / per), j = (int)(i - (long long)rr * per);
  const int iray = P.ray_lo + rr;
  if (P.rect && iray > 0 && !(P.rb[iray] < P.bskip)) return;
  const double x0 = P.x0[iray], z0 = P.z0[iray];
  const double costh0 = P.costh0, sinth0 = P.sinth0;
  const double costh02 = costh0 * costh0, sinth02 = sinth0 * sinth0;
  if (j < nth) {
    double a1, a2;
    double4 v;
    if (th_roots_math(x0, z0, costh0, costh02, sinth02, P.tan2[j + 1], a1, a2)) {
      v.x = a1;
      v.y = a2;
      v.z = sqrt(x0 * x0 + z0 * z0 + a1 * a1 + 2.0 * z0 * costh0 * a1);
      v.w = sqrt(x0 * x0 + z0 * z0 + a2 * a2 + 2.0 * z0 * costh0 * a2);
    } else {
      v.x = v.y = v.z = v.w = __longlong_as_double(0x7ff8000000000000LL);
    }
    P.thr[(size_t)rr * nth + j] = v;
  } else {
    const int ix = j - nth;
    const double r = (ix == 0) ? P.rstar : RCf(g, ix);
    const double b = 2.0 * costh0 * z0;
    const double c = x0 * x0 + z0 * z0 - r * r;
    double sdiscr = b * b - 4.0 * 1.0 * c + kTelescEps * b * b;
    sdiscr = sqrt(sdiscr);  // (negative only for spheres inside the impact parameter, which are never read)
    P.rrt[(size_t)rr * (g.nr + 1) + ix] = make_double2((-b - sdiscr) / (2.0 * 1.0), (-b + sdiscr) / (2.0 * 1.0));
  }
}

// tan(theta_iy)^2 of the stored hemisphere's cones (telescope.F:2960-2962 evaluates it per ray and cone)
__global__ void tan2_kernel(GridDev g, double *tan2) {
  const int iy = blockIdx.x * blockDim.x + threadIdx.x;
  if (iy < 1 || iy > g.nt / 2) return;
  const double t = tan(TCf(g, iy));
  tan2[iy] = t * t;
}

// ------------------------------------------------------------------------------------------
// The merge of a ray done by all 32 lanes of its warp (every ray but the centre one).  The position of every
// element in the reference's 3-way merge (telescope.F:3607-3677: extras while strictly smaller than both
// stream heads; on a tie between the streams the R crossing goes first) follows from ranks alone:
//   theta crossing k :  (k-1) + #R{r_s <= s} + #extras{s_e < s}
//   R crossing k     :  (k-1) + #theta{th_s < s} + #extras{s_e < s}
//   extra j (sorted) :   j    + #theta{th_s <= s} + #R{r_s <= s}      (dropped if no crossing follows it)
// so each lane places its share of the elements into a shared-memory array with binary searches over the staged
// streams.  The filters (radius windows; distance to the previous ACCEPTED node > 1e-10 R, :3352-3366) run over
// contiguous chunks of that array, the "previous accepted" carried across lanes; they are exact as long as no
// node is rejected for being too close to its predecessor, which the warp detects -- such a ray (a degenerate
// coincidence of two surfaces) is redone serially by lane 0 over the same array.
// ------------------------------------------------------------------------------------------
struct MergeSm {
  int *pk;  // merged order: kind (1 R crossing, 2 theta crossing, 3 extra) | index into the source << 2
  double *ex_s, *ex_rad, *ex_theta, *ex_sorted;
  int *ex_ir, *ex_it, *ex_rank;
};
template <bool COUNT>
__device__ void geom_parallel_tail(const GeomParams &P, const Ray &R, int iray, int lane, double bimpact,
                                   const MergeSm &M) {
  const GridDev &g = R.g;
  const int nr = g.nr, nt = g.nt;
  const double x0 = R.x0, z0 = R.z0;
  const double eps = kTelescEps, epsplus = 1.e1 * kTelescEps;
  const double rmaxt = RCf(g, nr) * (1.0 - epsplus), rmaxr = RCf(g, nr) * (1.0 + epsplus);
  const double rmint = RCf(g, 1) * (1.0 + epsplus), rminr = RCf(g, 1) * (1.0 - epsplus);
  const int nth_s = R.ith_amount, nr_s = R.ir_amount, total = nth_s + nr_s;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  // ranks of the streams at a path length
  auto th_lt = [&](double v) {  // #theta crossings with s < v
    int lo = 0, hi = nth_s;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (R.sm_th_s[m] < v) lo = m + 1; else hi = m; }
    return lo;
  };
  auto th_le = [&](double v) {
    int lo = 0, hi = nth_s;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (R.sm_th_s[m] <= v) lo = m + 1; else hi = m; }
    return lo;
  };
  auto r_le = [&](double v) {
    int lo = 0, hi = nr_s;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (R.sm_r_s[m] <= v) lo = m + 1; else hi = m; }
    return lo;
  };
  // ---- extremum sampling points (telescope.F:3417-3600): slot e = 0 radial extremum, 1..16 its sub-points,
  // 17 theta extremum, 18..33 its sub-points; one lane per slot ----
  const bool fam_r = bimpact > RCf(g, 1);
  const double s3r = 0.0 - z0 * R.costh0;
  int isrt_r = 0;
  if (fam_r) {
    int jlo = 0, jhi = nr_s + 1;
    while (jhi - jlo != 1) {
      int jm = (jhi + jlo) >> 1;
      if (s3r > r_s_at(R, jm)) jlo = jm;
      else jhi = jm;
    }
    isrt_r = jlo;
  }
  const double s3t = x0 * x0 * R.costh0 / (z0 * R.sinth02);
  const double rr_t = radius_of_s(R, s3t);
  const bool fam_t = rr_t > RCf(g, 1) && rr_t < RCf(g, nr);
  int isrt_t = 0;
  bool sub_t = false;
  if (fam_t) {
    int jlo = 0, jhi = nth_s + 1;
    while (jhi - jlo != 1) {
      int jm = (jhi + jlo) >> 1;
      if (s3t > th_s_at(R, jm)) jlo = jm;
      else jhi = jm;
    }
    isrt_t = jlo;
    sub_t = !((isrt_t - kRayRnpt + 1 < 1) || (isrt_t + kRayRnpt > nth_s));
  }
  for (int e = lane; e < 34; e += 32) {
    double es = nan, erad = 0.0, eth = 0.0;
    int eir = 0, eit = 0;
    bool valid = false;
    if (e == 0) {
      if (fam_r) {
        es = s3r;
        erad = bimpact;
        eth = theta_of_s(R, s3r);
        eir = R.ir_min - 1;
        eit = hunt_bisect(g.tc + 2, nt, eth);
        valid = true;
      }
    } else if (e <= 16) {
      if (fam_r) {
        const int irng = 1 + (e - 1) / 4, iad = 1 + (e - 1) % 4;
        double s1, s2;
        if (irng == 1) { s1 = s3r; s2 = r_s_at(R, isrt_r + 1); }
        else if (irng == 2) { s1 = r_s_at(R, isrt_r); s2 = s3r; }
        else if (irng == 3) { s1 = r_s_at(R, isrt_r + 1); s2 = r_s_at(R, isrt_r + 2); }
        else { s1 = r_s_at(R, isrt_r - 1); s2 = r_s_at(R, isrt_r); }
        const double ds = (s2 - s1) / (1.0 + 1.0 * kRayAdpt);
        es = iad * ds + s1;
        erad = radius_of_s(R, es);
        eth = theta_of_s(R, es);
        eir = hunt_bisect(g.rc + 2, nr, erad);
        eit = hunt_bisect(g.tc + 2, nt, eth);
        valid = true;
      }
    } else if (e == 17) {
      if (fam_t) {
        es = s3t;
        erad = rr_t;
        eth = theta_of_s(R, s3t);
        eir = hunt_bisect(g.rc + 2, nr, rr_t);
        if (eir == 0 || eir == nr) atomicCAS(P.status, 0, 192);
        eit = hunt_bisect(g.tc + 2, nt, eth);
        valid = true;
      }
    } else if (fam_t && sub_t) {
      const int irng = 1 + (e - 18) / 4, iad = 1 + (e - 18) % 4;
      double s1, s2;
      if (irng == 1) { s1 = s3t; s2 = th_s_at(R, isrt_t + 1); }
      else if (irng == 2) { s1 = th_s_at(R, isrt_t); s2 = s3t; }
      else if (irng == 3) { s1 = th_s_at(R, isrt_t + 1); s2 = th_s_at(R, isrt_t + 2); }
      else { s1 = th_s_at(R, isrt_t - 1); s2 = th_s_at(R, isrt_t); }
      if (s2 == s1) atomicCAS(P.status, 0, 987);
      if (s2 < s1) atomicCAS(P.status, 0, 988);
      const double ds = (s2 - s1) / (1.0 + 1.0 * kRayAdpt);
      const double s0 = iad * ds + s1;
      const double rad = radius_of_s(R, s0);
      const int ixx = hunt_bisect(g.rc + 2, nr, rad);
      if (!(ixx == 0 || ixx == nr)) {
        es = s0;
        erad = rad;
        eth = theta_of_s(R, s0);
        eir = ixx;
        eit = hunt_bisect(g.tc + 2, nt, eth);
        valid = true;
      }
    }
    // an extra that no crossing follows is never emitted by the reference's merge loop (telescope.F:3607-3677)
    if (valid && !(th_le(es) + r_le(es) < total)) valid = false;
    M.ex_s[e] = valid ? es : nan;
    M.ex_rad[e] = erad;
    M.ex_theta[e] = eth;
    M.ex_ir[e] = eir;
    M.ex_it[e] = eit;
  }
  __syncwarp();
  // stable order of the emitted extras by s (nrecip.F:833 ray_sort as an insertion sort keeps equal keys in
  // generation order)
  for (int e = lane; e < 34; e += 32) {
    const double v = M.ex_s[e];
    int rank = -1;
    if (v == v) {
      rank = 0;
      for (int j = 0; j < 34; j++) {
        const double w = M.ex_s[j];
        if (w < v || (w == v && j < e)) rank++;  // (NaN compares false)
      }
      M.ex_sorted[rank] = v;
    }
    M.ex_rank[e] = rank;
  }
  __syncwarp();
  int nex = 0;
  for (int j = 0; j < 34; j++) nex += M.ex_rank[j] >= 0;
  auto ex_lt = [&](double v) {  // #emitted extras with s < v
    int lo = 0, hi = nex;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (M.ex_sorted[m] < v) lo = m + 1; else hi = m; }
    return lo;
  };
  // ---- place every element ----
  const int Mtot = total + nex;
  for (int k = lane; k < nth_s; k += 32) {
    const double v = R.sm_th_s[k];
    M.pk[k + r_le(v) + ex_lt(v)] = 2 | (k << 2);
  }
  for (int k = lane; k < nr_s; k += 32) {
    const double v = R.sm_r_s[k];
    M.pk[k + th_lt(v) + ex_lt(v)] = 1 | (k << 2);
  }
  for (int e = lane; e < 34; e += 32) {
    if (M.ex_rank[e] < 0) continue;
    const double v = M.ex_s[e];
    M.pk[M.ex_rank[e] + th_le(v) + r_le(v)] = 3 | (e << 2);
  }
  __syncwarp();
  // an element of the merged order from its source
  auto r_index = [&](int k) {  // sphere of R crossing k (0-based): telescope.F:3284-3335
    const int j = (k + 1 <= nr_s / 2) ? k + 1 : (nr_s + 1 - (k + 1));
    return nr - (j - 1);
  };
  auto elem = [&](int i, double &sv, double &rad, int &kind, int &idx) {
    const int p = M.pk[i];
    kind = p & 3;
    idx = p >> 2;
    if (kind == 2) {
      sv = R.sm_th_s[idx];
      rad = R.sm_th_rad[idx];
    } else if (kind == 1) {
      sv = R.sm_r_s[idx];
      const int ix = r_index(idx);
      rad = (ix == 0) ? R.rstar : RCf(g, ix);
    } else {
      sv = M.ex_s[idx];
      rad = M.ex_rad[idx];
    }
  };
  auto window_ok = [&](int kind, double rad) {
    return (kind == 1) ? (rad <= rmaxr && rad >= rminr) : (rad <= rmaxt && rad >= rmint);
  };
  auto light_of = [&](int kind, int idx, double sv, double rad) {
    LightNode ln;
    ln.s = sv;
    ln.radius = rad;
    ln.iray = iray;
    if (kind == 2) {
      ln.theta = TCf(g, R.sm_th_it[idx]);
      ln.pk = 2u | ((uint32_t)R.sm_th_it[idx] << 17);
    } else if (kind == 1) {
      ln.theta = 0.0;
      ln.pk = 1u | ((uint32_t)r_index(idx) << 2);
    } else {
      ln.theta = M.ex_theta[idx];
      ln.pk = 3u | ((uint32_t)M.ex_ir[idx] << 2) | ((uint32_t)M.ex_it[idx] << 17);
    }
    return ln;
  };
  // ---- filters over contiguous chunks ----
  const int per = (Mtot + 31) / 32, c0 = min(Mtot, lane * per), c1 = min(Mtot, c0 + per);
  double last_ok = -1.e30;  // path length of the last node of my chunk that passes its window
  bool have = false;
  int nacc = 0;
  bool close_reject = false;
  // first walk: with the previous window-passing node taken from inside the chunk only
  double first_ok_s = 0.0, first_ok_rad = 0.0;
  bool first_seen = false;
  {
    double sp = 0.0;
    for (int i = c0; i < c1; i++) {
      double sv, rad;
      int kind, idx;
      elem(i, sv, rad, kind, idx);
      if (!window_ok(kind, rad)) continue;
      if (!first_seen) {
        first_seen = true;
        first_ok_s = sv;
        first_ok_rad = rad;
        nacc++;  // (judged against the previous chunks below)
      } else if ((sv - sp) > eps * rad) {
        nacc++;
      } else {
        close_reject = true;
      }
      sp = sv;  // exact while nothing is rejected for closeness (checked below)
      last_ok = sv;
      have = true;
    }
  }
  // previous window-passing node before my chunk: the nearest lane below that has one
  double sprev = -1.e30;
  for (int src = 0; src < 32; src++) {
    const double v = __shfl_sync(0xffffffffu, last_ok, src);
    const bool h = __shfl_sync(0xffffffffu, have ? 1 : 0, src) != 0;
    if (src < lane && h) sprev = v;
  }
  if (first_seen && !((first_ok_s - sprev) > eps * first_ok_rad)) close_reject = true;
  const long long base = COUNT ? 0 : P.node_off[iray];
  LightNode *out = COUNT ? nullptr : static_cast<LightNode *>(P.light) + base;
  if (__any_sync(0xffffffffu, close_reject)) {
    // degenerate ray: lane 0 applies the reference's sequential rule over the merged order
    if (lane == 0) {
      int n = 0;
      double sp = -1.e30;
      for (int i = 0; i < Mtot; i++) {
        double sv, rad;
        int kind, idx;
        elem(i, sv, rad, kind, idx);
        if (!window_ok(kind, rad) || !((sv - sp) > eps * rad)) continue;
        if (!COUNT) out[n] = light_of(kind, idx, sv, rad);
        sp = sv;
        n++;
      }
      if (COUNT) P.node_cnt[iray] = n;
    }
    return;
  }
  int off = nacc;  // exclusive prefix of the accepted counts
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, off, o);
    if (lane >= o) off += v;
  }
  const int tot_acc = __shfl_sync(0xffffffffu, off, 31);
  off -= nacc;
  if (COUNT) {
    if (lane == 0) P.node_cnt[iray] = tot_acc;
    return;
  }
  for (int i = c0; i < c1; i++) {
    double sv, rad;
    int kind, idx;
    elem(i, sv, rad, kind, idx);
    if (!window_ok(kind, rad)) continue;
    out[off++] = light_of(kind, idx, sv, rad);
  }
}

// One warp per ray.  The lanes first stage the ray's two crossing streams in path order in shared memory -- theta
// crossings (telescope.F:3024-3194 ordering) and R crossings (:3284-3335), each element an independent table
// read -- then lane 0 runs the extremum sampling, the 3-way merge and the filters of telescope.F:3352-3677 over
// them: what is serial along a ray now touches shared memory only.
// WARP = false is the same code with one thread per ray and the streams read from the tables: many rays per
// warp in lockstep, which wins when a render has enough rays to fill the GPU that way (the merge is a few
// hundred dependent instructions per node: one active lane per warp issues them at a small fraction of the
// rate 32 do); WARP = true scales down to the few thousand rays of one rank's ring block.
template <bool COUNT, bool WARP>
__global__ void __launch_bounds__(WARP ? 32 : 128) geom_kernel(GeomParams P) {
  extern __shared__ double geom_sm[];
  const int iray = P.ray_lo + (WARP ? (int)blockIdx.x : (int)(blockIdx.x * blockDim.x + threadIdx.x));
  const int lane = WARP ? (int)threadIdx.x : 0;
  if (iray >= P.nray || iray > P.ray_hi) return;
  const GridDev &g = P.g;
  const int nr = g.nr, nt = g.nt, nth = nt / 2;
  if (P.rect && iray > 0 && !(P.rb[iray] < P.bskip)) {  // a pixel off the model
    if (COUNT && lane == 0) P.node_cnt[iray] = 0;
    return;
  }
  double *sm_th_s = geom_sm, *sm_th_rad = sm_th_s + 2 * nth, *sm_r_s = sm_th_rad + 2 * nth;
  int *sm_th_it = reinterpret_cast<int *>(sm_r_s + 2 * (nr + 1));
  Ray R;
  R.g = g;
  R.tan2 = P.tan2;
  R.thr = P.thr + (size_t)(iray - P.ray_lo) * (size_t)nth;
  R.rrt = P.rrt + (size_t)(iray - P.ray_lo) * (size_t)(nr + 1);
  R.sm_th_s = sm_th_s;
  R.sm_th_rad = sm_th_rad;
  R.sm_r_s = sm_r_s;
  R.sm_th_it = sm_th_it;
  R.staged = WARP;
  R.x0 = P.x0[iray];
  R.z0 = P.z0[iray];
  R.rstar = P.rstar;
  const double theta0 = P.theta0;
  R.pitheta0 = 0.5 * kPi - theta0;
  R.costh0 = P.costh0;
  R.sinth0 = P.sinth0;
  R.costh02 = R.costh0 * R.costh0;
  R.sinth02 = R.sinth0 * R.sinth0;
  const double x0 = R.x0, z0 = R.z0;
  const double eps = kTelescEps, epsplus = 1.e1 * kTelescEps;

  // ---- theta crossings: which cones are hit (telescope.F:2958-3014) ----
  {
    int cnt = 0, first = 0x7fffffff, dbl = 0;
    for (int iy = 1 + lane; iy <= nth; iy += (WARP ? 32 : 1)) {
      double a1, a2;
      if (th_roots(R, iy, a1, a2)) {
        first = min(first, iy);
        cnt++;
        double pitheta = 0.5 * kPi - TCf(g, iy);
        if (fabs(pitheta) > fabs(R.pitheta0)) dbl++;
      }
    }
    if (WARP) {
      for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        dbl += __shfl_xor_sync(0xffffffffu, dbl, o);
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      }
    }
    if (cnt == 0) first = 0;
    // hit cones form a contiguous tail iy_first..nth (disc>0 <=> tan^2 theta above a threshold)
    if (cnt > 0 && first + cnt - 1 != nth) atomicCAS(P.status, 0, 9001);
    R.isnr = cnt;
    R.isdbl = dbl;
    R.iy_first = first;
    R.iup = (R.pitheta0 > 0.0) ? 1 : 0;
    R.branchA = (z0 * R.pitheta0 > 0.0) ? 1 : 0;
    R.ith_amount = 2 * cnt;
  }
  // ---- R crossings (telescope.F:3241-3276) ----
  const double bimpact = sqrt(x0 * x0 + z0 * z0 * (1.0 - R.costh02));
  {
    // first ix with R_ix > bimpact (telescope.F:3246-3252): R ascending, so bisect
    int lo = 0, hi = nr + 1;  // R_lo <= bimpact < R_hi with R_0 = -inf, R_{nr+1} = +inf
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (RCf(g, mid) > bimpact) hi = mid;
      else lo = mid;
    }
    const int ix = hi;
    if (ix > nr) {
      atomicCAS(P.status, 0, 13);
      if (COUNT && lane == 0) P.node_cnt[iray] = 0;
      return;
    }
    R.ir_min = ix;
    if (bimpact <= P.rstar) R.ir_min = 0;
    R.ir_amount = 2 * (nr + 1 - R.ir_min);
  }
  if (WARP) {
    // ---- stage both streams (every element an independent table read) ----
    for (int k = 1 + lane; k <= R.ith_amount; k += 32) {
      double s, rad;
      int it;
      th_elem_tab(R, k, s, it, rad);
      sm_th_s[k - 1] = s;
      sm_th_rad[k - 1] = rad;
      sm_th_it[k - 1] = it;
    }
    for (int k = 1 + lane; k <= R.ir_amount; k += 32) sm_r_s[k - 1] = r_elem_tab(R, k);
    __syncwarp();
    {  // the merge by all lanes
      const int Mmax = 2 * nth + 2 * (nr + 1) + 34;
      MergeSm M;
      M.ex_s = reinterpret_cast<double *>(sm_th_it + 2 * nth);
      M.ex_rad = M.ex_s + 34;
      M.ex_theta = M.ex_rad + 34;
      M.ex_sorted = M.ex_theta + 34;
      M.pk = reinterpret_cast<int *>(M.ex_sorted + 34);
      M.ex_ir = M.pk + Mmax;
      M.ex_it = M.ex_ir + 34;
      M.ex_rank = M.ex_it + 34;
      geom_parallel_tail<COUNT>(P, R, iray, lane, bimpact, M);
      return;
    }
  }
  const double rmaxt = RCf(g, nr) * (1.0 - epsplus);
  const double rmaxr = RCf(g, nr) * (1.0 + epsplus);
  const double rmint = RCf(g, 1) * (1.0 + epsplus);
  const double rminr = RCf(g, 1) * (1.0 - epsplus);
  const double send = 1.e30, sbeg = -1.e30;

  // ---- extremum sampling points (telescope.F:3417-3600) ----
  Extra ex[kMaxExtra];
  int nex = 0;
  if (bimpact > RCf(g, 1)) {
    double s2 = 0.0 - z0 * R.costh0;
    ex[nex].s = s2;
    ex[nex].radius = bimpact;
    ex[nex].theta = theta_of_s(R, s2);
    ex[nex].ir = R.ir_min - 1;
    ex[nex].it = hunt_bisect(g.tc + 2, nt, ex[nex].theta);
    nex++;
    const double s3 = s2;
    // hunt0(r_s, ir_amount, s2): largest k with r_s(k) < s2
    int isrt;
    {
      int jlo = 0, jhi = R.ir_amount + 1;
      while (jhi - jlo != 1) {
        int jm = (jhi + jlo) >> 1;
        if (s2 > r_s_at(R, jm)) jlo = jm;
        else jhi = jm;
      }
      isrt = jlo;
    }
    for (int irng = 1; irng <= 4; irng++) {
      double s1;
      if (irng == 1) { s1 = s3; s2 = r_s_at(R, isrt + 1); }
      else if (irng == 2) { s1 = r_s_at(R, isrt); s2 = s3; }
      else if (irng == 3) { s1 = r_s_at(R, isrt + 1); s2 = r_s_at(R, isrt + 2); }
      else { s1 = r_s_at(R, isrt - 1); s2 = r_s_at(R, isrt); }
      double ds = (s2 - s1) / (1.0 + 1.0 * kRayAdpt);
      for (int iad = 1; iad <= kRayAdpt; iad++) {
        double s0 = iad * ds + s1;
        ex[nex].s = s0;
        ex[nex].radius = radius_of_s(R, s0);
        ex[nex].theta = theta_of_s(R, s0);
        ex[nex].ir = hunt_bisect(g.rc + 2, nr, ex[nex].radius);
        ex[nex].it = hunt_bisect(g.tc + 2, nt, ex[nex].theta);
        nex++;
      }
    }
  }
  {
    double s2 = x0 * x0 * R.costh0 / (z0 * R.sinth02);  // NaN for the centre ray: falls through
    double rr = radius_of_s(R, s2);
    if (rr > RCf(g, 1) && rr < RCf(g, nr)) {
      ex[nex].s = s2;
      ex[nex].radius = rr;
      ex[nex].theta = theta_of_s(R, s2);
      int ix = hunt_bisect(g.rc + 2, nr, rr);
      if (ix == 0 || ix == nr) atomicCAS(P.status, 0, 192);
      ex[nex].ir = ix;
      ex[nex].it = hunt_bisect(g.tc + 2, nt, ex[nex].theta);
      nex++;
      const double s3 = s2;
      int isrt;
      {
        int jlo = 0, jhi = R.ith_amount + 1;
        while (jhi - jlo != 1) {
          int jm = (jhi + jlo) >> 1;
          if (s3 > th_s_at(R, jm)) jlo = jm;
          else jhi = jm;
        }
        isrt = jlo;
      }
      if (!((isrt - kRayRnpt + 1 < 1) || (isrt + kRayRnpt > R.ith_amount))) {
        for (int irng = 1; irng <= 4; irng++) {
          double s1;
          if (irng == 1) { s1 = s3; s2 = th_s_at(R, isrt + 1); }
          else if (irng == 2) { s1 = th_s_at(R, isrt); s2 = s3; }
          else if (irng == 3) { s1 = th_s_at(R, isrt + 1); s2 = th_s_at(R, isrt + 2); }
          else { s1 = th_s_at(R, isrt - 1); s2 = th_s_at(R, isrt); }
          if (s2 == s1) atomicCAS(P.status, 0, 987);
          if (s2 < s1) atomicCAS(P.status, 0, 988);
          double ds = (s2 - s1) / (1.0 + 1.0 * kRayAdpt);
          for (int iad = 1; iad <= kRayAdpt; iad++) {
            double s0 = iad * ds + s1;
            double rad = radius_of_s(R, s0);
            int ixx = hunt_bisect(g.rc + 2, nr, rad);
            if (ixx == 0 || ixx == nr) continue;
            ex[nex].s = s0;
            ex[nex].radius = rad;
            ex[nex].theta = theta_of_s(R, s0);
            ex[nex].ir = ixx;
            ex[nex].it = hunt_bisect(g.tc + 2, nt, ex[nex].theta);
            nex++;
          }
        }
      }
    }
  }
  // sort the extra points by s (nrecip.F:833 ray_sort); insertion sort, n <= 34
  for (int i = 1; i < nex; i++) {
    Extra e = ex[i];
    int j = i - 1;
    while (j >= 0 && ex[j].s > e.s) {
      ex[j + 1] = ex[j];
      j--;
    }
    ex[j + 1] = e;
  }

  // ---- 3-way merge by s (telescope.F:3607-3677) ----
  const long long base = COUNT ? 0 : P.node_off[iray];
  // (the bracketing searches of a node -- theta cell of an R crossing, radial cell of a theta crossing -- and
  // everything else per node are node_kernel's)
  Emit E;
  E.n = 0;
  E.ir_old = -99;
  E.icr_old = -99;
  E.s_prev = 0.0;
  E.dvmu_prev = 0.0;
  E.lw_prev = 0.0;
  E.star_done = 0;
  double sprev = -1.e30;
  int ist = 1, isr = 1, isex = 0;
  double th_s = 1.e30, r_s = 1.e30, r_rad = 0.0, th_rad = 0.0;
  int th_it = 0, r_ix = 0;
  if (ist <= R.ith_amount) th_elem(R, ist, th_s, th_it, th_rad);
  if (isr <= R.ir_amount) r_elem(R, isr, r_s, r_ix, r_rad);
  const int total = R.ir_amount + R.ith_amount;
  for (int iss = 1; iss <= total; iss++) {
    for (;;) {
      double exs = (isex < nex) ? ex[isex].s : 1.e30;
      if (!(exs < fmin(th_s, r_s))) break;
      const Extra &e = ex[isex];
      if (e.radius <= rmaxt && e.radius >= rmint && (e.s - sprev) > eps * e.radius && e.s >= sbeg &&
          e.s <= send) {
        emit_node<COUNT>(P, R, E, base, iray, 3, e.radius, e.theta, e.ir, e.it, e.s);
        sprev = e.s;
      }
      isex++;
    }
    if (th_s < r_s) {
      if (th_rad <= rmaxt && th_rad >= rmint && (th_s - sprev) > eps * th_rad && th_s >= sbeg &&
          th_s <= send) {
        emit_node<COUNT>(P, R, E, base, iray, 2, th_rad, TCf(g, th_it), 0, th_it, th_s);
        sprev = th_s;
      }
      ist++;
      if (ist <= R.ith_amount) th_elem(R, ist, th_s, th_it, th_rad);
      else th_s = 1.e30;
    } else {
      if (r_rad <= rmaxr && r_rad >= rminr && (r_s - sprev) > eps * r_rad && r_s >= sbeg && r_s <= send) {
        emit_node<COUNT>(P, R, E, base, iray, 1, r_rad, 0.0, r_ix, 0, r_s);
        sprev = r_s;
      }
      isr++;
      if (isr <= R.ir_amount) r_elem(R, isr, r_s, r_ix, r_rad);
      else r_s = 1.e30;
    }
  }
  if (COUNT) P.node_cnt[iray] = E.n;
}

}  // namespace

void launch_tan2(const GridDev &g, double *tan2, cudaStream_t st) {
  tan2_kernel<<<(g.nt / 2 + 1 + 127) / 128, 128, 0, st>>>(g, tan2);
}

void launch_roots(const GeomParams &P, cudaStream_t st) {
  const long long n = (long long)(P.ray_hi - P.ray_lo + 1) * (P.g.nt / 2 + P.g.nr + 1);
  if (n > 0) roots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P);
}

// shared memory of one geom_kernel warp: the staged streams of its ray
size_t geom_smem_bytes(const GridDev &g) {
  const size_t nth2 = (size_t)(2 * (g.nt / 2)), nr2 = (size_t)(2 * (g.nr + 1)), mmax = nth2 + nr2 + 34;
  return nth2 * (8 + 8 + 4) + nr2 * 8 + mmax * 4 + 34 * (8 * 4 + 4 * 3) + 64;
}
void launch_geom(const GeomParams &P, bool count, cudaStream_t st) {
  const int nrays = P.ray_hi - P.ray_lo + 1;
  if (nrays <= 0) return;
  // one warp per ray below ~100 rays per SM, one thread per ray above
  const bool warp = nrays < kGeomWarpMax;
  const size_t smem = geom_smem_bytes(P.g);
  static size_t attr = 0;
  if (warp && smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(geom_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(geom_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  const int tb = (nrays + 127) / 128;
  if (count) {
    if (warp) geom_kernel<true, true><<<nrays, 32, smem, st>>>(P);
    else geom_kernel<true, false><<<tb, 128, 0, st>>>(P);
  } else {
    if (warp) geom_kernel<false, true><<<nrays, 32, smem, st>>>(P);
    else geom_kernel<false, false><<<tb, 128, 0, st>>>(P);
    if (P.ntot > 0) node_kernel<<<(unsigned)((P.ntot + 255) / 256), 256, 0, st>>>(P, P.ntot);
  }
}

}  // namespace rl

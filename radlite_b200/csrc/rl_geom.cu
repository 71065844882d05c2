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
  double x0, z0, costh0, sinth0, costh02, sinth02, pitheta0;
  // theta-crossing stream
  int isnr, isdbl, iy_first, iup, branchA, ith_amount;
  // R-crossing stream
  int ir_min, ir_amount;
  double rstar;
};

// both roots of the theta=const cone iy (telescope.F:2960-2990), ordered sar1 <= sar2
__device__ bool th_roots(const Ray &R, int iy, double &sar1, double &sar2) {
  double tanth2 = R.tan2[iy];  // tan(theta_iy)^2, tabulated once per grid by tan2_kernel
  double a = tanth2 * R.costh02 - R.sinth02;
  double b = 2.0 * tanth2 * R.costh0 * R.z0;
  double c = tanth2 * R.z0 * R.z0 - R.x0 * R.x0;
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

// k-th theta crossing along the ray, k = 1..ith_amount (telescope.F:3024-3194 ordering)
__device__ void th_elem(const Ray &R, int k, double &s, int &itheta) {
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
  double a1 = 0.0, a2 = 0.0;
  th_roots(R, iy, a1, a2);
  s = which ? a2 : a1;
  bool mirrored = (pat == 1) ? (R.iup == 1) : (R.iup != 1);
  itheta = mirrored ? (R.g.nt + 1 - iy) : iy;
}
__device__ double th_s_at(const Ray &R, int k) {
  if (k > R.ith_amount) return 1.e30;
  double s;
  int it;
  th_elem(R, k, s, it);
  return s;
}

// k-th R crossing, k = 1..ir_amount (telescope.F:3284-3335)
__device__ void r_elem(const Ray &R, int k, double &s, int &ix, double &r) {
  int j = (k <= R.ir_amount / 2) ? k : (R.ir_amount + 1 - k);
  ix = R.g.nr - (j - 1);
  r = (ix == 0) ? R.rstar : RCf(R.g, ix);
  double b = 2.0 * R.costh0 * R.z0;
  double c = R.x0 * R.x0 + R.z0 * R.z0 - r * r;
  double sdiscr = b * b - 4.0 * 1.0 * c + kTelescEps * b * b;
  sdiscr = sqrt(sdiscr);  // sdiscr<0 cannot occur for ix>=ir_min (reference: stop 13)
  s = (k <= R.ir_amount / 2) ? (-b - sdiscr) / (2.0 * 1.0) : (-b + sdiscr) / (2.0 * 1.0);
}
__device__ double r_s_at(const Ray &R, int k) {
  if (k <= 0) return 0.0;  // r_s(0): read uninitialised by the reference for the outer ring; 0 by convention
  if (k > R.ir_amount) return 1.e30;
  double s, r;
  int ix;
  r_elem(R, k, s, ix, r);
  return s;
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
  double phi = (x0 < 0.0) ? asin(sinphi) : (kPi - asin(sinphi));
  while (phi < 0.0) phi = phi + 2.0 * kPi;
  while (phi >= 2.0 * kPi) phi = phi - 2.0 * kPi;
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
  o.dvmu = 3.335668e-11 * (mu * v1 + sqrt(1.0 - mu * mu) * (v2 * sin(phi) + v3 * cos(phi)));
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

// COUNT: count only.  Otherwise the centre ray (the only one with state that runs along the ray: the
// star is mixed in once) is finished here, every other ray only leaves its LightNode list.
template <bool COUNT>
__device__ void emit_node(const GeomParams &P, const Ray &R, Emit &E, long long base, int iray,
                          int icr, double radius, double theta, int ir, int it, double s,
                          double znew, double bnew) {
  if (COUNT) {
    E.n++;
    return;
  }
  const long long idx = base + E.n;
  if (iray != 0) {
    LightNode ln;
    ln.s = s;
    ln.radius = radius;
    ln.theta = theta;
    ln.pk = (uint32_t)icr | ((uint32_t)ir << 2) | ((uint32_t)it << 17);
    ln.iray = iray;
    static_cast<LightNode *>(P.light)[idx] = ln;
    E.n++;
    return;
  }
  const NodePoint pt = node_point(P, R.x0, R.z0, R.costh0, znew, bnew, icr, radius, theta, ir, it, s);
  const bool star_hit = sqrt(R.x0 * R.x0 + R.z0 * R.z0 * (1.0 - R.costh02)) <= P.rstar;  // tr_b (telescope.F:3268)
  P.nodes.rec[idx] = node_record(P, iray, pt, icr, ir, s, E.n > 0, E.s_prev, E.ir_old, E.icr_old, E.dvmu_prev,
                                 E.lw_prev, E.star_done, star_hit);
  E.dvmu_prev = pt.dvmu;
  E.lw_prev = pt.lw;
  E.ir_old = ir;
  E.icr_old = icr;
  E.s_prev = s;
  E.n++;
}

// one thread per node: the record of node i from its LightNode and the one before it on the same ray
// (whose point values are simply evaluated again: cheaper than a second pass over the records)
__global__ void __launch_bounds__(256) node_kernel(GeomParams P, long long ntot) {
  // what the next node of the ray needs from this one travels through shared memory (the threads of a
  // block hold consecutive nodes); only thread 0 evaluates its predecessor a second time
  __shared__ double s_dvmu[256], s_lw[256], s_s[256];
  __shared__ int s_ir[256], s_icr[256];
  const int t = threadIdx.x;
  const long long i = (long long)blockIdx.x * blockDim.x + t;
  // (the centre ray was finished by geom_kernel: its slots hold no LightNode)
  const bool active = i < ntot && i >= P.node_off[1];
  const LightNode *light = static_cast<const LightNode *>(P.light);
  LightNode me;
  me.s = me.radius = me.theta = 0.0;
  me.pk = 0;
  me.iray = 1;
  if (active) me = light[i];
  const int iray = me.iray;
  const double x0 = P.x0[iray], z0 = P.z0[iray];
  const double costh0 = cos(P.theta0), sinth0 = sin(P.theta0);
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
  int star_done = 1;  // (the circular camera mixes the star into the centre ray only)
  const bool star_hit = sqrt(x0 * x0 + z0 * z0 * (1.0 - costh0 * costh0)) <= P.rstar;  // tr_b (telescope.F:3268)
  P.nodes.rec[i] = node_record(P, iray, pt, icr, ir, me.s, have_prev, s_prev, ir_old, icr_old, dvmu_prev, lw_prev,
                               star_done, star_hit);
}

// tan(theta_iy)^2 of the stored hemisphere's cones (telescope.F:2960-2962 evaluates it per ray and cone)
__global__ void tan2_kernel(GridDev g, double *tan2) {
  const int iy = blockIdx.x * blockDim.x + threadIdx.x;
  if (iy < 1 || iy > g.nt / 2) return;
  const double t = tan(TCf(g, iy));
  tan2[iy] = t * t;
}

template <bool COUNT>
__global__ void __launch_bounds__(128) geom_kernel(GeomParams P) {
  const int iray = blockIdx.x * blockDim.x + threadIdx.x;
  if (iray >= P.nray) return;
  const GridDev &g = P.g;
  const int nr = g.nr, nt = g.nt;
  if (iray < P.ray_lo || iray > P.ray_hi ||
      (P.rect && iray > 0 && !(P.rb[iray] < P.bskip))) {  // another rank's ring block / a pixel off the model
    if (COUNT) P.node_cnt[iray] = 0;
    return;
  }
  Ray R;
  R.g = g;
  R.tan2 = P.tan2;
  R.x0 = P.x0[iray];
  R.z0 = P.z0[iray];
  R.rstar = P.rstar;
  const double theta0 = P.theta0;
  R.pitheta0 = 0.5 * kPi - theta0;
  R.costh0 = cos(theta0);
  R.sinth0 = sin(theta0);
  R.costh02 = R.costh0 * R.costh0;
  R.sinth02 = R.sinth0 * R.sinth0;
  const double x0 = R.x0, z0 = R.z0;
  const double eps = kTelescEps, epsplus = 1.e1 * kTelescEps;

  // ---- theta crossings: which cones are hit (telescope.F:2958-3014) ----
  {
    int cnt = 0, first = 0, dbl = 0;
    const int iyeq = nt / 2;
    for (int iy = 1; iy <= iyeq; iy++) {
      double a1, a2;
      if (th_roots(R, iy, a1, a2)) {
        if (cnt == 0) first = iy;
        cnt++;
        double pitheta = 0.5 * kPi - TCf(g, iy);
        if (fabs(pitheta) > fabs(R.pitheta0)) dbl++;
      }
    }
    // hit cones form a contiguous tail iy_first..nth (disc>0 <=> tan^2 theta above a threshold)
    if (cnt > 0 && first + cnt - 1 != iyeq) atomicCAS(P.status, 0, 9001);
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
    int ix;
    for (ix = 1; ix <= nr; ix++)
      if (RCf(g, ix) > bimpact) break;
    if (ix > nr) {
      atomicCAS(P.status, 0, 13);
      if (COUNT) P.node_cnt[iray] = 0;
      return;
    }
    R.ir_min = ix;
    if (bimpact <= P.rstar) R.ir_min = 0;
    R.ir_amount = 2 * (nr + 1 - R.ir_min);
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
  const double znew = z0 * R.sinth02;
  const double bnew = sqrt(x0 * x0 + z0 * z0 * R.sinth02);
  const long long base = COUNT ? 0 : P.node_off[iray];
  // rays finished by node_kernel: the bracketing searches of a node (theta cell of an R crossing, radial
  // cell of a theta crossing) are left to it as well -- they are the expensive serial part of the merge
  const bool lazy = COUNT || iray != 0;
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
  double th_s = 1.e30, r_s = 1.e30, r_rad = 0.0;
  int th_it = 0, r_ix = 0;
  if (ist <= R.ith_amount) th_elem(R, ist, th_s, th_it);
  if (isr <= R.ir_amount) r_elem(R, isr, r_s, r_ix, r_rad);
  const int total = R.ir_amount + R.ith_amount;
  for (int iss = 1; iss <= total; iss++) {
    for (;;) {
      double exs = (isex < nex) ? ex[isex].s : 1.e30;
      if (!(exs < fmin(th_s, r_s))) break;
      const Extra &e = ex[isex];
      if (e.radius <= rmaxt && e.radius >= rmint && (e.s - sprev) > eps * e.radius && e.s >= sbeg &&
          e.s <= send) {
        emit_node<COUNT>(P, R, E, base, iray, 3, e.radius, e.theta, e.ir, e.it, e.s, znew, bnew);
        sprev = e.s;
      }
      isex++;
    }
    if (th_s < r_s) {
      double th_rad = radius_of_s(R, th_s);
      if (th_rad <= rmaxt && th_rad >= rmint && (th_s - sprev) > eps * th_rad && th_s >= sbeg &&
          th_s <= send) {
        int th_ir = 0;
        if (!lazy) th_ir = ir_of_radius(g, th_rad);
        emit_node<COUNT>(P, R, E, base, iray, 2, th_rad, TCf(g, th_it), th_ir, th_it, th_s, znew, bnew);
        sprev = th_s;
      }
      ist++;
      if (ist <= R.ith_amount) th_elem(R, ist, th_s, th_it);
      else th_s = 1.e30;
    } else {
      if (r_rad <= rmaxr && r_rad >= rminr && (r_s - sprev) > eps * r_rad && r_s >= sbeg && r_s <= send) {
        double th = 0.0;
        int it = 0;
        if (!lazy) {
          th = theta_of_s(R, r_s);
          it = hunt_bisect(g.tc + 2, nt, th);
        }
        emit_node<COUNT>(P, R, E, base, iray, 1, r_rad, th, r_ix, it, r_s, znew, bnew);
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

void launch_geom(const GeomParams &P, bool count, cudaStream_t st) {
  const int threads = 128;
  const int blocks = (P.nray + threads - 1) / threads;
  if (count) {
    geom_kernel<true><<<blocks, threads, 0, st>>>(P);
  } else {
    geom_kernel<false><<<blocks, threads, 0, st>>>(P);
    if (P.ntot > 0) node_kernel<<<(unsigned)((P.ntot + 255) / 256), 256, 0, st>>>(P, P.ntot);
  }
}

}  // namespace rl

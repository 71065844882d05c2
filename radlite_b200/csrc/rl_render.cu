// rl_render.cu -- per-line preparation, ray integration and flux reduction kernels.
//
//   prep_cells_kernel   line.F:3502-3608 (dust source term at line centre), setup.F:937 (bplanck),
//                       line.F:4069-4088 (N_up, N_down per cell)
//   span_kernel         telescope.F:4265-4270 (minvel/maxvel of a ray) + the NONREDUNDANT channel
//                       selection of telescope.F:544-612 -> work items
//   integrate_kernel    telescope.F:3889-4312 (charintline), line.F:4636-4848
//                       (clever_integrate_element_linedust), line.F:4515-4624
//                       (integrate_element_linedust), line.F:2280 (voigt_profile),
//                       transfer.F:1498 (qdr_src_2).  tile_kernel: one block = one ray x a tile of
//                       (line, channel) items, node data staged in shared memory; center_kernel:
//                       the centre ray (star mixing, char_tau_center).
//   fill_kernel         telescope.F:582-612 continuum copy for the skipped channels
//   ringsum/flux        telescope.F:1388-1433 (calc_freq_flux_observer), fixed summation order
#include "rl_types.h"

#include <math.h>

namespace rl {

namespace {

__device__ __forceinline__ double4 ldg4(const double4 *p) {
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
  const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

struct Node {
  double ds, dvmu, lw, q, wr, wt;
  int4 cells;
  uint32_t flags;
};

__device__ __forceinline__ Node load_node(const NodeRec *__restrict__ rec, long long i) {
  const double2 *p = reinterpret_cast<const double2 *>(rec + i);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  const int4 d = __ldg(reinterpret_cast<const int4 *>(p + 3));
  Node n;
  n.ds = a.x;
  n.dvmu = a.y;
  n.lw = b.x;
  n.q = b.y;
  n.wr = c.x;
  n.wt = c.y;
  n.cells = d;
  n.flags = (uint32_t)d.x >> kCellFlagShift;
  n.cells.x = d.x & kCellMask;
  return n;
}

// interpolate the per-line cell record {src_dust, alp_dust, N_up, N_down} at a node
// (line.F:4054-4197; three cases by crossing type)
__device__ __forceinline__ double4 gather_line(const double4 *__restrict__ cellL, int4 c, double dr,
                                               double dt, int icr) {
  double4 a = ldg4(cellL + c.x), o;
  if (icr == 1) {
    double4 b = ldg4(cellL + c.y);
    o.x = (1.0 - dt) * a.x + dt * b.x;
    o.y = (1.0 - dt) * a.y + dt * b.y;
    o.z = (1.0 - dt) * a.z + dt * b.z;
    o.w = (1.0 - dt) * a.w + dt * b.w;
  } else if (icr == 2) {
    double4 b = ldg4(cellL + c.z);
    o.x = (1.0 - dr) * a.x + dr * b.x;
    o.y = (1.0 - dr) * a.y + dr * b.y;
    o.z = (1.0 - dr) * a.z + dr * b.z;
    o.w = (1.0 - dr) * a.w + dr * b.w;
  } else {
    double4 b = ldg4(cellL + c.y), cc = ldg4(cellL + c.z), d = ldg4(cellL + c.w);
    o.x = (1.0 - dr) * ((1.0 - dt) * a.x + dt * b.x) + dr * ((1.0 - dt) * cc.x + dt * d.x);
    o.y = (1.0 - dr) * ((1.0 - dt) * a.y + dt * b.y) + dr * ((1.0 - dt) * cc.y + dt * d.y);
    o.z = (1.0 - dr) * ((1.0 - dt) * a.z + dt * b.z) + dr * ((1.0 - dt) * cc.z + dt * d.z);
    o.w = (1.0 - dr) * ((1.0 - dt) * a.w + dt * b.w) + dr * ((1.0 - dt) * cc.w + dt * d.w);
  }
  return o;
}

// ------------------------------------------------------------------------------------------
// per-line preparation
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double bplanck_dev(double temp, double nu) {
  if (temp == 0.0) return 0.0;
  return 1.47455e-47 * nu * nu * nu / (exp(4.7989e-11 * nu / temp) - 1.0) + 1.e-290;
}

__global__ void __launch_bounds__(256) prep_cells_kernel(PrepParams P) {
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  if (cell >= P.ncell) return;
  double src = 0.0, alp = 0.0;
  if (P.use_dust) {
    const int inu = P.inudust[l];
    if (!(inu == 0 || inu == P.ncf)) {
      const double wgt = P.wgt[l], freq = P.freq[l];
      for (int is = 0; is < P.nspec; is++)
        for (int iz = 0; iz < P.nsize[is]; iz++) {
          const double *ka = P.kabs + ((size_t)is * P.maxsize + iz) * P.ncf - 1;
          double kappawgt = wgt * ka[inu + 1] + (1.0 - wgt) * ka[inu];
          double rho = P.drho[cell * P.nspec + is];
          double temp = P.dtemp[(cell * P.nspec + is) * P.maxsize + iz];
          src = src + rho * kappawgt * bplanck_dev(temp, freq);
          alp = alp + rho * kappawgt;
        }
      if (P.scat) {
        const double *sc = P.scat + cell * P.ncf - 1;
        src = src + wgt * sc[inu + 1] + (1.0 - wgt) * sc[inu];
      }
      for (int is = 0; is < P.nspec; is++)
        for (int iz = 0; iz < P.nsize[is]; iz++) {
          const double *ks = P.kscat + ((size_t)is * P.maxsize + iz) * P.ncf - 1;
          double kappawgt = wgt * ks[inu + 1] + (1.0 - wgt) * ks[inu];
          alp = alp + P.drho[cell * P.nspec + is] * kappawgt;
        }
    }
  } else {
    src = P.ld_src[(size_t)l * P.ncell + cell];
    alp = P.ld_alp[(size_t)l * P.ncell + cell];
  }
  const double ab = P.abund[cell], rh = P.rho[cell];
  const double *pp = P.popul + cell * P.nlevels;
  double4 o;
  o.x = src;
  o.y = alp;
  o.z = pp[P.lev_up[l] - 1] * ab * rh * P.molpg;
  o.w = pp[P.lev_down[l] - 1] * ab * rh * P.molpg;
  P.cellL[(size_t)l * P.ncell + cell] = o;
}

// ------------------------------------------------------------------------------------------
// velocity span of a ray for one line -> which channels the reference integrates
// one warp per (line, ray)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) span_kernel(RenderParams P) {
  const int lane = threadIdx.x & 31;
  const long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long ntask = (long long)P.nl * P.nray;
  if (task >= ntask) return;
  const int ray = (int)(task / P.nl), l = (int)(task % P.nl);  // ray-major tasks
  if (ray == 0 || !P.nonredundant) {
    if (lane == 0) {
      P.rng[task] = make_int4(1, P.nfr - 1, -1, ray == 0 ? 2 : 1);
      // the centre ray is traced by center_kernel (it also produces char_tau_center)
      P.nitems[task] = ray == 0 ? 0 : P.nfr;
    }
    return;
  }
  const long long n0 = P.node_off[ray], n1 = P.node_off[ray + 1];
  const double4 *cellL = P.cellL + (size_t)l * P.ncell;
  double vmin = 2.0, vmax = -2.0;
  // telescope.F:4265-4270: start node of every segment, i.e. all nodes but the last
  for (long long i = n0 + lane; i < n1 - 1; i += 32) {
    const Node nd = load_node(P.nodes.rec, i);
    const double4 v = gather_line(cellL, nd.cells, nd.wr, nd.wt, nd.flags & kFlagIcrMask);
    if (v.z + v.w > P.levthres) {
      vmin = fmin(vmin, nd.dvmu);
      vmax = fmax(vmax, nd.dvmu);
    }
  }
  for (int o = 16; o; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  // REAL*4 minvel/maxvel (common_telescope.h:17), initial values 1 and -1 (telescope.F:388-391)
  const float minvel = (vmin < 1.0) ? (float)vmin : 1.0f;
  const float maxvel = (vmax > -1.0) ? (float)vmax : -1.0f;
  const double hi_lim = (double)maxvel + 2.f * P.aksmax_c;
  const double lo_lim = (double)minvel - 2.f * P.aksmax_c;
  const double *velo = P.velo + (size_t)l * P.nfr;
  // channel 0 (reference inu=1) is always integrated; 1..nfr-1 only inside [lo_lim, hi_lim]
  int lo = P.nfr, hi = -1, c0 = P.nfr;
  for (int c = 1 + lane; c < P.nfr; c += 32) {
    const double v = velo[c];
    const bool in = (v <= hi_lim) && (v >= lo_lim);
    if (in) {
      lo = min(lo, c);
      hi = max(hi, c);
    } else {
      c0 = min(c0, c);
    }
  }
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    c0 = min(c0, __shfl_xor_sync(0xffffffffu, c0, o));
  }
  if (lane == 0) {
    const double v0 = velo[0];
    const bool ch0_out = (v0 > hi_lim) || (v0 < lo_lim);  // telescope.F:550-551
    int n = 1 + ((hi >= lo) ? (hi - lo + 1) : 0);
    int c0i = -1;
    if (!ch0_out && c0 < P.nfr) {  // continuum not known after channel 0: the first skipped channel is integrated
      c0i = c0;
      n++;
    }
    if (hi < lo) { lo = 1; hi = 0; }
    P.rng[task] = make_int4(lo, hi, c0i, ch0_out ? 0 : 1);
    P.nitems[task] = (unsigned)n;
  }
}

// number of channels the reference integrates for a task and the j-th of them
__device__ __forceinline__ int task_nchan(const RenderParams &P, const int4 rg) {
  if (rg.w == 2 || !P.nonredundant) return P.nfr;
  return 1 + ((rg.y >= rg.x) ? (rg.y - rg.x + 1) : 0) + (rg.z >= 0 ? 1 : 0);
}
__device__ __forceinline__ int task_chan(const RenderParams &P, const int4 rg, int j, bool &masked) {
  if (rg.w == 2) { masked = false; return j; }  // centre ray: the reference never sets its mask
  if (!P.nonredundant) { masked = (j == 0); return j; }  // telescope.F:548 only
  const int nin = (rg.y >= rg.x) ? (rg.y - rg.x + 1) : 0;
  if (j == 0) { masked = true; return 0; }
  if (j <= nin) { masked = true; return rg.x + j - 1; }
  masked = false;
  return rg.z;
}

// ------------------------------------------------------------------------------------------
// the formal solution, reference-ordered scalar version: used for the centre ray (which also
// yields char_tau_center), for sub-gridded segments and for the rare zero-continuum fallback
// ------------------------------------------------------------------------------------------
struct Carry {
  double srcl0, alpl0;
  int init;
};

// transfer.F:1498-1571
__device__ __forceinline__ double qdr_src_2(double inten, double js1, double alp1, double js2,
                                            double alp2, double ds) {
  double a, b, xp, src1, src2, q;
  const double dtau1 = 0.5 * (alp1 + alp2) * ds;
  const double theomax = 0.5 * (js1 + js2) * ds;
  if (dtau1 > 1.e-6) {
    xp = exp(-dtau1);
    const double e0 = 1.0 - xp;
    const double e1 = dtau1 - e0;
    b = e1 / dtau1;
    a = e0 - b;
  } else {
    a = 0.5 * dtau1;
    b = 0.5 * dtau1;
    xp = 1.0 - dtau1;
  }
  if (alp1 > 0.0) src1 = js1 / alp1;
  else if (alp2 > 0.0) src1 = js2 / alp2;
  else src1 = 0.0;
  if (alp2 > 0.0) src2 = js2 / alp2;
  else if (alp1 > 0.0) src2 = js1 / alp1;
  else src2 = 0.0;
  if (dtau1 > (double)1e-9f) q = a * src1 + b * src2;
  else q = theomax;
  q = fmin(q, theomax);
  return inten * xp + q;
}

// line.F:4515-4624 (+ voigt_profile line.F:2280-2314)
__device__ __noinline__ void integrate_element(const LineDev &L, double dnu_ch, double &inten,
                                               double ds, double srcd0, double srcd1, double alpd0,
                                               double alpd1, double lw0, double lw1, double dvmu0,
                                               double dvmu1, double nup0, double nup1,
                                               double ndown0, double ndown1, Carry &k, double &tau) {
  const double lwav = 0.5 * (lw0 + lw1);
  const double aa = 3.33567e-6 * L.nu0 * lwav;
  const double norm = 0.56419583546 / aa;
  if (k.init) {
    const double dnu0 = dnu_ch - L.nu0 * dvmu0;
    const double u0 = dnu0 / aa;
    const double phiprof0 = norm * exp(-(u0 * u0));
    k.srcl0 = 5.27296241956e-28 * L.nu0 * nup0 * L.aud * phiprof0;
    k.alpl0 = 5.27296241956e-28 * L.nu0 * phiprof0 * (ndown0 * L.bdu - nup0 * L.bud);
  }
  const double dnu1 = dnu_ch - L.nu0 * dvmu1;
  const double u1 = dnu1 / aa;
  const double phiprof1 = norm * exp(-(u1 * u1));
  const double srcl1 = 5.27296241956e-28 * L.nu0 * nup1 * L.aud * phiprof1;
  const double alpl1 = 5.27296241956e-28 * L.nu0 * phiprof1 * (ndown1 * L.bdu - nup1 * L.bud);
  const double src0 = srcd0 + k.srcl0, src1 = srcd1 + srcl1;
  const double alp0 = alpd0 + k.alpl0, alp1 = alpd1 + alpl1;
  inten = qdr_src_2(inten, src0, alp0, src1, alp1, ds);
  tau = tau + 0.5 * (alp0 + alp1) * ds;
  k.srcl0 = srcl1;
  k.alpl0 = alpl1;
  k.init = 0;
}

// the sub-gridded segment of line.F:4745-4833: returns the number of element integrations
__device__ __noinline__ int subgrid_segment(const LineDev &L, double dnu_ch, double &inten, double ds,
                                            double sleft, double sright, const double4 &v0,
                                            const double4 &v1, double dvmu0, double dvmu1, double lw,
                                            Carry &k, double &tau) {
  const double lg_ds = (sright - sleft) / (kLgNrMax - 1.0);
  double sp = 0.0, nup_p = v0.z, ndn_p = v0.w, dv_p = dvmu0, sd_p = v0.x, ad_p = v0.y;
  int n = 0;
  for (int j = 1; j <= kLgNrMax + 1; j++) {
    double s, nup_c, ndn_c, dv_c, sd_c, ad_c;
    if (j <= kLgNrMax) {
      s = sleft + (j - 1) * lg_ds;
      if (!(s > 0.0 && s < ds)) continue;
      const double eps = s / ds, epsp = 1.0 - eps;
      nup_c = epsp * v0.z + eps * v1.z;
      ndn_c = epsp * v0.w + eps * v1.w;
      dv_c = epsp * dvmu0 + eps * dvmu1;
      sd_c = epsp * v0.x + eps * v1.x;
      ad_c = epsp * v0.y + eps * v1.y;
    } else {
      s = ds; nup_c = v1.z; ndn_c = v1.w; dv_c = dvmu1; sd_c = v1.x; ad_c = v1.y;
    }
    integrate_element(L, dnu_ch, inten, s - sp, sd_p, sd_c, ad_p, ad_c, lw, lw, dv_p, dv_c, nup_p,
                      nup_c, ndn_p, ndn_c, k, tau);
    n++;
    sp = s; nup_p = nup_c; ndn_p = ndn_c; dv_p = dv_c; sd_p = sd_c; ad_p = ad_c;
  }
  return n;
}

// telescope.F:3889-4312 for ray `ray`, channel `ch` (0-based) of line slot `l`, reference order
__device__ __noinline__ double integrate_ray_channel(const RenderParams &P, int l, int ray, int ch,
                                                     double &tau, unsigned &nelem, int &maser) {
  const LineDev L = P.lines[l];
  const double4 *__restrict__ cellL = P.cellL + (size_t)l * P.ncell;
  const long long n0 = P.node_off[ray], n1 = P.node_off[ray + 1];
  const double dnu_ch = P.line_dnu[(size_t)l * P.nfr + ch];
  const double velo_ch = dnu_ch / L.nu0;
  double inten = (P.out_itype == 3) ? P.isrf_line[(size_t)l * P.nfr + ch] : L.i_outer;
  tau = 0.0;
  nelem = 0;
  if (n1 <= n0) return inten;
  Carry k;
  k.init = 1;
  k.srcl0 = k.alpl0 = 0.0;
  Node nd = load_node(P.nodes.rec, n0);
  double4 v0 = gather_line(cellL, nd.cells, nd.wr, nd.wt, nd.flags & kFlagIcrMask);
  double dvmu0 = nd.dvmu, lw0 = nd.lw;
  for (long long i = n0 + 1; i < n1; i++) {
    nd = load_node(P.nodes.rec, i);
    const uint32_t fl = nd.flags;
    const double ds = nd.ds, dvmu1 = nd.dvmu, lw1 = nd.lw;
    const double4 v1 = gather_line(cellL, nd.cells, nd.wr, nd.wt, fl & kFlagIcrMask);
    if (fl & (kFlagInit | kFlagStar | kFlagZero)) {
      if (fl & kFlagZero) inten = 0.0;
      if (fl & kFlagStar)
        inten = (1.0 - P.starfract) * inten + P.starfract * P.star_line[(size_t)l * P.nfr + ch];
      k.init = 1;
    }
    bool done = false;
    if (P.subgrid && 2.0 * 3.0 * nd.q > 1.0) {  // line.F:4715
      const double s_c = ds * (velo_ch - dvmu0) / (dvmu1 - dvmu0);
      const double dls = ds / nd.q;
      const double sright = s_c + 3.0 * dls, sleft = s_c - 3.0 * dls;
      if (sright > 0.0 && sleft < ds) {
        nelem += subgrid_segment(L, dnu_ch, inten, ds, sleft, sright, v0, v1, dvmu0, dvmu1,
                                 0.5 * (lw0 + lw1), k, tau);
        done = true;
      }
    }
    if (!done) {
      integrate_element(L, dnu_ch, inten, ds, v0.x, v1.x, v0.y, v1.y, lw0, lw1, dvmu0, dvmu1, v0.z,
                        v1.z, v0.w, v1.w, k, tau);
      nelem++;
    }
    if (k.alpl0 * ds < (double)(-0.01f)) maser = 1;  // telescope.F:4295
    v0 = v1;
    dvmu0 = dvmu1;
    lw0 = lw1;
  }
  return inten;
}

__device__ __forceinline__ long long img_row(const RenderParams &P, int ray) {
  return ray == 0 ? 0 : (long long)P.nphi + (ray - 1);
}

// reciprocal to ~1 ulp: hardware seed (2^-23) + two Newton steps.  Not correctly rounded; used
// where the reference divides (source function j/alpha, e1/dtau) -- differences are O(1e-16).
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// ------------------------------------------------------------------------------------------
// tile_kernel: the formal solution for one ray x a tile of (line, channel) items.
//
// One thread block owns one camera ray and kTileThreads consecutive entries of the ray's item list
// (all channels the reference integrates, lines in index order: "ray x line tile").  The ray's nodes
// are processed in chunks: first the block cooperatively stages, for every (node, line of the tile),
// the interpolated cell values and all per-segment constants of the line profile in shared memory
// (one thread per pair, so this work is done once per ray, line and node instead of once per
// channel); then every thread walks the chunk for its own channel, reading the staged values
// (shared-memory broadcast within a line) and advancing I <- I e^{-dtau} + Q.  All threads of a
// block walk the same ray, so there is no trip-count divergence; the only divergent paths are the
// rare ones of the reference itself (velocity sub-gridding, re-initialisation in the inner hole).
// ------------------------------------------------------------------------------------------
struct __align__(16) StagedLine {  // per (node, line): 64 bytes
  double srcd, alpd;   // dust source / opacity at the node (line.F:4058-4063)
  double cN, kk;       // c_src N_up ; c_alp (N_down B_du - N_up B_ud)
  double inv_aa, nudv; // 1/(k_aa * mean width of the segment ending here) ; nu0 * Omega.v/c
  double A1, K1;       // cN, kk times the profile norm 0.5641.../aa of the segment ending here
};
struct __align__(16) StagedNode {  // per node: 48 bytes
  double ds, dvmu, q, lwav;
  uint32_t flags, pad0, pad1, pad2;
};

// exp(x) for x <= 0: x = (64 k + j) ln2/64 + r, |r| <= ln2/128; exp(r) by a degree-5 polynomial
// (truncation 3.5e-17), 2^(j/64) from a 64-entry shared-memory table, 2^k through the exponent
// field.  Branch free; results below exp(-700) are flushed to 0.  ~1 ulp, 10 FP64 instructions.
__device__ __forceinline__ double exp_neg_tab(double x, const double *__restrict__ tab) {
  const double t = fma(x, 92.332482616893656877, 6755399441055744.0);  // 64/ln2, 1.5*2^52
  const int n = __double2loint(t);
  const double fn = t - 6755399441055744.0;
  double r = fma(fn, -1.08304246932675596327e-02, x);   // ln2_hi/64 (ln2_hi has 32 trailing zero bits)
  r = fma(fn, -2.98158582698529328128e-12, r);           // ln2_lo/64
  double p = fma(r, 8.33333333333333333333e-03, 4.16666666666666666667e-02);
  p = fma(p, r, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double tj = tab[n & 63];
  const int hi = __double2hiint(tj) + ((n >> 6) << 20);
  const double sc = __hiloint2double(hi, __double2loint(tj));
  const bool tiny = (unsigned)__double2hiint(x) > 0xC085E000u;  // x < -700
  return tiny ? 0.0 : p * sc;
}

// sub-gridded segment in the staged (cN, kk) form (line.F:4745-4833); reference-ordered arithmetic
__device__ __noinline__ int subgrid_tile(const LineDev &L, double dnu_ch, double &inten, double ds,
                                         double sleft, double sright, double sd0, double ad0,
                                         double cN0, double kk0, double dv0, double sd1, double ad1,
                                         double cN1, double kk1, double dv1, double lw, double &srcl0,
                                         double &alpl0, int init) {
  const double lg_ds = (sright - sleft) / (kLgNrMax - 1.0);
  const double aa = 3.33567e-6 * L.nu0 * (0.5 * (lw + lw));
  const double norm = 0.56419583546 / aa;
  double sp = 0.0, cN_p = cN0, kk_p = kk0, dv_p = dv0, sd_p = sd0, ad_p = ad0;
  int n = 0;
  for (int j = 1; j <= kLgNrMax + 1; j++) {
    double s, cN_c, kk_c, dv_c, sd_c, ad_c;
    if (j <= kLgNrMax) {
      s = sleft + (j - 1) * lg_ds;
      if (!(s > 0.0 && s < ds)) continue;
      const double eps = s / ds, epsp = 1.0 - eps;
      cN_c = epsp * cN0 + eps * cN1;
      kk_c = epsp * kk0 + eps * kk1;
      dv_c = epsp * dv0 + eps * dv1;
      sd_c = epsp * sd0 + eps * sd1;
      ad_c = epsp * ad0 + eps * ad1;
    } else {
      s = ds; cN_c = cN1; kk_c = kk1; dv_c = dv1; sd_c = sd1; ad_c = ad1;
    }
    if (init) {
      const double u0 = (dnu_ch - L.nu0 * dv_p) / aa;
      const double phi0 = norm * exp(-(u0 * u0));
      srcl0 = cN_p * phi0;
      alpl0 = kk_p * phi0;
      init = 0;
    }
    const double u1 = (dnu_ch - L.nu0 * dv_c) / aa;
    const double phi1 = norm * exp(-(u1 * u1));
    const double srcl1 = cN_c * phi1, alpl1 = kk_c * phi1;
    inten = qdr_src_2(inten, sd_p + srcl0, ad_p + alpl0, sd_c + srcl1, ad_c + alpl1, s - sp);
    srcl0 = srcl1;
    alpl0 = alpl1;
    n++;
    sp = s; cN_p = cN_c; kk_p = kk_c; dv_p = dv_c; sd_p = sd_c; ad_p = ad_c;
  }
  return n;
}

__global__ void __launch_bounds__(kTileThreads, 4) tile_kernel(RenderParams P) {
  extern __shared__ double4 smem_raw[];
  __shared__ double s_exptab[64];
  __shared__ int s_ray, s_l0, s_l1;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 64) s_exptab[tid] = exp2((double)tid * (1.0 / 64.0));
  if (tid == 0) {
    // which ray does this block belong to: largest r with cta_off[r] <= blockIdx.x
    int lo = 0, hi = P.nray;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (P.cta_off[mid] <= blockIdx.x) lo = mid;
      else hi = mid;
    }
    s_ray = lo;
    const unsigned *off = P.item_off + (size_t)lo * P.nl;
    const unsigned base = off[0], M = off[P.nl] - base;
    const unsigned g0 = (blockIdx.x - P.cta_off[lo]) * kTileThreads;
    const unsigned g1 = min(g0 + kTileThreads, M) - 1;
    // lines of the first and last item: largest l with off[l]-base <= g
    int a = 0, b = P.nl;
    while (b - a > 1) {
      const int mid = (a + b) >> 1;
      if (off[mid] - base <= g0) a = mid;
      else b = mid;
    }
    s_l0 = a;
    b = P.nl;
    while (b - a > 1) {
      const int mid = (a + b) >> 1;
      if (off[mid] - base <= g1) a = mid;
      else b = mid;
    }
    s_l1 = a;
  }
  __syncthreads();
  const int ray = s_ray, l0 = s_l0, nlc = s_l1 - s_l0 + 1;
  const unsigned *off = P.item_off + (size_t)ray * P.nl;
  const unsigned base = off[0], M = off[P.nl] - base;
  const unsigned my = (blockIdx.x - P.cta_off[ray]) * kTileThreads + tid;
  const bool active = my < M;
  // my line and channel
  int l = l0;
  if (active) {
    int a = l0, b = s_l1 + 1;
    while (b - a > 1) {
      const int mid = (a + b) >> 1;
      if (off[mid] - base <= my) a = mid;
      else b = mid;
    }
    l = a;
  }
  const int ml = l - l0;
  const long long task = (long long)ray * P.nl + l;
  const int4 rg = P.rng[task];
  bool masked = false;
  const int ch = active ? task_chan(P, rg, (int)(my - (off[l] - base)), masked) : 0;
  const LineDev L = P.lines[l];
  const double dnu = P.line_dnu[(size_t)l * P.nfr + ch];
  double inten = (P.out_itype == 3) ? P.isrf_line[(size_t)l * P.nfr + ch] : L.i_outer;
  const long long n0 = P.node_off[ray];
  const int N = (int)(P.node_off[ray + 1] - n0);
  // shared-memory carve-up: StagedNode[nch+1] then StagedLine[(nch+1)*nlc]
  int nch = (P.smem_budget - (int)sizeof(StagedNode)) / (int)(nlc * sizeof(StagedLine) + sizeof(StagedNode)) - 1;
  nch = max(1, min(kTileChunk, nch));
  StagedNode *sn = reinterpret_cast<StagedNode *>(smem_raw);
  StagedLine *sl = reinterpret_cast<StagedLine *>(sn + (nch + 1));
  double src0 = 0.0, alp0 = 0.0, r0 = 0.0;
  int init = 1, maser = 0;
  unsigned nelem = 0;
  for (int c0 = 1; c0 < N; c0 += nch) {
    const int cnt = min(nch, N - c0);
    // ---- stage nodes c0-1 .. c0+cnt-1 (slot 0 = the previous node) ----
    for (int p = tid; p < (cnt + 1) * nlc; p += kTileThreads) {
      const int slot = p / nlc, m = p - slot * nlc;
      const int node = c0 - 1 + slot;
      const Node nd = load_node(P.nodes.rec, n0 + node);
      const LineDev Lm = P.lines[l0 + m];
      const double4 v = gather_line(P.cellL + (size_t)(l0 + m) * P.ncell, nd.cells, nd.wr, nd.wt,
                                    nd.flags & kFlagIcrMask);
      const double lw_prev = (node > 0) ? __ldg(&P.nodes.rec[n0 + node - 1].lw) : nd.lw;
      const double lwav = 0.5 * (lw_prev + nd.lw);
      StagedLine s;
      s.srcd = v.x;
      s.alpd = v.y;
      s.cN = Lm.c_src * v.z;
      s.kk = Lm.c_alp * (v.w * Lm.bdu - v.z * Lm.bud);
      s.inv_aa = 1.0 / (Lm.k_aa * lwav);
      s.nudv = Lm.nu0 * nd.dvmu;
      const double norm = 0.56419583546 * s.inv_aa;
      s.A1 = s.cN * norm;
      s.K1 = s.kk * norm;
      sl[slot * nlc + m] = s;
      if (m == 0) {
        StagedNode t;
        t.ds = nd.ds;
        t.dvmu = nd.dvmu;
        t.q = nd.q;
        t.lwav = lwav;
        t.flags = nd.flags;
        t.pad0 = t.pad1 = t.pad2 = 0;
        sn[slot] = t;
      }
    }
    __syncthreads();
    if (active) {
      for (int slot = 1; slot <= cnt; slot++) {
        const StagedNode ns = sn[slot];
        const StagedLine s1 = sl[slot * nlc + ml];
        const double ds = ns.ds, hds = 0.5 * ns.ds;
        if (ns.flags & (kFlagInit | kFlagStar | kFlagZero)) {  // inner hole / inner boundary (rare)
          if (ns.flags & kFlagZero) inten = 0.0;
          if (ns.flags & kFlagStar)
            inten = (1.0 - P.starfract) * inten + P.starfract * P.star_line[(size_t)l * P.nfr + ch];
          init = 1;
        }
        if (P.subgrid && (2.0 * 3.0 * ns.q > 1.0)) {  // line.F:4715 (rare)
          const double dvmu0 = sn[slot - 1].dvmu;
          const double s_c = ds * (dnu * L.inv_nu0 - dvmu0) / (ns.dvmu - dvmu0);
          const double dls3 = 3.0 * (ds / ns.q);
          const double sright = s_c + dls3, sleft = s_c - dls3;
          if (sright > 0.0 && sleft < ds) {
            const StagedLine s0 = sl[(slot - 1) * nlc + ml];
            double srcl0 = src0 - s0.srcd, alpl0 = alp0 - s0.alpd;
            nelem += subgrid_tile(L, dnu, inten, ds, sleft, sright, s0.srcd, s0.alpd, s0.cN, s0.kk, dvmu0,
                                  s1.srcd, s1.alpd, s1.cN, s1.kk, ns.dvmu, ns.lwav, srcl0, alpl0, init);
            src0 = s1.srcd + srcl0;
            alp0 = s1.alpd + alpl0;
            r0 = src0 * rcp_fast(alp0);
            init = 0;
            if (alpl0 * ds < (double)(-0.01f)) maser = 1;
            continue;
          }
        }
        if (init) {  // first segment of the ray / after the inner hole (rare): line.F:4559-4586
          const StagedLine s0 = sl[(slot - 1) * nlc + ml];
          const double u0 = (dnu - s0.nudv) * s1.inv_aa;
          const double phi0 = 0.56419583546 * s1.inv_aa * exp_neg_tab(-(u0 * u0), s_exptab);
          src0 = s0.srcd + s0.cN * phi0;
          alp0 = s0.alpd + s0.kk * phi0;
          r0 = src0 * rcp_fast(alp0);
          init = 0;
        }
        // line.F:4554-4597 + transfer.F:1498-1571, straight line
        const double u1 = (dnu - s1.nudv) * s1.inv_aa;
        const double e1g = exp_neg_tab(-(u1 * u1), s_exptab);
        const double alpl1 = s1.K1 * e1g;
        const double src1 = fma(s1.A1, e1g, s1.srcd);
        const double alp1 = s1.alpd + alpl1;
        const double dtau = hds * (alp0 + alp1);
        const double theomax = hds * (src0 + src1);
        const double r1 = src1 * rcp_fast(alp1);
        const double xpe = exp_neg_tab(-fabs(dtau), s_exptab);
        const bool p0 = alp0 > 0.0, p1 = alp1 > 0.0;
        const double s_a = p0 ? r0 : (p1 ? r1 : 0.0);
        const double s_b = p1 ? r1 : (p0 ? r0 : 0.0);
        const bool thick = dtau > 1.e-6;
        const double e0 = 1.0 - xpe;
        const double ee1 = dtau - e0;
        const double bt = ee1 * rcp_fast(dtau);
        const double hb = 0.5 * dtau;
        const double b = thick ? bt : hb;
        const double a = thick ? (e0 - bt) : hb;
        const double x = thick ? xpe : (1.0 - dtau);
        double qv = (dtau > (double)1e-9f) ? fma(a, s_a, b * s_b) : theomax;
        qv = fmin(qv, theomax);
        inten = fma(inten, x, qv);
        src0 = src1;
        alp0 = alp1;
        r0 = r1;
        if (s1.K1 < 0.0 && alpl1 * ds < (double)(-0.01f)) maser = 1;  // telescope.F:4295
        nelem++;
      }
    }
    __syncthreads();
  }
  if (active) {
    const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
    P.img[row * P.nfr + ch] = inten;
    if (P.integ) P.integ[row * P.nfr + ch] = masked ? 1 : 2;
    if (maser) atomicOr(&P.maser[l], 1);
  }
  // work counters
  unsigned long long e = nelem, s = active ? (unsigned long long)(N > 0 ? N - 1 : 0) : 0, r = active ? 1 : 0;
  for (int o = 16; o; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if (lane == 0 && r) {
    atomicAdd(&P.counters[0], r);
    atomicAdd(&P.counters[1], e);
    atomicAdd(&P.counters[2], s);
  }
}

// the centre ray (telescope.F:498-527): one thread per (line, channel), reference-ordered scalar
// path; also yields char_tau_center
__global__ void __launch_bounds__(128) center_kernel(RenderParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool active = i < P.nl * P.nfr;
  unsigned long long e = 0, s = 0, r = 0;
  if (active) {
    const int l = i / P.nfr, ch = i % P.nfr;
    double tau;
    unsigned ne;
    int maser = 0;
    const double inten = integrate_ray_channel(P, l, 0, ch, tau, ne, maser);
    const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi;
    P.img[row * P.nfr + ch] = inten;
    if (P.integ) P.integ[row * P.nfr + ch] = 2;
    if (ch == P.nfr - 1) P.tau_center[l] = tau;
    if (maser) atomicOr(&P.maser[l], 1);
    e = ne;
    s = (unsigned long long)(P.node_off[1] - P.node_off[0] - 1);
    r = 1;
  }
  for (int o = 16; o; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if (lane == 0 && r) {
    atomicAdd(&P.counters[0], r);
    atomicAdd(&P.counters[1], e);
    atomicAdd(&P.counters[2], s);
  }
}

// thread blocks of tile_kernel per ray
__global__ void plan_kernel(RenderParams P) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray > P.nray) return;
  unsigned n = 0;
  if (ray < P.nray) {
    const unsigned M = P.item_off[(size_t)(ray + 1) * P.nl] - P.item_off[(size_t)ray * P.nl];
    n = (M + kTileThreads - 1) / kTileThreads;
  }
  P.ncta[ray] = n;
}

// continuum copy for the channels the reference skips (telescope.F:557-612); one warp per task
__global__ void __launch_bounds__(256) fill_kernel(RenderParams P) {
  const int lane = threadIdx.x & 31;
  const long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long ntask = (long long)P.nl * P.nray;
  if (task >= ntask) return;
  const int ray = (int)(task / P.nl), l = (int)(task % P.nl);
  if (ray == 0 || !P.nonredundant) return;
  const int4 rg = P.rng[task];
  const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
  double *I = P.img + row * P.nfr;
  // continuum known after channel 0 (if out of range) or after the pre-integrated channel c0
  double cont = (rg.w == 0) ? I[0] : (rg.z >= 0 ? I[rg.z] : 0.0);
  if (cont != 0.0) {
    for (int c = 1 + lane; c < P.nfr; c += 32) {
      const bool in = (c >= rg.x && c <= rg.y);
      if (!in && c != rg.z) I[c] = cont;
    }
    return;
  }
  // rare: the continuum is exactly zero, so the reference integrates every skipped channel until
  // one is non-zero (imcir_cont.ne.0 test, telescope.F:583).  Sequential, lane 0.
  if (lane == 0) {
    unsigned long long e = 0, r = 0, s = 0;
    for (int c = 1; c < P.nfr; c++) {
      const bool in = (c >= rg.x && c <= rg.y);
      if (in) continue;
      if (cont != 0.0) {
        I[c] = cont;
      } else {
        if (c != rg.z) {
          double tau;
          unsigned ne;
          int maser = 0;
          I[c] = integrate_ray_channel(P, l, ray, c, tau, ne, maser);
          if (maser) atomicOr(&P.maser[l], 1);
          e += ne;
          r += 1;
          s += (unsigned)(P.node_off[ray + 1] - P.node_off[ray] - 1);
        }
        cont = I[c];
      }
    }
    if (r) {
      atomicAdd(&P.counters[0], r);
      atomicAdd(&P.counters[1], e);
      atomicAdd(&P.counters[2], s);
    }
  }
}

// replicate the centre ray over phi (telescope.F:524-526) -- only when the cube is requested
__global__ void center_replicate_kernel(RenderParams P) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)P.nl * (P.nphi - 1) * P.nfr;
  if (i >= n) return;
  const int c = (int)(i % P.nfr);
  const int ip = 1 + (int)((i / P.nfr) % (P.nphi - 1));
  const int l = (int)(i / ((long long)P.nfr * (P.nphi - 1)));
  const size_t base = (size_t)l * (size_t)(P.nrr + 1) * P.nphi;
  P.img[(base + ip) * P.nfr + c] = P.img[base * P.nfr + c];
  if (P.integ) P.integ[(base + ip) * P.nfr + c] = 0;
}

// telescope.F:1418-1423: mean over phi of one ring times the ring area, phi in index order
__global__ void __launch_bounds__(128) ringsum_kernel(RenderParams P, const double *surf, double *ring) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)P.nl * P.nrr * P.nfr;
  if (i >= n) return;
  const int c = (int)(i % P.nfr);
  const int ir = 1 + (int)((i / P.nfr) % P.nrr);
  const int l = (int)(i / ((long long)P.nfr * P.nrr));
  const double *I = P.img + ((size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)ir * P.nphi) * P.nfr + c;
  double dslum = 0.0;
  for (int ip = 0; ip < P.nphi; ip++) dslum = dslum + I[(size_t)ip * P.nfr];
  dslum = dslum / (1.0 * P.nphi);
  dslum = dslum * surf[ir];
  ring[i] = dslum;
}

// telescope.F:1388-1433: sum over rings in index order, divide by distance^2
__global__ void __launch_bounds__(128) flux_kernel(RenderParams P, const double *surf, const double *ring,
                                                   double dist2, double *flux) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nl * P.nfr) return;
  const int c = i % P.nfr, l = i / P.nfr;
  double slum = 0.0;
  const double dslum = surf[0] * P.img[((size_t)l * (size_t)(P.nrr + 1) * P.nphi) * P.nfr + c];
  slum = slum + dslum;
  const double *rg = ring + (size_t)l * P.nrr * P.nfr + c;
  for (int ir = 0; ir < P.nrr; ir++) slum = slum + rg[(size_t)ir * P.nfr];
  flux[i] = slum / dist2;
}

// imcir_cmask is never cleared between lines (telescope.F:548,575): running OR over the lines of
// this call on top of the state left by earlier calls
__global__ void cmask_kernel(RenderParams P, unsigned char *accum, int *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)(P.nrr + 1) * P.nphi * P.nfr;
  if (i >= per) return;
  unsigned char a = accum[i];
  for (int l = 0; l < P.nl; l++) {
    if (P.integ[(size_t)l * per + i] == 1) a = 1;
    out[(size_t)l * per + i] = a;
  }
  accum[i] = a;
}

// FP64 FMA peak probe (bench.py roofline denominator): 16 independent DFMA chains per thread
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *sink, int iters) {
  double a[16];
  const double x = 1.0000001, y = 1e-9 * (threadIdx.x + 1);
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = 1.0 + k + threadIdx.x * 1e-3;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = fma(a[k], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
  if (s == 12345.678) sink[(blockIdx.x * blockDim.x + threadIdx.x) & 0xfffff] = s;
}

}  // namespace

// ---- launchers (host) ---------------------------------------------------------------------
void launch_prep(const PrepParams &P, cudaStream_t st) {
  dim3 grid((unsigned)((P.ncell + 255) / 256), (unsigned)P.nl);
  prep_cells_kernel<<<grid, 256, 0, st>>>(P);
}
void launch_span(const RenderParams &P, cudaStream_t st) {
  const long long ntask = (long long)P.nl * P.nray;
  const long long threads = ntask * 32;
  span_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P);
}
void launch_plan(const RenderParams &P, cudaStream_t st) {
  plan_kernel<<<(P.nray + 1 + 255) / 256, 256, 0, st>>>(P);
}
int tile_smem_limit() {
  static int done = 0;
  const int want = 44 * 1024;  // 4 blocks/SM of 128 threads stay resident
  if (!done) {
    cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    done = 1;
  }
  return want;
}
void launch_integrate(const RenderParams &P, unsigned total_ctas, cudaStream_t st) {
  center_kernel<<<(P.nl * P.nfr + 127) / 128, 128, 0, st>>>(P);
  if (!total_ctas) return;
  tile_kernel<<<total_ctas, kTileThreads, P.smem_budget, st>>>(P);
}
void launch_fill(const RenderParams &P, cudaStream_t st) {
  const long long ntask = (long long)P.nl * P.nray;
  const long long threads = ntask * 32;
  fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P);
}
void launch_center_replicate(const RenderParams &P, cudaStream_t st) {
  const long long n = (long long)P.nl * (P.nphi - 1) * P.nfr;
  if (n <= 0) return;
  center_replicate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P);
}
void launch_flux(const RenderParams &P, const double *surf, double *ring, double dist2, double *flux,
                 cudaStream_t st) {
  const long long n = (long long)P.nl * P.nrr * P.nfr;
  ringsum_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, surf, ring);
  flux_kernel<<<(P.nl * P.nfr + 127) / 128, 128, 0, st>>>(P, surf, ring, dist2, flux);
}
void launch_dfma_peak(double *sink, int iters, int blocks, int threads, cudaStream_t st) {
  dfma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters);
}
void launch_cmask(const RenderParams &P, unsigned char *accum, int *out, cudaStream_t st) {
  const long long per = (long long)(P.nrr + 1) * P.nphi * P.nfr;
  cmask_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(P, accum, out);
}

}  // namespace rl

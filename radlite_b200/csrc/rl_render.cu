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
  double ds, dvmu, lw, inv_lwav, wr, wt;
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
  n.inv_lwav = b.y;
  n.wr = c.x;
  n.wt = c.y;
  n.cells = d;
  n.flags = (uint32_t)d.x >> kCellFlagShift;
  n.cells.x = d.x & kCellMask;
  return n;
}

// interpolate the per-line cell record {src_dust, alp_dust, N_up, N_down} at a node
// (line.F:4054-4197; three cases by crossing type)
__device__ __forceinline__ double4 gather_line(const double4 *__restrict__ cellL, size_t nl, int4 c,
                                               double dr, double dt, int icr) {
  // cell-major layout: record of line l at cell k is cellL[k * nl + l]; cellL points at line l of cell 0
  double4 a = ldg4(cellL + (size_t)c.x * nl), o;
  if (icr == 1) {
    double4 b = ldg4(cellL + (size_t)c.y * nl);
    o.x = (1.0 - dt) * a.x + dt * b.x;
    o.y = (1.0 - dt) * a.y + dt * b.y;
    o.z = (1.0 - dt) * a.z + dt * b.z;
    o.w = (1.0 - dt) * a.w + dt * b.w;
  } else if (icr == 2) {
    double4 b = ldg4(cellL + (size_t)c.z * nl);
    o.x = (1.0 - dr) * a.x + dr * b.x;
    o.y = (1.0 - dr) * a.y + dr * b.y;
    o.z = (1.0 - dr) * a.z + dr * b.z;
    o.w = (1.0 - dr) * a.w + dr * b.w;
  } else {
    double4 b = ldg4(cellL + (size_t)c.y * nl), cc = ldg4(cellL + (size_t)c.z * nl), d = ldg4(cellL + (size_t)c.w * nl);
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
  P.cellL[(size_t)cell * P.nl + l] = o;
  P.cellD[(size_t)cell * P.nl + l] = make_double2(src, alp);
}

// ------------------------------------------------------------------------------------------
// velocity span of a ray for every line of the batch -> which channels the reference integrates
//
// minvel/maxvel of telescope.F:4265-4270 run over the start node of every segment whose
// interpolated N_up + N_down exceeds LEVTHRES.  The interpolation weights are convex, so the test is
// decided by the stencil cells alone whenever they agree: mask_kernel packs, per cell, one bit per
// line "surely above" / "surely below" the threshold (128 lines = one uint4 each); span_kernel ANDs
// the masks of a node's stencil cells and only evaluates the interpolation where they disagree.
// One block per ray: threads first act as nodes (load the record, combine the masks, park them in
// shared memory), then as lines (walk the parked nodes, min/max of Omega.v/c).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_kernel(RenderParams P) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cell = i >> 2;
  const int w = (int)(i & 3);
  if (cell >= P.ncell) return;
  const double hi = P.levthres * (1.0 + 1.0e-9), lo = P.levthres * (1.0 - 1.0e-9);
  uint32_t on = 0, off = 0;
  for (int b = 0; b < 32; b++) {
    const int l = 32 * w + b;
    if (l < P.nl) {
      const double2 zw = __ldg(reinterpret_cast<const double2 *>(P.cellL + (size_t)cell * P.nl + l) + 1);
      const double s = zw.x + zw.y;
      on |= (s > hi ? 1u : 0u) << b;
      off |= (s < lo ? 1u : 0u) << b;
    }
  }
  reinterpret_cast<uint32_t *>(&P.masks[cell].on)[w] = on;
  reinterpret_cast<uint32_t *>(&P.masks[cell].off)[w] = off;
}

__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) {
  return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w);
}

__global__ void __launch_bounds__(kSpanThreads) span_kernel(RenderParams P) {
  // Omega.v/c of the parked nodes as REAL*4: minvel / maxvel are REAL*4 in the reference (common_telescope.h:17)
  // and rounding to float is monotonic, so the float of the minimum is the minimum of the floats -- and
  // single-precision min / max are one instruction each where the double ones are compare-and-select sequences
  __shared__ float s_dv[kSpanThreads];
  __shared__ uint32_t s_on[4][kSpanThreads];
  const int ray = blockIdx.x, tid = threadIdx.x;
  const int l = tid;  // the line this thread owns in the second phase
  const long long task = (long long)ray * P.nl + l;
  {  // ring-block sharding (multi-GPU single-line configs): rays of other ranks carry no items
    const int ir = ray == 0 ? 0 : 1 + (ray - 1) / P.nphi;
    if (ir < P.ring_lo || ir > P.ring_hi) {
      if (l < P.nl) {
        P.rng[task] = make_int4(1, 0, -1, 3);
        P.nitems[task] = 0;
      }
      return;
    }
  }
  if (ray == 0 || !P.nonredundant) {
    if (l < P.nl) {
      P.rng[task] = make_int4(1, P.nfr - 1, -1, ray == 0 ? 2 : 1);
      // the centre ray is traced by center_kernel (it also produces char_tau_center)
      P.nitems[task] = ray == 0 ? 0 : P.nfr;
    }
    return;
  }
  const long long n0 = P.node_off[ray], n1 = P.node_off[ray + 1];
  const int w = (l >> 5) & 3;
  const uint32_t bit = 1u << (l & 31);
  float vmin = 1.0f, vmax = -1.0f;  // initial values of telescope.F:388-391
  // telescope.F:4265-4270: start node of every segment, i.e. all nodes but the last
  for (long long c = n0; c < n1 - 1; c += kSpanThreads) {
    const int cnt = (int)min((long long)kSpanThreads, n1 - 1 - c);
    if (tid < cnt) {
      const Node nd = load_node(P.nodes.rec, c + tid);
      const int icr = nd.flags & kFlagIcrMask;
      const CellMask *mk = P.masks;
      uint4 on = __ldg(&mk[nd.cells.x].on), off = __ldg(&mk[nd.cells.x].off);
      if (icr != 2) {
        on = and4(on, __ldg(&mk[nd.cells.y].on));
        off = and4(off, __ldg(&mk[nd.cells.y].off));
      }
      if (icr != 1) {
        on = and4(on, __ldg(&mk[nd.cells.z].on));
        off = and4(off, __ldg(&mk[nd.cells.z].off));
      }
      if (icr == 3) {
        on = and4(on, __ldg(&mk[nd.cells.w].on));
        off = and4(off, __ldg(&mk[nd.cells.w].off));
      }
      // lines whose stencil cells disagree (or sit on the threshold): the interpolation decides.  Resolved here,
      // by the thread that owns the node -- the 128 nodes of the chunk in parallel -- so that the walk over the
      // parked nodes below touches shared memory only
      uint32_t o4[4] = {on.x, on.y, on.z, on.w};
      const uint32_t f4[4] = {off.x, off.y, off.z, off.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int nk = min(32, max(0, P.nl - 32 * k));  // lines of this word
        uint32_t un = ~(o4[k] | f4[k]) & (nk >= 32 ? 0xffffffffu : ((1u << nk) - 1u));
        while (un) {
          const int b = __ffs((int)un) - 1;
          un &= un - 1u;
          const double4 v = gather_line(P.cellL + (32 * k + b), (size_t)P.nl, nd.cells, nd.wr, nd.wt, icr);
          if (v.z + v.w > P.levthres) o4[k] |= 1u << b;
        }
        s_on[k][tid] = o4[k];
      }
      s_dv[tid] = (float)nd.dvmu;
    }
    __syncthreads();
    if (l < P.nl) {
#pragma unroll 4
      for (int t = 0; t < cnt; t++) {
        if (s_on[w][t] & bit) {
          const float dv = s_dv[t];
          vmin = fminf(vmin, dv);
          vmax = fmaxf(vmax, dv);
        }
      }
    }
    __syncthreads();
  }
  if (l >= P.nl) return;
  // REAL*4 minvel/maxvel (common_telescope.h:17), initial values 1 and -1 (telescope.F:388-391)
  const float minvel = vmin, maxvel = vmax;
  const double hi_lim = (double)maxvel + 2.f * P.aksmax_c;
  const double lo_lim = (double)minvel - 2.f * P.aksmax_c;
  const double *velo = P.velo + (size_t)l * P.nfr;
  // channel 0 (reference inu=1) is always integrated; 1..nfr-1 only inside [lo_lim, hi_lim]
  int lo = P.nfr, hi = -1, c0 = P.nfr;
  for (int ch = 1; ch < P.nfr; ch++) {
    const double v = __ldg(&velo[ch]);
    const bool in = (v <= hi_lim) && (v >= lo_lim);
    if (in) {
      lo = min(lo, ch);
      hi = max(hi, ch);
    } else {
      c0 = min(c0, ch);
    }
  }
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

// the j-th channel the reference integrates for a task (rng = {lo, hi, c0, kind})
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
  const double4 *__restrict__ cellL = P.cellL + l;
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
  double4 v0 = gather_line(cellL, (size_t)P.nl, nd.cells, nd.wr, nd.wt, nd.flags & kFlagIcrMask);
  double dvmu0 = nd.dvmu, lw0 = nd.lw;
  for (long long i = n0 + 1; i < n1; i++) {
    nd = load_node(P.nodes.rec, i);
    const uint32_t fl = nd.flags;
    const double ds = nd.ds, dvmu1 = nd.dvmu, lw1 = nd.lw;
    const double4 v1 = gather_line(cellL, (size_t)P.nl, nd.cells, nd.wr, nd.wt, fl & kFlagIcrMask);
    if (fl & (kFlagInit | kFlagStar | kFlagZero)) {
      if (fl & kFlagZero) inten = 0.0;
      if (fl & kFlagStar)
        inten = (1.0 - P.starfract) * inten + P.starfract * P.star_line[(size_t)l * P.nfr + ch];
      k.init = 1;
    }
    bool done = false;
    if (P.subgrid && (fl & kFlagSub)) {  // line.F:4706-4715 (trigger evaluated by geom_kernel)
      const double q = fabs((dvmu1 - dvmu0) / (0.5 * (lw0 + lw1) / 2.99792458e5));
      const double s_c = ds * (velo_ch - dvmu0) / (dvmu1 - dvmu0);
      const double dls = ds / q;
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

// ------------------------------------------------------------------------------------------
// tile_kernel: the formal solution for one ray x a tile of (line, channel) items.
//
// One thread block owns one camera ray and a tile of consecutive entries of the ray's item list (all
// channels the reference integrates, lines in index order: "ray x line tile"), one item per consumer
// thread.  A producer warp walks the ray's nodes in chunks: it gathers and interpolates the cell values
// of every (node, line of the tile) pair and precomputes every channel-independent constant of the line
// profile -- that work is done once per ray, line and node instead of once per channel -- and stages
// the chunk in a ring of kTileBufs shared-memory buffers.  The consumer warps never touch global
// memory inside the ray loop.  Hand-over by mbarriers: full[b] (producer -> consumers) and empty[b]
// (each consumer warp -> producer), so consumer warps are not coupled to each other.
//
// All threads of a block walk the same ray: no trip-count divergence.  Segments are integrated in
// groups of RL_JAM whose dependency chains interleave (jam_steps); the case split of transfer.F:1517,
// 1542 is a warp vote per group:
//   * every lane has dtau <= 1e-9 on every segment of the group (about 2/3 of all steps in disk
//     atmospheres): I <- I (1-dtau) + theomax
//   * otherwise the full qdr_src_2 step with e^-dtau and the source-function ratios
//   * nodes flagged by the geometry (first segment, inner hole / star mixing, 6q > 1 sub-grid
//     candidates) take a separate reference-ordered path (slow_step).
//
// Arithmetic notes (all deviations from the reference are << the 1e-6 tolerance):
//   exp(x), x <= 0: n = round(x kTabN/ln2), exp = 2^(n>>kTabBits) T1[n & (kTabN-1)] P(r), with one small
//   shared-memory table and |r| <= ln2/(2 kTabN); P = degree kDegGauss for the line profile (< 1.5e-13),
//   kDegTau for exp(-dtau) (< 2e-18).  The Gaussian argument is carried pre-scaled so that n and r fall
//   out of two FMAs.  Profile values below exp(-345) are flushed to 0.  Divisions: hardware reciprocal
//   seed + two Newton steps (< 2e-11).
// ------------------------------------------------------------------------------------------
constexpr double kExpMagic = 6755399441055744.0;       // 1.5 * 2^52
constexpr double kLn2 = 0.6931471805599453;
constexpr double kLog2eS = (double)kTabN / kLn2;       // kTabN / ln 2
constexpr double kLn2S = kLn2 / (double)kTabN;         // L = ln 2 / kTabN
constexpr double kCnorm = 0.56419583546 / kTabSqrtScale;
// Taylor coefficients L^k / k! of exp(L rp), |rp| <= 1/2
template <int K>
__host__ __device__ constexpr double exp_coef() {
  return exp_coef<K - 1>() * kLn2S / (double)K;
}
template <>
__host__ __device__ constexpr double exp_coef<0>() {
  return 1.0;
}
// the same coefficients in constant memory: an FP64 instruction takes one operand straight from a constant bank,
// while a 64-bit literal with a non-zero low word costs two UMOVs every time it is used (14 per three exp(-dtau))
__constant__ double c_expc[9] = {exp_coef<0>(), exp_coef<1>(), exp_coef<2>(), exp_coef<3>(), exp_coef<4>(),
                                 exp_coef<5>(), exp_coef<6>(), exp_coef<7>(), exp_coef<8>()};
template <int K>
struct Horner {
  static __device__ __forceinline__ double run(double p, double rp) {
    return Horner<K - 1>::run(fma(p, rp, c_expc[K]), rp);
  }
};
template <>
struct Horner<-1> {
  static __device__ __forceinline__ double run(double p, double) { return p; }
};
constexpr unsigned kHiTauMax = 0x40859000u;            // hi word of 690.0
constexpr double kAlpTiny = 1.0e-280;                  // alpha <= this is treated like alpha <= 0

struct __align__(16) HotLine {  // per (node, line): read by every step
  double srcd, alpd;  // dust source / opacity at the node (line.F:4058-4063)
  double A1, K1;      // c_src N_up and c_alp (N_down B_du - N_up B_ud), times the profile norm of the
                      // segment ending here
  double ia, nv;      // scaled reciprocal Doppler width of that segment ; nu0 Omega.v/c times ia
};
struct __align__(16) ColdLine {  // per (node, line): flagged nodes only
  double cN, kk;      // c_src N_up ; c_alp (N_down B_du - N_up B_ud)
};
struct __align__(16) HotNode {
  double hds;         // ds / 2
  uint32_t flags, pad;
};
struct __align__(16) ColdNode {
  double ds, dvmu, lwav, pad;
};
// shared-window byte addresses of one stage buffer (32-bit: the hot loop steps them directly)
struct TileBuf {
  uint32_t hn, cn, hl, cl;
};
template <class T>
__device__ __forceinline__ T *smem_ptr(uint32_t a) {
  return reinterpret_cast<T *>(__cvta_shared_to_generic(a));
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds_u32x2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
constexpr int kSlotBytes = (int)(sizeof(HotNode) + sizeof(ColdNode));
constexpr int kPairBytes = (int)(sizeof(HotLine) + sizeof(ColdLine));

// The case thresholds of qdr_src_2 in the integrate kernels: dtau > 1e-9 (REAL literal, transfer.F:1542) and
// dtau > 1e-6 (transfer.F:1517) are tested on the high words of the bit patterns -- one 32-bit integer compare,
// false for negative values, off the FP64 pipe.  That moves each threshold up by at most 2^-20 of itself (a value
// in that sliver takes the neighbouring branch, whose result differs by O(dtau^2) of an already negligible term);
// every integrate kernel uses these same predicates, so a lane's result does not depend on the kernel or on the
// outcome of a warp vote.  (The reference-ordered scalar path -- centre ray, rectangular imager -- keeps the
// reference's compares.)
constexpr int kThinHi = 0x3E112E0B;  // high word of (double)1e-9f = 0x3E112E0BE0000000
constexpr int kMidHi = 0x3EB0C6F7;   // high word of 1e-6 = 0x3EB0C6F7A0B5ED8D
__device__ __forceinline__ bool gt_thin(double d) { return __double2hiint(d) > kThinHi; }
__device__ __forceinline__ bool gt_mid(double d) { return __double2hiint(d) > kMidHi; }
// min of two non-NaN-or-rescued values: a NaN in a (0/0 of a degenerate step) selects b like fmin does, without
// fmin's NaN-quieting instruction sequence
__device__ __forceinline__ double min_sel(double a, double b) { return a < b ? a : b; }

// n/x: hardware reciprocal seed (MUFU.RCP64H, about 2^-9) and two Newton steps, the second fused
// with the multiplication: relative error ~ seed^4 < 2e-11
__device__ __forceinline__ double div_fast(double n, double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  const double q = n * y;
  return fma(q, e, q);
}

// 2^(n / kTabN) e^r, r = rp ln2/kTabN, |rp| <= 1/2: one table read of 2^(j/kTabN) (a single 8-byte
// shared-memory read per exp: the LSU data pipe, not the FP64 pipe, is the scarce resource) and a short
// polynomial of degree DEG (relative error r^(DEG+1)/(DEG+1)!).  exp(-dtau) needs the longer one: it
// feeds the cancelling differences e0 = 1 - xp, e1 = dtau - e0 of transfer.F:1519-1520, whose error is
// the absolute error of xp.  T1 = shared-window address of the table.
template <int DEG>
__device__ __forceinline__ double exp_tab(double t, double rp, uint32_t T1) {
  const int n = __double2loint(t);
  const double v = lds_f64(T1 + ((n & (kTabN - 1)) << 3));
  const double p = Horner<DEG - 1>::run(c_expc[DEG], rp);
  const int hi = __double2hiint(v) + ((n >> kTabBits) << 20);
  return __hiloint2double(hi, __double2loint(v)) * p;
}
// exp(-u^2 ln2/kTabN) for the pre-scaled argument u; exactly 0 beyond exp(-345)
__device__ __forceinline__ double gauss_tab(double u, uint32_t T1, uint32_t T2) {
  const double t = fma(-u, u, kExpMagic);
  const double fn = t - kExpMagic;
  const double rp = fma(-u, u, -fn);
  const double e = exp_tab<kDegGauss>(t, rp, T1);
  const bool far = ((unsigned)__double2hiint(u) & 0x7fffffffu) > kHiUmax;
  return far ? 0.0 : e;
}
// exp(-d), d >= 0 (clamped at ~690; lanes with d < 0 get an unused finite value)
__device__ __forceinline__ double expneg_tab(double d, uint32_t T1, uint32_t T2) {
  const unsigned h = min((unsigned)__double2hiint(d), kHiTauMax);
  const double dc = __hiloint2double((int)h, __double2loint(d));
  const double t = fma(dc, -kLog2eS, kExpMagic);
  const double fn = t - kExpMagic;
  const double rp = fma(dc, -kLog2eS, -fn);
  return exp_tab<kDegTau>(t, rp, T1);
}

// the qdr_src_2 step (transfer.F:1498-1571) with all case selections branch free: r0 = src0/alp0 comes
// in, r1 = src1/alp1 goes out; the step itself is I <- I x + qv
__device__ __forceinline__ void step_coeffs(double alp0, double r0, double src1, double alp1, double &r1,
                                            double dtau, double theomax, uint32_t T1, double &x, double &qv) {
  r1 = div_fast(src1, alp1);
  const double xpe = expneg_tab(dtau, T1, 0);
  const double e0 = 1.0 - xpe;
  const double ee1 = dtau - e0;
  const double bt = div_fast(ee1, dtau);
  const double hb = 0.5 * dtau;
  const bool thick = gt_mid(dtau);
  const double b = thick ? bt : hb;
  const double a = thick ? (e0 - bt) : hb;
  x = thick ? xpe : (1.0 - dtau);
  const bool p0 = alp0 > kAlpTiny, p1 = alp1 > kAlpTiny;
  const double s_a = p0 ? r0 : (p1 ? r1 : 0.0);
  const double s_b = p1 ? r1 : (p0 ? r0 : 0.0);
  qv = fma(a, s_a, b * s_b);
  qv = gt_thin(dtau) ? min_sel(qv, theomax) : theomax;
}

__device__ __forceinline__ void full_step(double &inten, double alp0, double r0, double src1, double alp1,
                                          double &r1, double dtau, double theomax, uint32_t T1, uint32_t) {
  double x, qv;
  step_coeffs(alp0, r0, src1, alp1, r1, dtau, theomax, T1, x, qv);
  inten = fma(inten, x, qv);
}

// The same step with ONE division: a S_a + b S_b (a = e0 - b, b = e1 / dtau on the thick branch, a = b = dtau / 2
// on the thin one; S = j / alpha of the node itself or, where its opacity is not positive, of the other node:
// transfer.F:1517-1541) brought over the common denominator.  Used by ztile_kernel and zcont_kernel for every
// step that is not all-thin; a lane with dtau > 1e-6 and both opacities positive performs exactly the
// operations of ztile_kernel's select-free variant, so its result does not depend on which of the two the
// warp's vote picks.
constexpr double kAlpMin = 1.0e-150;  // alpha <= this is treated like alpha <= 0 (the product of two opacities and
                                      // dtau must stay a normal number)
__device__ __forceinline__ void step_onediv(double alp0, double src0, double alp1, double src1, double dtau,
                                            double theomax, uint32_t T1, double &x, double &qv) {
  const double xpe = expneg_tab(dtau, T1, 0);
  const double e0 = 1.0 - xpe, e1 = dtau - e0;
  const bool thick = gt_mid(dtau);
  const bool p0 = alp0 > kAlpMin, p1 = alp1 > kAlpMin;
  const double nA = p0 ? src0 : (p1 ? src1 : 0.0), dA = p0 ? alp0 : (p1 ? alp1 : 1.0);
  const double nB = p1 ? src1 : (p0 ? src0 : 0.0), dB = p1 ? alp1 : (p0 ? alp0 : 1.0);
  const double hb = 0.5 * dtau;
  const double ca = thick ? fma(e0, dtau, -e1) : hb, cb = thick ? e1 : hb, dd = thick ? dtau : 1.0;
  const double den = dd * (dA * dB);
  const double num = fma(ca, nA * dB, cb * (nB * dA));
  x = thick ? xpe : (1.0 - dtau);
  qv = div_fast(num, den);
  qv = gt_thin(dtau) ? min_sel(qv, theomax) : theomax;
}

__shared__ double s_T1[kTabN];       // 2^(j/kTabN): the exp table every integrate kernel fills at block start

// Sub-gridded segment in the staged (cN, kk) form (line.F:4745-4833): the sub-points and their order are the
// reference's; each sub-step uses the kernels' own exp (table + polynomial) and the one-division qdr_src_2 step,
// like every other step of the integrate kernels (a sub-gridded segment is up to 32 of them).  sub_point is
// the one place a sub-point is evaluated.
struct SubCtx {
  double ds, sd0, ad0, cN0, kk0, dv0, sd1, ad1, cN1, kk1, dv1;
  double un, uv, norm;  // dnu and nu0 times the scaled reciprocal Doppler width ; profile norm
};
__device__ __forceinline__ SubCtx sub_ctx(double nu0, double k_aa, double dnu_ch, double ds, double sd0, double ad0,
                                          double cN0, double kk0, double dv0, double sd1, double ad1, double cN1,
                                          double kk1, double dv1, double lw) {
  SubCtx k;
  k.ds = ds;
  k.sd0 = sd0; k.ad0 = ad0; k.cN0 = cN0; k.kk0 = kk0; k.dv0 = dv0;
  k.sd1 = sd1; k.ad1 = ad1; k.cN1 = cN1; k.kk1 = kk1; k.dv1 = dv1;
  const double aa = k_aa * (0.5 * (lw + lw));
  k.norm = 0.56419583546 / aa;
  const double ias = kTabSqrtScale / aa;  // pre-scaled reciprocal Doppler width (gauss_tab's argument scale)
  k.un = dnu_ch * ias;
  k.uv = nu0 * ias;
  return k;
}
__device__ __forceinline__ double lerp_rn(double a, double b, double w, double w1);
// dust pair and line terms at distance s from the segment's start node (endpoint: the end node itself)
__device__ __forceinline__ void sub_point(const SubCtx &k, double s, bool endpoint, uint32_t T1, double &sd,
                                          double &ad, double &srcl, double &alpl) {
  const double eps = s / k.ds, epsp = 1.0 - eps;
  const double cN = endpoint ? k.cN1 : lerp_rn(k.cN0, k.cN1, eps, epsp);
  const double kk = endpoint ? k.kk1 : lerp_rn(k.kk0, k.kk1, eps, epsp);
  const double dv = endpoint ? k.dv1 : lerp_rn(k.dv0, k.dv1, eps, epsp);
  sd = endpoint ? k.sd1 : lerp_rn(k.sd0, k.sd1, eps, epsp);
  ad = endpoint ? k.ad1 : lerp_rn(k.ad0, k.ad1, eps, epsp);
  const double phi = k.norm * gauss_tab(fma(-k.uv, dv, k.un), T1, 0);
  srcl = cN * phi;
  alpl = kk * phi;
}
__device__ __noinline__ int subgrid_tile(double nu0, double k_aa, double dnu_ch, double &inten, double ds,
                                         double sleft, double sright, double sd0, double ad0, double cN0,
                                         double kk0, double dv0, double sd1, double ad1, double cN1,
                                         double kk1, double dv1, double lw, double &srcl0, double &alpl0,
                                         int init) {
  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1);
  const double lg_ds = (sright - sleft) / (kLgNrMax - 1.0);
  const SubCtx k = sub_ctx(nu0, k_aa, dnu_ch, ds, sd0, ad0, cN0, kk0, dv0, sd1, ad1, cN1, kk1, dv1, lw);
  if (init) {  // the start point with this segment's width (line.F:4559-4586)
    const double phi0 = k.norm * gauss_tab(fma(-k.uv, dv0, k.un), T1, 0);
    srcl0 = cN0 * phi0;
    alpl0 = kk0 * phi0;
  }
  // which of the 31 sub-points lie inside the segment: a run jlo .. jlo+cnt-1 (s grows with j).  Found by a
  // branch-free scan, so that the lanes of a warp then walk THEIR sub-points in step (k-th point of every lane
  // together) instead of all 31 indices with the lanes whose point is outside idling
  int jlo = 0, cnt = 0;
#pragma unroll 1
  for (int j = 1; j <= kLgNrMax; j++) {
    const double s = fma((double)(j - 1), lg_ds, sleft);
    const bool in = s > 0.0 && s < ds;
    jlo = (in && cnt == 0) ? j : jlo;
    cnt += in ? 1 : 0;
  }
  double sp = 0.0, sd_p = sd0, ad_p = ad0;
  for (int kk = 0; kk <= cnt; kk++) {
    const bool endpoint = kk == cnt;
    const double s = endpoint ? ds : fma((double)(jlo + kk - 1), lg_ds, sleft);
    double sd_c, ad_c, srcl1, alpl1;
    sub_point(k, s, endpoint, T1, sd_c, ad_c, srcl1, alpl1);
    const double src0 = sd_p + srcl0, alp0 = ad_p + alpl0, src1 = sd_c + srcl1, alp1 = ad_c + alpl1;
    const double hds = 0.5 * (s - sp);
    double x, q;
    step_onediv(alp0, src0, alp1, src1, hds * (alp0 + alp1), hds * (src0 + src1), T1, x, q);
    inten = fma(inten, x, q);
    srcl0 = srcl1;
    alpl0 = alpl1;
    sp = s; sd_p = sd_c; ad_p = ad_c;
  }
  const int n = cnt + 1;
  return n;
}
// per-block tables and per-item metadata of tile_kernel (file scope: the out-of-line slow path uses
// them too)
__shared__ int2 s_meta[128];      // {line slot, channel | cmask bit} of the thread's item
__shared__ unsigned s_flags[128]; // maser | extra elements << 8 of the thread's item
__shared__ double s_dnu[128];     // line_dnu of the thread's item

// per-item state of tile_kernel: kept in registers across the whole ray
struct Item {
  double inten, src0, alp0, r0;
};

// one flagged segment of one item (first segment, inner hole / star, sub-grid candidate), out of
// line.  st = {inten, src0, alp0, r0} in/out; returns maser | extra_elements << 1
__device__ __noinline__ unsigned slow_step(double *st, TileBuf B, int nlc, int slot, int l, int ml, int ch,
                                           double dnu, const LineDev *__restrict__ lines,
                                           const double *__restrict__ star_line, int nfr, double starfract) {
  double inten = st[0], src0 = st[1], alp0 = st[2], r0;
  unsigned ret = 0;
  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1), T2 = 0;
  const uint32_t fl = smem_ptr<HotNode>(B.hn)[slot].flags;
  const ColdNode c1 = smem_ptr<ColdNode>(B.cn)[slot], c0 = smem_ptr<ColdNode>(B.cn)[slot - 1];
  const HotLine h1 = smem_ptr<HotLine>(B.hl)[slot * nlc + ml], h0 = smem_ptr<HotLine>(B.hl)[(slot - 1) * nlc + ml];
  const ColdLine k1 = smem_ptr<ColdLine>(B.cl)[slot * nlc + ml], k0 = smem_ptr<ColdLine>(B.cl)[(slot - 1) * nlc + ml];
  const double nu0 = __ldg(&lines[l].nu0);
  const double ds = c1.ds;
  int init = 0;
  if (fl & (kFlagInit | kFlagStar | kFlagZero)) {
    if (fl & kFlagZero) inten = 0.0;
    if (fl & kFlagStar) inten = (1.0 - starfract) * inten + starfract * star_line[(size_t)l * nfr + ch];
    init = 1;
  }
  bool done = false;
  if (fl & kFlagSub) {  // line.F:4706-4745
    const double q = fabs((c1.dvmu - c0.dvmu) / (c1.lwav / 2.99792458e5));
    const double s_c = ds * (dnu * __ldg(&lines[l].inv_nu0) - c0.dvmu) / (c1.dvmu - c0.dvmu);
    const double dls3 = 3.0 * (ds / q);
    const double sright = s_c + dls3, sleft = s_c - dls3;
    if (sright > 0.0 && sleft < ds) {
      double srcl0 = src0 - h0.srcd, alpl0 = alp0 - h0.alpd;
      const int n = subgrid_tile(nu0, __ldg(&lines[l].k_aa), dnu, inten, ds, sleft, sright, h0.srcd, h0.alpd,
                                 k0.cN, k0.kk, c0.dvmu, h1.srcd, h1.alpd, k1.cN, k1.kk, c1.dvmu, c1.lwav,
                                 srcl0, alpl0, init);
      src0 = h1.srcd + srcl0;
      alp0 = h1.alpd + alpl0;
      r0 = div_fast(src0, alp0);
      if (alpl0 * ds < (double)(-0.01f)) ret |= 1u;
      ret += (unsigned)(n - 1) << 1;
      done = true;
    }
  }
  if (!done) {
    if (init) {  // line.F:4559-4586: the start point with this segment's width
      const double u0 = fma(dnu, h1.ia, -((nu0 * c0.dvmu) * h1.ia));
      const double phi0 = (kCnorm * h1.ia) * gauss_tab(u0, T1, T2);
      src0 = h0.srcd + k0.cN * phi0;
      alp0 = h0.alpd + k0.kk * phi0;
    }
    r0 = div_fast(src0, alp0);
    const double u1 = fma(dnu, h1.ia, -h1.nv);
    const double e = gauss_tab(u1, T1, T2);
    const double alpl1 = h1.K1 * e;
    const double src1 = fma(h1.A1, e, h1.srcd);
    const double alp1 = h1.alpd + alpl1;
    const double hds = 0.5 * ds;
    const double dtau = hds * (alp0 + alp1), theomax = hds * (src0 + src1);
    double r1;
    full_step(inten, alp0, r0, src1, alp1, r1, dtau, theomax, T1, T2);
    src0 = src1;
    alp0 = alp1;
    r0 = r1;
    if (alpl1 * ds < (double)(-0.01f)) ret |= 1u;
  }
  st[0] = inten;
  st[1] = src0;
  st[2] = alp0;
  st[3] = r0;
  return ret;
}
// named barrier (0 is __syncthreads): only used among the producer warps when there are several
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// mbarrier hand-over of the staged chunks: full[b] is signalled by the producer warp(s),
// empty[b] by every consumer warp on its own, so the consumer warps of a block are NOT coupled to each
// other -- each runs ahead as far as the ring allows.
__shared__ unsigned long long s_mbar[8];  // full[0..kTileBufs), empty[0..kTileBufs)
__device__ __forceinline__ void mbar_init(uint32_t a, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n}" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nRL_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RL_MBAR_DONE;\nbra RL_MBAR_WAIT;\nRL_MBAR_DONE:\n}" ::"r"(a), "r"(parity)
      : "memory");
}
constexpr int kTileBufs = 3;        // staged chunks in flight
#ifndef RL_PRODUCER_WARPS
#define RL_PRODUCER_WARPS 1
#endif
#ifndef RL_PAIRS_IN_FLIGHT
#define RL_PAIRS_IN_FLIGHT 2
#endif
constexpr int kProducerWarps = RL_PRODUCER_WARPS;      // staging warps per block
constexpr int kProducerThreads = 32 * kProducerWarps;
constexpr int kPairsInFlight = RL_PAIRS_IN_FLIGHT;     // (node, line) pairs a producer lane gathers at a time
constexpr int kBarProducers = 15;                      // named barrier among the producer warps

// ---- producer side of tile_kernel -----------------------------------------------------------
// per-tile line constants and the node records of the chunk being staged / the next one
struct __align__(16) LineC {
  double c_src, c_alp, bud, bdu, nu0, kia;
};
constexpr int kMaxTileLines = 64;
__shared__ LineC s_linec[kMaxTileLines];
__shared__ NodeRec s_nodes[2][32];

// (node, line) pair staging.  Nodes on a grid line (crossing types 1 and 2) interpolate between two
// stencil cells, extra points (type 3) between four (line.F:4054-4197, same expressions as gather_line).
// The two kinds are staged in separate, compacted passes so that the four-cell arithmetic never runs
// in a mostly idle warp.
__device__ __forceinline__ int node_icr(const NodeRec *nd) {
  return (int)(((uint32_t)nd->cells.x >> kCellFlagShift) & kFlagIcrMask);
}
// (explicitly rounded products + one fma per component: the same bits at every call site, whatever the
// compiler's contraction choices -- the opaque-wall start evaluates a node in the prologue of ztile_kernel
// that the full walk evaluates in its loop body)
__device__ __forceinline__ double lerp_rn(double a, double b, double w, double w1) {
  return fma(w, b, __dmul_rn(w1, a));
}
__device__ __forceinline__ double4 interp2(const double4 a, const double4 b, double w) {
  const double w1 = 1.0 - w;
  double4 o;
  o.x = lerp_rn(a.x, b.x, w, w1);
  o.y = lerp_rn(a.y, b.y, w, w1);
  o.z = lerp_rn(a.z, b.z, w, w1);
  o.w = lerp_rn(a.w, b.w, w, w1);
  return o;
}
__device__ __forceinline__ double4 interp4(const double4 a, const double4 b, const double4 c, const double4 d,
                                           double dr, double dt) {
  const double r1 = 1.0 - dr, t1 = 1.0 - dt;
  double4 o;
  o.x = lerp_rn(lerp_rn(a.x, b.x, dt, t1), lerp_rn(c.x, d.x, dt, t1), dr, r1);
  o.y = lerp_rn(lerp_rn(a.y, b.y, dt, t1), lerp_rn(c.y, d.y, dt, t1), dr, r1);
  o.z = lerp_rn(lerp_rn(a.z, b.z, dt, t1), lerp_rn(c.z, d.z, dt, t1), dr, r1);
  o.w = lerp_rn(lerp_rn(a.w, b.w, dt, t1), lerp_rn(c.w, d.w, dt, t1), dr, r1);
  return o;
}
__device__ __forceinline__ void pair_store(const TileBuf &B, int p, int m, const NodeRec &nd, const double4 v) {
  const LineC lc = s_linec[m];
  ColdLine c;
  c.cN = lc.c_src * v.z;
  c.kk = lc.c_alp * (v.w * lc.bdu - v.z * lc.bud);
  HotLine h;
  h.srcd = v.x;
  h.alpd = v.y;
  h.ia = nd.inv_lwav * lc.kia;
  h.nv = (lc.nu0 * nd.dvmu) * h.ia;
  const double norm = kCnorm * h.ia;
  h.A1 = c.cN * norm;
  h.K1 = c.kk * norm;
  smem_ptr<HotLine>(B.hl)[p] = h;
  smem_ptr<ColdLine>(B.cl)[p] = c;
}
// per-node part of a staged slot (one lane per slot)
__device__ __forceinline__ void node_store(const TileBuf &B, int slot, int node, const NodeRec *chunk_nodes,
                                           int subgrid) {
  const NodeRec &nd = chunk_nodes[slot];
  HotNode a;
  a.hds = 0.5 * nd.ds;
  uint32_t fl = ((uint32_t)nd.cells.x >> kCellFlagShift) & ~kFlagIcrMask;
  if (!subgrid) fl &= ~kFlagSub;
  if (node == 1) fl |= kFlagInit;  // first segment of the ray: nothing carried yet
  a.flags = fl;
  a.pad = 0;
  smem_ptr<HotNode>(B.hn)[slot] = a;
  ColdNode b;
  // mean width of the segment ending here (slot 0 only serves as a start point: never read)
  const double lw_prev = (slot > 0) ? chunk_nodes[slot - 1].lw : nd.lw;
  b.ds = nd.ds;
  b.dvmu = nd.dvmu;
  b.lwav = 0.5 * (lw_prev + nd.lw);
  b.pad = 0.0;
  smem_ptr<ColdNode>(B.cn)[slot] = b;
}

// the producer warp: stages chunk after chunk (nodes c0-1 .. c0+cnt-1 -> slots 0 .. cnt) into the
// buffer ring.  Node records are fetched one chunk ahead (one per lane, coalesced) and parked in
// shared memory; every lane then handles four (node, line) pairs at a time with all their gathers
// in flight together.
__device__ __noinline__ void producer_loop(const RenderParams &P, uint32_t smem0, int bufbytes, int nch,
                                           int nlc, int l0, long long n0, int N, int nchunks, int plane,
                                           int nall) {
  constexpr int NP = kProducerThreads;
  for (int m = plane; m < nlc; m += NP) {
    const LineDev *Lm = P.lines + (l0 + m);
    LineC lc;
    lc.c_src = __ldg(&Lm->c_src); lc.c_alp = __ldg(&Lm->c_alp); lc.bud = __ldg(&Lm->bud);
    lc.bdu = __ldg(&Lm->bdu); lc.nu0 = __ldg(&Lm->nu0); lc.kia = __ldg(&Lm->kia);
    s_linec[m] = lc;
  }
  const NodeRec *rec = P.nodes.rec + n0;
  // node records travel global -> shared without passing through registers (cp.async, 4 x 16 B)
  auto fetch_nodes = [&](int buf, int first, int count) {
    if (plane < count) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_nodes[buf][plane]);
      const char *src = reinterpret_cast<const char *>(rec + first + plane);
#pragma unroll
      for (int k = 0; k < 4; k++)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * k), "l"(src + 16 * k) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto producers_sync = [&]() {
    if (kProducerWarps == 1) __syncwarp();
    else bar_sync(kBarProducers, NP);
  };
  // q / nlc by multiplication: exact for q < 2^11, nlc <= 64 (q (M nlc - 2^20) < 2^17 < 2^20)
  const uint32_t Mdiv = ((1u << 20) + (uint32_t)nlc - 1u) / (uint32_t)nlc;
  fetch_nodes(0, 0, min(nch, N - 1) + 1);
  for (int c = 0, b = 0; c < nchunks; c++, b = (b + 1 == kTileBufs) ? 0 : b + 1) {
    const int c0 = 1 + c * nch, cnt = min(nch, N - c0);
    // this chunk's node records have landed; the next chunk's go in flight while this one is staged
    asm volatile("cp.async.wait_all;" ::: "memory");
    producers_sync();
    const int c0n = c0 + nch;
    if (c + 1 < nchunks) fetch_nodes((c + 1) & 1, c0n - 1, min(nch, N - c0n) + 1);
    if (c >= kTileBufs)  // empty[b]: the consumers are done with the chunk staged here kTileBufs chunks ago
      mbar_wait((uint32_t)__cvta_generic_to_shared(&s_mbar[kTileBufs + b]), (uint32_t)((c / kTileBufs - 1) & 1));
    TileBuf B;
    B.hn = smem0 + (uint32_t)(b * bufbytes);
    B.hl = B.hn + (uint32_t)((nch + 2) * sizeof(HotNode));
    B.cn = B.hl + (uint32_t)((nch + 2) * nlc * sizeof(HotLine));
    B.cl = B.cn + (uint32_t)((nch + 1) * sizeof(ColdNode));
    const NodeRec *cn = s_nodes[c & 1];
    const int npair = (cnt + 1) * nlc;
    const size_t nl = (size_t)P.nl;
    const double4 *cell0 = P.cellL + l0;
    // per-node part, one lane per slot; which slots are extra points (four-cell stencil)
    const int lane = plane & 31;
    const bool mine = lane <= cnt;
    const uint32_t mask3 = __ballot_sync(0xffffffffu, mine && node_icr(cn + lane) == 3);
    if (plane <= cnt) node_store(B, plane, c0 - 1 + plane, cn, P.subgrid);
    // pass 1: pairs on grid lines.  kPairsInFlight pairs per lane and round with their gathers in flight
    // together; consecutive lanes take consecutive lines of one node: with the cell-major layout their
    // records are contiguous
    for (int p = plane; p < npair; p += kPairsInFlight * NP) {
      int slot[kPairsInFlight], m[kPairsInFlight];
      bool has[kPairsInFlight];
      double4 a[kPairsInFlight], b[kPairsInFlight];
#pragma unroll
      for (int k = 0; k < kPairsInFlight; k++) {
        const int q = p + NP * k;
        slot[k] = (q < npair) ? (int)(((uint32_t)q * Mdiv) >> 20) : 0;
        m[k] = q - slot[k] * nlc;
        has[k] = (q < npair) && !((mask3 >> slot[k]) & 1u);
        if (has[k]) {
          const NodeRec *nd = cn + slot[k];
          const int4 cl = nd->cells;
          a[k] = ldg4(cell0 + m[k] + (size_t)(cl.x & kCellMask) * nl);
          b[k] = ldg4(cell0 + m[k] + (size_t)((((uint32_t)cl.x >> kCellFlagShift) & kFlagIcrMask) == 2 ? cl.z : cl.y) * nl);
        }
      }
#pragma unroll
      for (int k = 0; k < kPairsInFlight; k++) {
        if (has[k]) {
          const NodeRec &nd = cn[slot[k]];
          pair_store(B, p + NP * k, m[k], nd, interp2(a[k], b[k], node_icr(&nd) == 2 ? nd.wr : nd.wt));
        }
      }
    }
    // pass 2: the extra points of the chunk, compacted
    const int n3 = __popc(mask3) * nlc;
    for (int q = plane; q < n3; q += NP) {
      const int k3 = (int)(((uint32_t)q * Mdiv) >> 20), m3 = q - k3 * nlc;
      const int slot3 = (int)__fns(mask3, 0, k3 + 1);
      const NodeRec &nd = cn[slot3];
      const int4 cl = nd.cells;
      const double4 *base = cell0 + m3;
      const double4 a = ldg4(base + (size_t)(cl.x & kCellMask) * nl), b = ldg4(base + (size_t)cl.y * nl);
      const double4 c4 = ldg4(base + (size_t)cl.z * nl), d = ldg4(base + (size_t)cl.w * nl);
      pair_store(B, slot3 * nlc + m3, m3, nd, interp4(a, b, c4, d, nd.wr, nd.wt));
    }
    __threadfence_block();
    __syncwarp();
    if ((plane & 31) == 0) mbar_arrive((uint32_t)__cvta_generic_to_shared(&s_mbar[b]));  // full[b]
  }
}

// ---- consumer side: the segments of one staged chunk ------------------------------------------------
// The only true recurrence along the ray is I <- I x + q; the profile, exp(-dtau) and the source-function
// weights of a segment depend on staged node data alone.  NS consecutive segments are therefore
// evaluated in ONE basic block (their dependency chains interleave: the warp's instruction-level
// parallelism is what hides the FP64 latency at 4-5 resident warps per scheduler), then chained into
// the intensity with NS fused multiply-adds.  The qdr_src_2 case split is voted once per group.
template <int NS>
__device__ __forceinline__ void jam_steps(uint32_t an, uint32_t ah, uint32_t stride, double dnu, Item &it,
                                          int &r0ok, uint32_t T1) {
  double hds[NS], src1[NS], alp1[NS], alpl1[NS];
  int k1hi[NS];
#pragma unroll
  for (int k = 0; k < NS; k++) {
    hds[k] = lds_f64(an + (uint32_t)k * (uint32_t)sizeof(HotNode));
    const uint32_t a = ah + (uint32_t)k * stride;
    const double2 iv = lds_f64x2(a + 32), ak = lds_f64x2(a + 16), sa = lds_f64x2(a);
    const double u = fma(dnu, iv.x, -iv.y);
    const double e = gauss_tab(u, T1, 0);
    alpl1[k] = ak.y * e;
    src1[k] = fma(ak.x, e, sa.x);
    alp1[k] = sa.y + alpl1[k];
    k1hi[k] = __double2hiint(ak.y);
  }
  double dtau[NS], theo[NS];
  bool work = false;
  int anyk1 = 0;
#pragma unroll
  for (int k = 0; k < NS; k++) {
    const double a0 = k ? alp1[k - 1] : it.alp0, s0 = k ? src1[k - 1] : it.src0;
    dtau[k] = hds[k] * (a0 + alp1[k]);
    theo[k] = hds[k] * (s0 + src1[k]);
    work = work || gt_thin(dtau[k]);
    anyk1 |= k1hi[k];
  }
  work = work || (anyk1 < 0);  // inverted populations force the full path, which carries the maser test
  if (!__any_sync(0xffffffffu, work)) {
    // transfer.F:1522-1524,1545: Q = theomax, xp = 1 - dtau
#pragma unroll
    for (int k = 0; k < NS; k++) it.inten = fma(it.inten, 1.0 - dtau[k], theo[k]);
    r0ok = 0;
  } else {
    double r = r0ok ? it.r0 : div_fast(it.src0, it.alp0);
    double x[NS], q[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
      const double a0 = k ? alp1[k - 1] : it.alp0;
      double rn;
      step_coeffs(a0, r, src1[k], alp1[k], rn, dtau[k], theo[k], T1, x[k], q[k]);
      r = rn;
    }
#pragma unroll
    for (int k = 0; k < NS; k++) it.inten = fma(it.inten, x[k], q[k]);
    if (__any_sync(0xffffffffu, anyk1 < 0)) {  // telescope.F:4295
      bool ms = false;
#pragma unroll
      for (int k = 0; k < NS; k++) ms = ms || (alpl1[k] * (hds[k] + hds[k]) < (double)(-0.01f));
      if (ms) s_flags[threadIdx.x] |= 1u;
    }
    it.r0 = r;
    r0ok = 1;
  }
  it.src0 = src1[NS - 1];
  it.alp0 = alp1[NS - 1];
}

#ifndef RL_JAM
#define RL_JAM 3
#endif
__device__ __forceinline__ void integrate_chunk(const RenderParams &P, const TileBuf B, int nlc, int cnt,
                                                    Item &it, int l0, int &r0ok) {
  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1);
  const uint32_t stride = (uint32_t)nlc * (uint32_t)sizeof(HotLine);
  const double dnu = s_dnu[threadIdx.x];
  const uint32_t ah0 = B.hl + (uint32_t)(s_meta[threadIdx.x].x - l0) * (uint32_t)sizeof(HotLine);
  int slot = 1;
  while (slot <= cnt) {
    const uint32_t an = B.hn + (uint32_t)slot * (uint32_t)sizeof(HotNode);
    const uint32_t ah = ah0 + (uint32_t)slot * stride;
    // flags of the next RL_JAM slots (block-uniform); slots past the chunk count as flagged
    uint32_t fl[RL_JAM];
#pragma unroll
    for (int k = 0; k < RL_JAM; k++)
      fl[k] = (slot + k <= cnt) ? smem_ptr<HotNode>(B.hn)[slot + k].flags : 1u;
    int n = 0;  // number of leading unflagged slots
#pragma unroll
    for (int k = 0; k < RL_JAM; k++) {
      if (fl[k]) break;
      n = k + 1;
    }
    if (n == RL_JAM) {
      jam_steps<RL_JAM>(an, ah, stride, dnu, it, r0ok, T1);
      slot += RL_JAM;
    } else if (n >= 1) {
      jam_steps<1>(an, ah, stride, dnu, it, r0ok, T1);
      slot += 1;
    } else {
      const int2 mt = s_meta[threadIdx.x];
      double st[4] = {it.inten, it.src0, it.alp0, it.r0};
      const unsigned f = slow_step(st, B, nlc, slot, mt.x, mt.x - l0, mt.y & 0x3fffffff, s_dnu[threadIdx.x],
                                   P.lines, P.star_line, P.nfr, P.starfract);
      it.inten = st[0];
      it.src0 = st[1];
      it.alp0 = st[2];
      it.r0 = st[3];
      if (f) s_flags[threadIdx.x] = (s_flags[threadIdx.x] | (f & 1u)) + ((f >> 1) << 8);
      r0ok = 1;
      slot++;
    }
  }
}

// Block = NT consumer threads (one (line, channel) item each, threads 0..NT-1) + one producer warp.
// The producer gathers, interpolates and stages chunk after chunk of the ray's nodes into a ring of
// kTileBufs shared-memory buffers; the consumers never touch global memory inside the ray loop (the
// flagged-node path excepted).  Hand-over by named barriers: full[b] (producer arrives, consumers
// wait) and empty[b] (consumers arrive, producer waits).
#ifndef RL_BLOCKS128
#define RL_BLOCKS128 4  // resident 128-item tiles per SM the register budget is set for
#endif
template <int NT>
__global__ void __launch_bounds__(NT + kProducerThreads, NT == 128 ? RL_BLOCKS128 : 2 * RL_BLOCKS128)
    tile_kernel(const __grid_constant__ RenderParams P) {
  extern __shared__ double4 smem_raw[];
  constexpr int NALL = NT + kProducerThreads;
  const int tid = threadIdx.x, lane = tid & 31;
  for (int j = tid; j < kTabN; j += NALL) s_T1[j] = exp2((double)j * (1.0 / kTabN));
  if (tid < 2 * kTileBufs)
    mbar_init((uint32_t)__cvta_generic_to_shared(&s_mbar[tid]), tid < kTileBufs ? kProducerWarps : NT / 32);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const TileDesc td = P.tiles[blockIdx.x];
  const int ray = td.ray, l0 = td.l0, nlc = td.nlc;
  const unsigned g0 = td.g0, g1 = td.g1;
  // opaque-wall start (wall_kernel): the walk begins one segment BEFORE the first segment that matters;
  // that segment, the "first segment of the ray" here, sets up the carried profile state at its end node
  // exactly (line.F:4613-4615), and whatever it contributes itself lies behind the whole wall
  const int Nfull = (int)(P.node_off[ray + 1] - P.node_off[ray]);
  const int shift = P.nstart ? max(0, min(__ldg(&P.nstart[ray]), Nfull - 1) - 2) : 0;
  const long long n0 = P.node_off[ray] + shift;
  const int N = Nfull - shift;
  // shared-memory carve-up: kTileBufs buffers of nch+1 slots (+1 phantom slot of hn and hl, see
  // integrate_chunk)
  const int slot_bytes = nlc * kPairBytes + kSlotBytes, ph_bytes = nlc * (int)sizeof(HotLine) + (int)sizeof(HotNode);
  int nch = (P.smem_budget / kTileBufs - ph_bytes) / slot_bytes - 1;
  nch = max(1, min(kTileChunk, nch));
  const int bufbytes = (nch + 1) * slot_bytes + ph_bytes;
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
  auto make_buf = [&](int b) {
    TileBuf t;
    t.hn = smem0 + (uint32_t)(b * bufbytes);
    t.hl = t.hn + (uint32_t)((nch + 2) * sizeof(HotNode));
    t.cn = t.hl + (uint32_t)((nch + 2) * nlc * sizeof(HotLine));
    t.cl = t.cn + (uint32_t)((nch + 1) * sizeof(ColdNode));
    return t;
  };
  const int nchunks = (N > 1) ? (N - 1 + nch - 1) / nch : 0;

  if (tid >= NT) {
    // ---------------- producer warp ----------------
    __syncthreads();  // (exp tables; keeps the barrier-0 count uniform)
    producer_loop(P, smem0, bufbytes, nch, nlc, l0, n0, N, nchunks, tid - NT, NALL);
    return;
  }

  // ---------------- consumers ----------------
  // my item: g0 + tid; surplus threads shadow the tile's first item and store nothing
  Item it;
  bool mine;
  {
    const unsigned *off = P.item_off + (size_t)ray * P.nl;
    const unsigned base = off[0];
    unsigned my = g0 + tid;
    mine = my < g1;
    if (my >= g1) my = g0;
    int a = l0, b = l0 + nlc;
    while (b - a > 1) {
      const int mid = (a + b) >> 1;
      if (off[mid] - base <= my) a = mid;
      else b = mid;
    }
    const int4 rg = P.rng[(long long)ray * P.nl + a];
    bool msk = false;
    const int ch = task_chan(P, rg, (int)(my - (off[a] - base)), msk);
    s_meta[tid] = make_int2(a, ch | (msk ? 0x40000000 : 0));
    s_dnu[tid] = P.line_dnu[(size_t)a * P.nfr + ch];
    s_flags[tid] = 0;
    it.inten = (P.out_itype == 3) ? P.isrf_line[(size_t)a * P.nfr + ch] : P.lines[a].i_outer;
    it.src0 = it.alp0 = it.r0 = 0.0;
  }
  int r0ok = 0;
  __syncthreads();  // exp tables
  for (int c = 0, b = 0; c < nchunks; c++, b = (b + 1 == kTileBufs) ? 0 : b + 1) {
    mbar_wait((uint32_t)__cvta_generic_to_shared(&s_mbar[b]), (uint32_t)((c / kTileBufs) & 1));  // full[b]
    const int c0 = 1 + c * nch;
    integrate_chunk(P, make_buf(b), nlc, min(nch, N - c0), it, l0, r0ok);
    __syncwarp();
    if (lane == 0 && c + kTileBufs < nchunks)
      mbar_arrive((uint32_t)__cvta_generic_to_shared(&s_mbar[kTileBufs + b]));  // empty[b]
  }
  unsigned long long r = 0, x = 0;
  if (mine) {
    const int2 mt = s_meta[tid];
    const int ch = mt.y & 0x3fffffff;
    const size_t row = (size_t)mt.x * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
    P.img[row * P.nfr + ch] = it.inten;
    if (P.sparse && it.inten == 0.0) P.dense[(long long)ray * P.nl + mt.x] = 2;  // see fill_sparse_kernel
    if (P.integ) P.integ[row * P.nfr + ch] = (mt.y & 0x40000000) ? 1 : 2;
    const unsigned f = s_flags[tid];
    if (f & 1u) atomicOr(&P.maser[mt.x], 1);
    r = 1;
    x = f >> 8;
  }
  // work counters: every item walks the ray's N-1 segments; sub-gridding adds extra elements
  // (the reference's counts, whether or not an opaque wall shortened the walk; ex = executed elements)
  unsigned long long s = r * (unsigned long long)(Nfull > 0 ? Nfull - 1 : 0), e = s + x;
  unsigned long long ex = r * (unsigned long long)(N > 0 ? N - 1 : 0) + x;
  for (int o = 16; o; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ex += __shfl_xor_sync(0xffffffffu, ex, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if (lane == 0 && r) {
    atomicAdd(&P.counters[0], r);
    atomicAdd(&P.counters[1], e);
    atomicAdd(&P.counters[2], s);
    atomicAdd(&P.counters[3], ex);
  }
}

// ------------------------------------------------------------------------------------------
// ztile_kernel: the formal solution with the LINES of a batch across the lanes of a warp.
//
// The argument of the line profile, (dnu - nu0 Omega.v/c) / (3.33567e-6 nu0 lwav) (line.F:4559-4566,
// 2301), does not depend on the line: every line of a render shares the velocity grid of the passband
// (line.F:462-469: dnu_k = nu0 * 3.33567e-6 * v_k) and the local width, so nu0 cancels.  The Gaussian
// is therefore evaluated once per (ray, node, channel) -- cooperatively by the lanes of the warp, into a
// small shared-memory table -- and not once per (ray, node, channel, line) as the reference does.
//
// One warp = one ray x up to 32 lines (one per lane; with fewer lines per tile the lane pattern repeats
// and the repeats take different channels) x up to kZCw channels per thread, intensities in registers.
// Per node every thread gathers and interpolates the cells of ITS line once (coalesced: the lines of a
// cell are contiguous in cellL; the loads of the next node are in flight while this one is integrated),
// folds everything that does not depend on the channel into six constants
//     dtau = D + P e0 + Q e1 ,  theomax = Th + R e0 + S e1       (e0, e1: profile at the two nodes)
// and then walks its channels: independent dependency chains, no per-channel state but the intensity.
// The case split of transfer.F:1517,1542 is voted: first for the whole node (D + |P| + |Q| <= 1e-9 on
// every lane: all channels take I <- I (1 - dtau) + theomax without further tests), else per group of
// three channels: all lanes thin -> the same update; all lanes dtau > 1e-6 with positive opacities ->
// qdr_src_2 without its case selections; else the branch-free general step.  Nodes flagged by the
// geometry (first segment, inner hole / star mixing, 6q > 1 sub-grid candidates) go through zflagged,
// out of line.  What every lane would otherwise recompute per node (cell offsets, flags, interpolation
// weight, ds/2, the profile scale) is derived once by the lane that stages the node record.
// The walk starts at the segment wall_kernel reports (opaque-wall start, below).
// There is no block-level synchronisation: a block is one warp (kZWarps = 1), 168 registers.
// ------------------------------------------------------------------------------------------
constexpr double kIanScale = kTabSqrtScale / 3.33567e-6;  // scaled reciprocal Doppler width per 1/lwav

struct ZVal {  // channel-independent values of one line at one node
  double sd, ad;  // dust source / opacity (line.F:4058-4063)
  double cN, kk;  // c_src N_up ; c_alp (N_down B_du - N_up B_ud)
};
struct __align__(16) ZNx {  // per staged node, derived once by the staging lane
  uint32_t offA, offB;  // element offsets (cell * nl) of the first stencil pair in cellL
  uint32_t fl, icr;     // segment flags (sub-grid switch and first-segment flag applied) ; crossing type
  double w, hds;        // interpolation weight of the pair ; ds / 2
  double ian, dvi;      // scaled reciprocal Doppler width of the segment ending here ; Omega.v/c times it
};
struct ZSeg {  // a flagged segment of one line, handed to zflagged through local memory
  double ds, lwav, dv0, dv1, ian;
  double nrm0, nrm1;  // profile norm 0.5642/aa of the previous segment (carried state) and of this one
  ZVal v0, v1;
};

__device__ __forceinline__ ZVal zvals(const double4 v, double c_src, double cb_du, double cb_ud) {
  ZVal o;
  o.sd = v.x;
  o.ad = v.y;
  o.cN = __dmul_rn(c_src, v.z);
  o.kk = fma(v.w, cb_du, -__dmul_rn(v.z, cb_ud));
  return o;
}

// flagged segment of one line for the cw channels of a thread (Ic in/out).  ep_a / ec_a: shared-window
// addresses of the thread's profile values at the two nodes.  Channel of slot c: list position
// min(jb + c gw, jmax).  Returns the maser bits (telescope.F:4295); xtra += the extra element
// integrations of sub-gridded segments (real items only).
__device__ __noinline__ unsigned zflagged(double *Ic, const ZSeg &g, uint32_t fl, int cw, uint32_t ep_a,
                                          uint32_t ec_a, const LineDev *__restrict__ L,
                                          const double *__restrict__ dnu_l, const double *__restrict__ star_l,
                                          const double *__restrict__ velo, int jb, int gw, int jmax, int cmin,
                                          double starfract, unsigned realbits, unsigned &xtra) {
  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1);
  const double nu0 = __ldg(&L->nu0), k_aa = __ldg(&L->k_aa), inv_nu0 = __ldg(&L->inv_nu0);
  const double ds = g.ds;
  unsigned mb = 0;
  for (int c = 0; c < cw; c++) {
    const int j = min(jb + c * gw, jmax);
    const int ch = j ? cmin + j - 1 : 0;
    const double dnu = __ldg(dnu_l + ch);
    const double ep = lds_f64(ep_a + 8u * (uint32_t)c), ec = lds_f64(ec_a + 8u * (uint32_t)c);
    double inten = Ic[c];
    int init = 0;
    if (fl & (kFlagInit | kFlagStar | kFlagZero)) {
      if (fl & kFlagZero) inten = 0.0;
      if (fl & kFlagStar) inten = (1.0 - starfract) * inten + starfract * __ldg(star_l + ch);
      init = 1;
    }
    double srcl0 = g.v0.cN * (g.nrm0 * ep), alpl0 = g.v0.kk * (g.nrm0 * ep);  // carried state (line.F:4613-4615)
    bool done = false, ms = false;
    if (fl & kFlagSub) {  // line.F:4706-4745
      const double q = fabs((g.dv1 - g.dv0) / (g.lwav / 2.99792458e5));
      const double s_c = ds * (dnu * inv_nu0 - g.dv0) / (g.dv1 - g.dv0);
      const double dls3 = 3.0 * (ds / q);
      const double sright = s_c + dls3, sleft = s_c - dls3;
      if (sright > 0.0 && sleft < ds) {
        const int n = subgrid_tile(nu0, k_aa, dnu, inten, ds, sleft, sright, g.v0.sd, g.v0.ad, g.v0.cN, g.v0.kk,
                                   g.dv0, g.v1.sd, g.v1.ad, g.v1.cN, g.v1.kk, g.dv1, g.lwav, srcl0, alpl0, init);
        ms = alpl0 * ds < (double)(-0.01f);
        if ((realbits >> c) & 1u) xtra += (unsigned)(n - 1);
        done = true;
      }
    }
    if (!done) {
      if (init) {  // line.F:4559-4586: the start point with this segment's width
        const double vel = __ldg(velo + ch);
        const double e0 = gauss_tab(fma(vel, g.ian, -(g.dv0 * g.ian)), T1, 0);
        srcl0 = g.v0.cN * (g.nrm1 * e0);
        alpl0 = g.v0.kk * (g.nrm1 * e0);
      }
      const double src0 = g.v0.sd + srcl0, alp0 = g.v0.ad + alpl0;
      const double alpl1 = g.v1.kk * (g.nrm1 * ec);
      const double src1 = fma(g.v1.cN, g.nrm1 * ec, g.v1.sd), alp1 = g.v1.ad + alpl1;
      const double hds = 0.5 * ds;
      const double dtau = hds * (alp0 + alp1), theomax = hds * (src0 + src1);
      double x, q;
      step_onediv(alp0, src0, alp1, src1, dtau, theomax, T1, x, q);
      inten = fma(inten, x, q);
      ms = alpl1 * ds < (double)(-0.01f);
    }
    if (ms) mb |= 1u << c;
    Ic[c] = inten;
  }
  return mb;
}

// Three channels of one unflagged segment through the fully general step (an opacity that is not positive
// somewhere in the warp: dust-free cells, inverted populations), out of line: rare, and ztile_kernel's node
// loop has to fit the instruction cache.  g: the segment's channel-independent values (ds, nrm0, nrm1, v0, v1);
// returns the maser bits of the three channels (telescope.F:4295).
__device__ __noinline__ unsigned zgeneral3(double *I3, const double *ep, const double *ec, const double *dtau,
                                           const double *theo, const ZSeg &g) {
  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1);
  const double K0 = g.v0.kk * g.nrm0, A0 = g.v0.cN * g.nrm0, K1 = g.v1.kk * g.nrm1, A1 = g.v1.cN * g.nrm1;
  const bool neg = g.v1.kk < 0.0;
  unsigned mb = 0;
#pragma unroll 1
  for (int k = 0; k < 3; k++) {
    const double alp1 = fma(K1, ec[k], g.v1.ad), src1 = fma(A1, ec[k], g.v1.sd);
    const double alp0 = fma(K0, ep[k], g.v0.ad), src0 = fma(A0, ep[k], g.v0.sd);
    double x, q;
    step_onediv(alp0, src0, alp1, src1, dtau[k], theo[k], T1, x, q);
    I3[k] = fma(I3[k], x, q);
    if (neg && (K1 * ec[k]) * g.ds < (double)(-0.01f)) mb |= 1u << k;
  }
  return mb;
}

#ifndef RL_ZPIPE
#define RL_ZPIPE 1
#endif
#ifndef RL_ZPD
#define RL_ZPD 1  // how many nodes ahead the stencil cells are requested
#endif
#if RL_ZPIPE && RL_ZPD != 1
#error "RL_ZPIPE needs RL_ZPD == 1"
#endif
#ifndef RL_ZFAST
#define RL_ZFAST 1
#endif
#ifndef RL_ZSTREAM
#define RL_ZSTREAM 1
#endif
#ifndef RL_ZONEDIV
#define RL_ZONEDIV 1
#endif
#ifndef RL_ZMINB
#define RL_ZMINB 6  // resident blocks per SM the register budget is set for (6: 168 registers, no hot-loop spills)
#endif
template <int CW>
#ifdef RL_ZMAXREG
__global__ void __maxnreg__(RL_ZMAXREG) ztile_kernel
#else
__global__ void __launch_bounds__(32 * kZWarps, RL_ZMINB * 2 / kZWarps) ztile_kernel
#endif
(const __grid_constant__ RenderParams P) {
  // per warp: profile table, staged node records, derived records -- one struct, so that the hot loop addresses
  // all three as (one 32-bit shared-window base) + constant
  struct __align__(16) ZShm {
    double et[kZTab];
    NodeRec nd[32];
    ZNx nx[32];
  };
  __shared__ ZShm s_z[kZWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kTabN) s_T1[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / kTabN));
  __syncthreads();
  const unsigned tix = blockIdx.x * kZWarps + warp;
  if (tix >= P.nztile) return;
  const ZTile t = P.ztiles[tix];
  const int ray = t.ray, lws = t.lwshift, gws = 5 - lws, GW = 1 << gws;
  const int ls = lane & ((1 << lws) - 1), g = lane >> lws;
  const bool lineok = ls < (int)t.nlt;
  const int l = (int)P.zlines[t.loff + (lineok ? ls : 0)];
  const int nchk = t.nchk, j0 = t.j0, cmin = t.cmin;
  const int cw = (nchk + GW - 1) >> gws;  // channels per thread
  const int cw3 = 3 * ((cw + 2) / 3);
  const int jmax = j0 + nchk - 1, jb = j0 + g;
  // channel of my slot c: position min(jb + c GW, jmax) of the list {0, cmin, cmin + 1, ...}
  auto chan_of = [&](int c) {
    const int j = min(jb + c * GW, jmax);
    return j ? cmin + j - 1 : 0;
  };
  const int4 rg = P.rng[(long long)ray * P.nl + l];
  unsigned realbits = 0;  // slots that are channels the reference integrates for this line
#pragma unroll
  for (int c = 0; c < CW; c++) {
    if (lineok && c * GW + g < nchk) {
      const int ch = chan_of(c);
      if (ch == 0 || (ch >= rg.x && ch <= rg.y) || ch == rg.z) realbits |= 1u << c;
    }
  }
  const LineDev *Lp = P.lines + l;
  const double c_src = __ldg(&Lp->c_src);
  const double cb_du = __ldg(&Lp->c_alp) * __ldg(&Lp->bdu), cb_ud = __ldg(&Lp->c_alp) * __ldg(&Lp->bud);
  const double knorm = 0.56419583546 / __ldg(&Lp->k_aa);
  double I[CW];
#pragma unroll
  for (int c = 0; c < CW; c++)
    I[c] = (P.out_itype == 3) ? __ldg(&P.isrf_line[(size_t)l * P.nfr + chan_of(c)]) : __ldg(&Lp->i_outer);

  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1);
  const uint32_t et0 = (uint32_t)__cvta_generic_to_shared(&s_z[warp].et[0]);
  constexpr uint32_t kNdOff = (uint32_t)(kZTab * sizeof(double)), kNxOff = kNdOff + 32u * (uint32_t)sizeof(NodeRec);
  static_assert(sizeof(ZNx) == 48 && sizeof(NodeRec) == 64, "record sizes the hot loop's offsets assume");
  NodeRec *snd = s_z[warp].nd;
  ZNx *snx = s_z[warp].nx;
  const long long n0 = P.node_off[ray];
  const int N = (int)(P.node_off[ray + 1] - n0);
  const NodeRec *__restrict__ rec = P.nodes.rec + n0;
  const double4 *__restrict__ cl = P.cellL + l;
  const size_t nl = (size_t)P.nl;
  // first segment to integrate: everything before it lies behind an opaque dust wall (wall_kernel)
  const int ns = P.nstart ? max(1, min(__ldg(&P.nstart[ray]), N - 1)) : 1;
  if (ns > 1) {
#pragma unroll
    for (int c = 0; c < CW; c++) I[c] = 0.0;
  }
  const int cwS = cw3 | 1;  // table columns per channel group (odd: the groups' reads never conflict)
  const int NB = min(31 - RL_ZPD, kZTab / (GW * cwS) - 1);  // nodes per batch
  const uint32_t rowB = (uint32_t)(GW * cwS) * 8u;  // bytes of one node's row of the profile table
  const int W = GW * cw3;                                // columns of the profile table
  const int fsi = 32 / W, fqi = 32 - fsi * W;            // the fill loop's step of 32 entries in (row, column)
  const int fs0 = lane / W, fq0 = lane - fs0 * W;        // this lane's first entry
  unsigned mbits = 0, xtra = 0;
#ifdef RL_STATS
  unsigned long long st[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  auto st_far = [&](int base, uint32_t ep_a, uint32_t ec_a) {
    for (int c = 0; c < cw; c++) {
      const double em = fmax(lds_f64(ep_a + 8u * (uint32_t)c), lds_f64(ec_a + 8u * (uint32_t)c));
      st[base]++;
      if (base && em == 0.0) st[base + 1]++;
      if (base && em <= 3.72e-44) st[base + 2]++;
      if (base && em <= 1.6e-28) st[base + 3]++;
    }
  };
#endif

  // the two (four for extra points) stencil cells of a node; the first pair is pulled into L1 one node
  // ahead (no registers held across the channel loop)
  // RL_ZPIPE = 1: the loads of the next node's first cell pair are issued before the channel loop of this
  // node and consumed after it (eight registers per cell in flight); RL_ZPIPE = 0: they are only pulled
  // into L1 (RL_ZPD nodes ahead)
  double4 pa, pb;
  auto derive = [&](const int4 cells, double ds, double wr, double wt, double inv_lwav, double dvmu, int node) {
    ZNx x;
    x.icr = ((uint32_t)cells.x >> kCellFlagShift) & kFlagIcrMask;
    x.offA = (uint32_t)(cells.x & kCellMask) * (uint32_t)nl;
    x.offB = (uint32_t)(x.icr == 2 ? cells.z : cells.y) * (uint32_t)nl;
    uint32_t fl = ((uint32_t)cells.x >> kCellFlagShift) & ~kFlagIcrMask;
    if (!P.subgrid) fl &= ~kFlagSub;
    if (node == 1) fl |= kFlagInit;  // first segment of the ray: nothing carried yet
    x.fl = fl;
    x.w = x.icr == 2 ? wr : wt;
    x.hds = 0.5 * ds;
    x.ian = inv_lwav * kIanScale;
    x.dvi = dvmu * x.ian;
    return x;
  };
  auto prefetch = [&](uint32_t offA, uint32_t offB) {
#if RL_ZPIPE
    pa = ldg4(cl + offA);
    pb = ldg4(cl + offB);
#else
    asm volatile("prefetch.global.L1 [%0];" ::"l"(cl + offA));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(cl + offB));
#endif
  };
  // interpolated cell values at a node (x: its derived record, raw: its node record, read only for the
  // four-cell extra points)
  auto gather = [&](const ZNx x, const NodeRec *raw) {
#if !RL_ZPIPE
    pa = ldg4(cl + x.offA);
    pb = ldg4(cl + x.offB);
#endif
    if (x.icr == 3) {
      const int4 cells = raw->cells;
      const double4 c4 = ldg4(cl + (size_t)cells.z * nl), d4 = ldg4(cl + (size_t)cells.w * nl);
      return interp4(pa, pb, c4, d4, raw->wr, raw->wt);
    }
    return interp2(pa, pb, x.w);
  };
  ZVal v0;
  v0.sd = v0.ad = v0.cN = v0.kk = 0.0;
  double nrm0 = 0.0;  // profile norm of the segment that ended at the previous node
  if (N > 0) {  // values at the node the first integrated segment starts from
    const NodeRec *r0 = rec + (ns - 1);
    const ZNx x0 = derive(__ldg(&r0->cells), 0.0, __ldg(&r0->wr), __ldg(&r0->wt), 1.0, 0.0, 0);
#if RL_ZPIPE
    prefetch(x0.offA, x0.offB);
#endif
    v0 = zvals(gather(x0, r0), c_src, cb_du, cb_ud);
    nrm0 = knorm * __ldg(&r0->inv_lwav);
  }
  for (int c0 = ns; c0 < N; c0 += NB) {
    const int cnt = min(NB, N - c0);
    __syncwarp();  // the previous batch is consumed
    {  // nodes c0-1 .. c0+cnt-1+RL_ZPD (the last ones only as look-ahead for the gathers) -> shared memory
      const int nstage = min(cnt + 1 + RL_ZPD, N - (c0 - 1));
      if (lane < nstage) {
        const int4 *src = reinterpret_cast<const int4 *>(rec + (c0 - 1 + lane));
        int4 *dst = reinterpret_cast<int4 *>(snd + lane);
#pragma unroll
        for (int k = 0; k < 4; k++) dst[k] = __ldg(src + k);
        const NodeRec *mine = snd + lane;  // (own writes: visible to this lane)
        snx[lane] = derive(mine->cells, mine->ds, mine->wr, mine->wt, mine->inv_lwav, mine->dvmu, c0 - 1 + lane);
      }
    }
    __syncwarp();
#if RL_ZPIPE
    if (c0 == ns) prefetch(snx[1].offA, snx[1].offB);  // later batches: in flight since the previous batch's last node
#else
    if (c0 == ns) {
#pragma unroll
      for (int k = 1; k <= RL_ZPD; k++)
        if (c0 - 1 + k < N) prefetch(snx[k].offA, snx[k].offB);
    }
#endif
    {  // profile table of the batch: rows = nodes c0-1 .. c0+cnt-1, columns = (channel group, slot).  The lanes
       // walk (row s, column q) with q = slot * GW + group -- the position in the tile's channel list is then
       // j0 + q, group and slot fall out of q by mask and shift -- advancing (s, q) by 32 entries per iteration
       // without a division
      int q = fq0, sr = fs0;
      const int rows = cnt + 1;
      while (sr < rows) {
        const int gg = q & (GW - 1), c = q >> gws;
        const int j = min(j0 + q, jmax);
        const double vel = __ldg(P.velz + (j ? cmin + j - 1 : 0));
        const double2 sc = *reinterpret_cast<const double2 *>(&snx[sr].ian);  // {ian, dvi}
        const double e = gauss_tab(fma(vel, sc.x, -sc.y), T1, 0);
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(et0 + (uint32_t)(((sr << gws) + gg) * cwS + c) * 8u), "d"(e)
                     : "memory");
        q += fqi;
        sr += fsi;
        if (q >= W) {
          q -= W;
          sr++;
        }
      }
    }
    __syncwarp();
    for (int s = 1; s <= cnt; s++) {
      const NodeRec *nd = snd + s;  // (generic pointer: the rare paths only)
      const uint32_t nxa = et0 + kNxOff + 48u * (uint32_t)s, nda = et0 + kNdOff + 64u * (uint32_t)s;
      ZNx nx;
      {
        const uint2 fi = lds_u32x2(nxa + 8u);
        const double2 wh = lds_f64x2(nxa + 16u);
        nx.fl = fi.x;
        nx.icr = fi.y;
        nx.w = wh.x;
        nx.hds = wh.y;
#if !RL_ZPIPE
        const uint2 ob = lds_u32x2(nxa);
        nx.offA = ob.x;
        nx.offB = ob.y;
#endif
      }
      const uint32_t fl = nx.fl;
      const ZVal v1 = zvals(gather(nx, nd), c_src, cb_du, cb_ud);
      if (c0 - 1 + s + RL_ZPD < N) {
        const uint2 ob = lds_u32x2(nxa + 48u * RL_ZPD);
        prefetch(ob.x, ob.y);
      }
      const double nrm1 = knorm * lds_f64(nda + 24u);  // NodeRec::inv_lwav
      const uint32_t ep_a = et0 + (uint32_t)(((s - 1) << gws) + g) * (uint32_t)(cwS * 8);
      const uint32_t ec_a = ep_a + rowB;
      if (fl == 0) {
        const double hds = nx.hds;
        const double hn0 = hds * nrm0, hn1 = hds * nrm1;
        const double D = hds * (v0.ad + v1.ad), Th = hds * (v0.sd + v1.sd);
        const double Pq = hn0 * v0.kk, Q = hn1 * v1.kk, R = hn0 * v0.cN, S = hn1 * v1.cN;
        const bool neg = v1.kk < 0.0;  // inverted populations: full path, which carries the maser test
        // lanes whose dust opacity alone is positive never see alpha <= 0 unless the line inverts
        const bool nonpos = neg | !(v0.ad > kAlpMin) | !(v1.ad > kAlpMin) | (v0.kk < 0.0);
#if RL_ZFAST
        // the profile is <= 1: if D + |P| + |Q| <= 1e-9 on every lane, every channel of the node takes the
        // thin branch of transfer.F:1522-1524,1545 (Q = theomax, xp = 1 - dtau): no votes, no branches, all
        // channels' chains independent
        const bool maybe = neg | gt_thin(D + fabs(Pq) + fabs(Q));
        if (cw3 == CW && !__any_sync(0xffffffffu, maybe)) {
#ifdef RL_STATS
          st_far(0, ep_a, ec_a);
#endif
#pragma unroll
          for (int c = 0; c < CW; c++) {
            const double ep = lds_f64(ep_a + 8u * (uint32_t)c), ec = lds_f64(ec_a + 8u * (uint32_t)c);
            const double dtau = fma(Pq, ep, fma(Q, ec, D)), theo = fma(R, ep, fma(S, ec, Th));
            I[c] = fma(I[c], 1.0 - dtau, theo);
          }
        } else
#endif
        {
#ifdef RL_STATS
          st_far(4, ep_a, ec_a);
#endif
#pragma unroll
          for (int cb = 0; cb < CW; cb += 3) {
            if (cb < cw) {
              double ep[3], ec[3], dtau[3], theo[3];
#pragma unroll
              for (int k = 0; k < 3; k++) {
                ep[k] = lds_f64(ep_a + 8u * (uint32_t)(cb + k));
                ec[k] = lds_f64(ec_a + 8u * (uint32_t)(cb + k));
                dtau[k] = fma(Pq, ep[k], fma(Q, ec[k], D));
                theo[k] = fma(R, ep[k], fma(S, ec[k], Th));
              }
              // gt_thin / gt_mid of the three channels at once: largest and smallest high word
              const int h0 = __double2hiint(dtau[0]), h1 = __double2hiint(dtau[1]), h2 = __double2hiint(dtau[2]);
              const bool work = neg | (max(max(h0, h1), h2) > kThinHi);
              const bool thin_some = !(min(min(h0, h1), h2) > kMidHi);
              if (!__any_sync(0xffffffffu, work)) {
#ifdef RL_STATS
                st[8]++;
#endif
                // transfer.F:1522-1524,1545: Q = theomax, xp = 1 - dtau
#pragma unroll
                for (int k = 0; k < 3; k++) I[cb + k] = fma(I[cb + k], 1.0 - dtau[k], theo[k]);
              } else {
                const double K0 = v0.kk * nrm0, A0 = v0.cN * nrm0, K1 = v1.kk * nrm1, A1 = v1.cN * nrm1;
#if RL_ZSTREAM
                const bool anyodd = __any_sync(0xffffffffu, nonpos | thin_some);
#ifdef RL_STATS
                st[anyodd ? 10 : 9]++;
#endif
                if (!anyodd) {
                  // every lane: dtau > 1e-6 and both opacities positive -- qdr_src_2 without its case
                  // selections (the operations of step_coeffs on this branch, bit for bit)
#pragma unroll
                  for (int k = 0; k < 3; k++) {
                    const double alp1 = fma(K1, ec[k], v1.ad), src1 = fma(A1, ec[k], v1.sd);
                    const double alp0 = fma(K0, ep[k], v0.ad), src0 = fma(A0, ep[k], v0.sd);
                    const double xpe = expneg_tab(dtau[k], T1, 0);
#if RL_ZONEDIV
                    // a S0 + b S1 with a = e0 - b, b = e1 / dtau, S = j / alpha (transfer.F:1519-1541) over the
                    // common denominator dtau alp0 alp1: one reciprocal instead of three
                    const double e0 = 1.0 - xpe, e1 = dtau[k] - e0;
                    const double den = dtau[k] * (alp0 * alp1);
                    const double num = fma(fma(e0, dtau[k], -e1), src0 * alp1, e1 * (src1 * alp0));
                    const double qv = min_sel(div_fast(num, den), theo[k]);
#else
                    const double r0 = div_fast(src0, alp0), r1 = div_fast(src1, alp1);
                    const double e0 = 1.0 - xpe;
                    const double bt = div_fast(dtau[k] - e0, dtau[k]);
                    const double qv = min_sel(fma(e0 - bt, r0, bt * r1), theo[k]);
#endif
                    I[cb + k] = fma(I[cb + k], xpe, qv);
                  }
                } else if (!__any_sync(0xffffffffu, nonpos)) {
                  // both opacities positive on every lane, thin (dtau <= 1e-6) and thick channels in the group:
                  // step_onediv with its opacity selections resolved (the same operations, bit for bit)
#pragma unroll
                  for (int k = 0; k < 3; k++) {
                    const double alp1 = fma(K1, ec[k], v1.ad), src1 = fma(A1, ec[k], v1.sd);
                    const double alp0 = fma(K0, ep[k], v0.ad), src0 = fma(A0, ep[k], v0.sd);
                    const double xpe = expneg_tab(dtau[k], T1, 0);
                    const double e0 = 1.0 - xpe, e1 = dtau[k] - e0;
                    const bool thick = gt_mid(dtau[k]);
                    const double hb = 0.5 * dtau[k];
                    const double ca = thick ? fma(e0, dtau[k], -e1) : hb, cbb = thick ? e1 : hb;
                    const double dd = thick ? dtau[k] : 1.0;
                    const double den = dd * (alp0 * alp1);
                    const double num = fma(ca, src0 * alp1, cbb * (src1 * alp0));
                    const double x = thick ? xpe : (1.0 - dtau[k]);
                    double qv = div_fast(num, den);
                    qv = gt_thin(dtau[k]) ? min_sel(qv, theo[k]) : theo[k];
                    I[cb + k] = fma(I[cb + k], x, qv);
                  }
                } else
#endif
                {
                  ZSeg sg;
                  sg.ds = nd->ds;
                  sg.nrm0 = nrm0;
                  sg.nrm1 = nrm1;
                  sg.v0 = v0;
                  sg.v1 = v1;
                  // (copies: the arrays of the hot paths must not have their address taken)
                  double I3[3] = {I[cb], I[cb + 1], I[cb + 2]};
                  const double ep3[3] = {ep[0], ep[1], ep[2]}, ec3[3] = {ec[0], ec[1], ec[2]};
                  const double dt3[3] = {dtau[0], dtau[1], dtau[2]}, th3[3] = {theo[0], theo[1], theo[2]};
                  mbits |= zgeneral3(I3, ep3, ec3, dt3, th3, sg) << cb;
                  I[cb] = I3[0];
                  I[cb + 1] = I3[1];
                  I[cb + 2] = I3[2];
                }
              }
            }
          }
        }
      } else {
#ifdef RL_STATS
        st[11]++;
#endif
        double tmp[CW];
#pragma unroll
        for (int c = 0; c < CW; c++) tmp[c] = I[c];
        ZSeg sg;
        sg.ds = nd->ds;
        sg.lwav = 0.5 * (snd[s - 1].lw + nd->lw);
        sg.dv0 = snd[s - 1].dvmu;
        sg.dv1 = nd->dvmu;
        sg.ian = snx[s].ian;
        sg.nrm0 = nrm0;
        sg.nrm1 = nrm1;
        sg.v0 = v0;
        sg.v1 = v1;
        mbits |= zflagged(tmp, sg, fl, cw, ep_a, ec_a, Lp, P.line_dnu + (size_t)l * P.nfr,
                          P.star_line + (size_t)l * P.nfr, P.velz, jb, GW, jmax, cmin, P.starfract, realbits, xtra);
#pragma unroll
        for (int c = 0; c < CW; c++) I[c] = tmp[c];
      }
      v0 = v1;
      nrm0 = nrm1;
    }
  }
  // results: only the slots that are channels the reference integrates for this line
  unsigned long long r = 0;
  const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
#pragma unroll
  for (int c = 0; c < CW; c++) {
    if ((realbits >> c) & 1u) {
      const int ch = chan_of(c);
      P.img[row * P.nfr + ch] = I[c];
      if (P.sparse && I[c] == 0.0) P.dense[(long long)ray * P.nl + l] = 2;  // see fill_sparse_kernel
      if (P.integ) {
        const bool masked = P.nonredundant ? (ch != rg.z) : (ch == 0);  // telescope.F:548,575
        P.integ[row * P.nfr + ch] = masked ? 1 : 2;
      }
      r++;
    }
  }
#ifdef RL_STATS
  if ((lane & ((1 << lws) - 1)) == 0)  // one lane per channel group
    for (int k = 0; k < 12; k++) atomicAdd(&P.counters[4 + k], st[k]);
#endif
  if (mbits & realbits) atomicOr(&P.maser[l], 1);
  // work counters: every item walks the ray's N-1 segments (the reference's count, whether or not an opaque
  // wall shortened the walk here); sub-gridding adds extra elements; executed = what this kernel integrated
  unsigned long long sct = r * (unsigned long long)(N > 0 ? N - 1 : 0), e = sct + xtra;
  unsigned long long ex = r * (unsigned long long)(N > ns ? N - ns : 0) + xtra;
  for (int o = 16; o; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    sct += __shfl_xor_sync(0xffffffffu, sct, o);
    ex += __shfl_xor_sync(0xffffffffu, ex, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if (lane == 0 && r) {
    atomicAdd(&P.counters[0], r);
    atomicAdd(&P.counters[1], e);
    atomicAdd(&P.counters[2], sct);
    atomicAdd(&P.counters[3], ex);
  }
}

// ------------------------------------------------------------------------------------------
// zcont_kernel: the (ray, line) pairs that carry no channel window on a ray -- the reference integrates
// only their first channel (telescope.F:544-554), whose intensity then stands for all others.  They are a
// quarter of all ray x node steps of a multi-line render but 4 % of its element integrations, and running
// them through ztile_kernel (nine channel slots per thread, 168 registers) wastes both.  Here one warp
// = one ray x 32 such lines (one per lane) x channel 0, four independent warps per block, few registers:
// many resident warps hide the gather latency, and two unflagged segments are evaluated per iteration with
// their loads and their dependency chains interleaved.
// The profile at channel 0 does not depend on the line (see ztile_kernel): one lane per node evaluates it
// while staging the node record.  Where it is exactly 0 at a node (the rule far from the line centre: it is
// flushed below exp(-345)) the line terms of that node vanish identically -- x + 0 y = x -- so only the
// dust pair {j_d, alpha_d} of the cell records is gathered (needfull = 0).
// Arithmetic per lane is that of ztile_kernel's general step (same functions), so a (ray, line) pair gives
// the same bits whichever kernel integrates it.
// ------------------------------------------------------------------------------------------
struct __align__(16) ZCNode {
  uint32_t offA, offB, offC, offD;  // element offsets (cell * nl) of the stencil cells in cellL (C, D: extra points)
  double w, w2;                     // weight between A and B (and C, D) ; between the pairs (extra points)
  double hds, inv_lwav;             // ds / 2 ; reciprocal mean width of the segment ending here
  double lw, dvmu;                  // node values the flagged path needs (line.F:4706-4715)
  double ian, e;                    // profile scale of that segment ; profile value at channel 0
  uint32_t fl, icr, needfull, pad;  // segment flags ; crossing type ; line terms needed at this node
};
constexpr int kZcWarps = 4;
#ifndef RL_ZCMINB
#define RL_ZCMINB 4
#endif

__global__ void __launch_bounds__(32 * kZcWarps, RL_ZCMINB) zcont_kernel(const __grid_constant__ RenderParams P, unsigned tile0,
                                                                 unsigned ntile) {
  __shared__ ZCNode s_zc[kZcWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kTabN) s_T1[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / kTabN));
  __syncthreads();
  const unsigned tix = blockIdx.x * kZcWarps + warp;
  if (tix >= ntile) return;
  const ZTile t = P.ztiles[tile0 + tix];
  const int ray = t.ray;
  const bool lineok = lane < (int)t.nlt;
  const int l = (int)P.zlines[t.loff + (lineok ? lane : 0)];
  const LineDev *Lp = P.lines + l;
  const double c_src = __ldg(&Lp->c_src);
  const double cb_du = __ldg(&Lp->c_alp) * __ldg(&Lp->bdu), cb_ud = __ldg(&Lp->c_alp) * __ldg(&Lp->bud);
  const double knorm = 0.56419583546 / __ldg(&Lp->k_aa);
  const long long n0 = P.node_off[ray];
  const int N = (int)(P.node_off[ray + 1] - n0);
  const int ns = P.nstart ? max(1, min(__ldg(&P.nstart[ray]), N - 1)) : 1;
  double I = ns > 1 ? 0.0 : ((P.out_itype == 3) ? __ldg(&P.isrf_line[(size_t)l * P.nfr]) : __ldg(&Lp->i_outer));
  const NodeRec *__restrict__ rec = P.nodes.rec + n0;
  const double4 *__restrict__ cl = P.cellL + l;
  const double2 *__restrict__ cd = P.cellD + l;
  const uint32_t nl = (uint32_t)P.nl;
  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1);
  ZCNode *sn = s_zc[warp];
  const double vel0 = __ldg(P.velz);
  unsigned mbits = 0, xtra = 0;

  // interpolated cell values of this lane's line at a staged node: all four fields, or the dust pair only
  auto full_at = [&](const ZCNode *x) {
    const double4 a = ldg4(cl + x->offA), b = ldg4(cl + x->offB);
    if (x->icr == 3) return zvals(interp4(a, b, ldg4(cl + x->offC), ldg4(cl + x->offD), x->w2, x->w), c_src, cb_du, cb_ud);
    return zvals(interp2(a, b, x->w), c_src, cb_du, cb_ud);
  };
  auto dust2 = [&](const ZCNode *x, double2 a, double2 b) {
    const double w = x->w, w1 = 1.0 - w;
    return make_double2(lerp_rn(a.x, b.x, w, w1), lerp_rn(a.y, b.y, w, w1));
  };
  auto dust_at = [&](const ZCNode *x) {
    ZVal o;
    o.cN = o.kk = 0.0;
    const double2 a = __ldg(cd + x->offA);
    const double2 b = __ldg(cd + x->offB);
    double2 r;
    if (x->icr == 3) {
      const double2 c = __ldg(cd + x->offC);
      const double2 d = __ldg(cd + x->offD);
      const double t1 = 1.0 - x->w, r1 = 1.0 - x->w2;
      r.x = lerp_rn(lerp_rn(a.x, b.x, x->w, t1), lerp_rn(c.x, d.x, x->w, t1), x->w2, r1);
      r.y = lerp_rn(lerp_rn(a.y, b.y, x->w, t1), lerp_rn(c.y, d.y, x->w, t1), x->w2, r1);
    } else {
      r = dust2(x, a, b);
    }
    o.sd = r.x;
    o.ad = r.y;
    return o;
  };
  // one unflagged segment with the line terms (the arithmetic of zstep/ztile_kernel's grouped path)
  auto seg_full = [&](const ZVal &v0, const ZVal &v1, double nrm0, double nrm1, const ZCNode *x, double ep, double ec) {
    const double hds = x->hds;
    const double hn0 = hds * nrm0, hn1 = hds * nrm1;
    const double D = hds * (v0.ad + v1.ad), Th = hds * (v0.sd + v1.sd);
    const double Pq = hn0 * v0.kk, Q = hn1 * v1.kk, R = hn0 * v0.cN, S = hn1 * v1.cN;
    const bool neg = v1.kk < 0.0;
    const double dtau = fma(Pq, ep, fma(Q, ec, D)), theo = fma(R, ep, fma(S, ec, Th));
    const bool work = neg | gt_thin(dtau);
    if (!__any_sync(0xffffffffu, work)) {
      I = fma(I, 1.0 - dtau, theo);
    } else {
      const double K0 = v0.kk * nrm0, A0 = v0.cN * nrm0, K1 = v1.kk * nrm1, A1 = v1.cN * nrm1;
      const double alp1 = fma(K1, ec, v1.ad), src1 = fma(A1, ec, v1.sd);
      const double alp0 = fma(K0, ep, v0.ad), src0 = fma(A0, ep, v0.sd);
      double xx, q;
      step_onediv(alp0, src0, alp1, src1, dtau, theo, T1, xx, q);
      I = fma(I, xx, q);
      if (neg && (K1 * ec) * (hds + hds) < (double)(-0.01f)) mbits |= 1u;  // telescope.F:4295
    }
  };

  ZVal v0;
  v0.sd = v0.ad = v0.cN = v0.kk = 0.0;
  double nrm0 = 0.0;
  for (int c0 = ns; c0 < N; c0 += 30) {
    const int cnt = min(30, N - c0);
    __syncwarp();  // the previous batch is consumed
    {  // stage nodes c0-1 .. c0+cnt (the last one only as look-ahead for needfull) -> rows 0 .. cnt+1
      const int nstage = min(cnt + 2, N - (c0 - 1));
      uint32_t fl = 0;
      double e = 0.0;
      ZCNode x;
      if (lane < nstage) {
        const int node = c0 - 1 + lane;
        const double2 *p = reinterpret_cast<const double2 *>(rec + node);
        const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        const int4 cells = __ldg(reinterpret_cast<const int4 *>(p + 3));
        x.icr = ((uint32_t)cells.x >> kCellFlagShift) & kFlagIcrMask;
        x.offA = (uint32_t)(cells.x & kCellMask) * nl;
        x.offB = (uint32_t)(x.icr == 2 ? cells.z : cells.y) * nl;
        x.offC = (uint32_t)cells.z * nl;
        x.offD = (uint32_t)cells.w * nl;
        fl = ((uint32_t)cells.x >> kCellFlagShift) & ~kFlagIcrMask;
        if (!P.subgrid) fl &= ~kFlagSub;
        if (node == 1) fl |= kFlagInit;  // first segment of the ray: nothing carried yet
        if (node == 0) fl = 0;           // (the first node ends no segment)
        x.fl = fl;
        x.w = x.icr == 2 ? c.x : c.y;  // wr : wt
        x.w2 = c.x;
        x.hds = 0.5 * a.x;
        x.inv_lwav = b.y;
        x.lw = b.x;
        x.dvmu = a.y;
        x.ian = b.y * kIanScale;
        e = gauss_tab(fma(vel0, x.ian, -(a.y * x.ian)), T1, 0);
        x.e = e;
        x.pad = 0;
      }
      // the line terms of a node enter through its own profile value -- as the end point of its segment
      // (which a flag sends to zflagged with both end points in full) and as the start point of the next
      // segment, where the carried state multiplies this same value (line.F:4613-4615) unless that
      // segment is flagged too
      const uint32_t fl_next = __shfl_down_sync(0xffffffffu, fl, 1);
      const bool last = lane + 1 >= nstage;
      x.needfull = (e != 0.0 || fl != 0 || (last ? 1u : fl_next) != 0) ? 1u : 0u;
      if (lane < nstage) sn[lane] = x;
    }
    __syncwarp();
    if (c0 == ns) {  // values at the node the first integrated segment starts from
      v0 = sn[0].needfull ? full_at(sn) : dust_at(sn);
      nrm0 = knorm * sn[0].inv_lwav;
    }
    int s = 1;
    while (s <= cnt) {
      const ZCNode *x1 = sn + s;
      if (s < cnt && (x1->fl | x1->needfull | x1[1].fl | x1[1].needfull | x1[-1].needfull) == 0 &&
          (x1->icr != 3) && (x1[1].icr != 3)) {
        // two dust-only segments at once (nodes s-1, s, s+1 carry no line terms)
        const ZCNode *x2 = x1 + 1;
        const double2 a1 = __ldg(cd + x1->offA);
        const double2 b1 = __ldg(cd + x1->offB);
        const double2 a2 = __ldg(cd + x2->offA);
        const double2 b2 = __ldg(cd + x2->offB);
        const double2 d1 = dust2(x1, a1, b1), d2 = dust2(x2, a2, b2);
        const double h1 = x1->hds, h2 = x2->hds;
        const double D1 = h1 * (v0.ad + d1.y), T1h = h1 * (v0.sd + d1.x);
        const double D2 = h2 * (d1.y + d2.y), T2h = h2 * (d1.x + d2.x);
        const bool work = gt_thin(D1) | gt_thin(D2);
        if (!__any_sync(0xffffffffu, work)) {
          I = fma(fma(I, 1.0 - D1, T1h), 1.0 - D2, T2h);
        } else {
          double xa, qa, xb, qb;
          step_onediv(v0.ad, v0.sd, d1.y, d1.x, D1, T1h, T1, xa, qa);
          step_onediv(d1.y, d1.x, d2.y, d2.x, D2, T2h, T1, xb, qb);
          I = fma(fma(I, xa, qa), xb, qb);
        }
        v0.sd = d2.x;
        v0.ad = d2.y;
        v0.cN = v0.kk = 0.0;
        nrm0 = knorm * x2->inv_lwav;
        s += 2;
        continue;
      }
      const ZVal v1 = x1->needfull ? full_at(x1) : dust_at(x1);
      const double nrm1 = knorm * x1->inv_lwav;
      if (x1->fl == 0) {
        seg_full(v0, v1, nrm0, nrm1, x1, x1[-1].e, x1->e);
      } else {
        double tmp[1] = {I};
        ZSeg sg;
        sg.ds = x1->hds + x1->hds;
        sg.lwav = 0.5 * (x1[-1].lw + x1->lw);
        sg.dv0 = x1[-1].dvmu;
        sg.dv1 = x1->dvmu;
        sg.ian = x1->ian;
        sg.nrm0 = nrm0;
        sg.nrm1 = nrm1;
        sg.v0 = v0;
        sg.v1 = v1;
        const uint32_t ep_a = (uint32_t)__cvta_generic_to_shared(&x1[-1].e);
        const uint32_t ec_a = (uint32_t)__cvta_generic_to_shared(&x1->e);
        mbits |= zflagged(tmp, sg, x1->fl, 1, ep_a, ec_a, Lp, P.line_dnu + (size_t)l * P.nfr,
                          P.star_line + (size_t)l * P.nfr, P.velz, 0, 1, 0, 1, P.starfract, lineok ? 1u : 0u, xtra);
        I = tmp[0];
      }
      v0 = v1;
      nrm0 = nrm1;
      s++;
    }
  }
  unsigned long long r = 0;
  if (lineok) {
    const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
    P.img[row * P.nfr] = I;
    if (P.sparse && I == 0.0) P.dense[(long long)ray * P.nl + l] = 2;  // see fill_sparse_kernel
    if (P.integ) P.integ[row * P.nfr] = 1;  // channel 0 is always a masked one (telescope.F:548)
    if (mbits) atomicOr(&P.maser[l], 1);
    r = 1;
  }
  unsigned long long sct = r * (unsigned long long)(N > 0 ? N - 1 : 0), e = sct + (lineok ? xtra : 0u);
  unsigned long long ex = r * (unsigned long long)(N > ns ? N - ns : 0) + (lineok ? xtra : 0u);
  for (int o = 16; o; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    sct += __shfl_xor_sync(0xffffffffu, sct, o);
    ex += __shfl_xor_sync(0xffffffffu, ex, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if (lane == 0 && r) {
    atomicAdd(&P.counters[0], r);
    atomicAdd(&P.counters[1], e);
    atomicAdd(&P.counters[2], sct);
    atomicAdd(&P.counters[3], ex);
  }
}

// ------------------------------------------------------------------------------------------
// chan_kernel: the formal solution with the CHANNELS of one line across the lanes of a warp -- renders with
// fewer than 8 lines per batch (BASELINE configs 1 and 3: one line, up to 199 channels per ray), where there
// are no lines to put side by side and nothing to share between them.
//
// One warp = one ray x one line x up to 32 x kCCw consecutive channels (lane = channel, kCCw blocks of 32
// channels per thread; intensity and the profile value at the previous node in registers); the tiles are
// zplan_kernel's with one line per tile.  Everything that does not depend on the channel is the same for all
// lanes, so it is computed ONCE per (ray, line, node) -- by the lanes acting as nodes: per batch of 31
// segments lane k loads node record k, gathers and interpolates its stencil cells (the 32 gathers of a batch
// in flight together), takes the previous node's values from lane k-1 and leaves in shared memory (CSeg)
//   * the six constants of  dtau = D + P e0 + Q e1 ,  theomax = Th + R e0 + S e1  (e0, e1: profile at the
//     two nodes),
//   * the complete step  I <- I xd + qd  of a channel that sees no line at either node (e0 = e1 = 0: the
//     dust-only segment, identical for all such channels).
// The lanes then walk the batch as channels.  Per node and block of 32 channels: the profile argument
// (v_k/c - Omega.v/c) / width (one FMA; formed in velocity like ztile_kernel's) tells whether the channel is
// beyond the point where the kernels flush the profile to 0 (exp(-345)); if that holds for the whole block at
// both nodes -- the rule in a cube, where a node's line covers a few of the ~200 channels -- the block takes
// the staged dust-only step: one FMA per channel.  Otherwise one Gaussian per channel (table-based exp) and the
// step; the case split of transfer.F:1517,1542 needs no vote at the node level (D + |P| + |Q| <= 1e-9 is a
// property of the node: CSeg.thin), and where every opacity of the two nodes is positive (CSeg.upos, a
// property of the node as well) qdr_src_2's selections on the sign of the opacities drop out.  The arithmetic
// per (ray, line, channel) is that of ztile_kernel (same functions, same operation order): the two kernels
// give the same bits.  Flagged nodes (first segment, inner hole / star mixing, 6q > 1 sub-grid candidates)
// go through chan_flagged, out of line.
// ------------------------------------------------------------------------------------------
struct __align__(16) CSeg {  // segment ending at a staged node
  double D, Th, Pq, Q, R, S;
  double ian, dvi;           // scaled reciprocal Doppler width of the segment ; Omega.v/c at its end node times it
  double xd, qd;             // the dust-only step of this segment
  uint32_t fl, thin, neg, upos;
};
struct __align__(16) CNode {  // staged node
  double sd, ad, A1, K1;      // dust pair ; c_src N_up and kk, times the profile norm of the segment ending here
  double cN, kk, nrm, lw;     // (flagged path) c_src N_up ; kk ; that norm ; line width
  double dvmu, ds;
};
// Flagged segment of chan_kernel's line for the cw channel blocks of a thread (Ic, and the profile values epl /
// ecl at the two nodes, in local memory): zflagged's arithmetic per channel.  Returns the maser bits; xtra += the
// extra element integrations of sub-gridded channels.
__device__ __noinline__ unsigned chan_flagged(double *Ic, const double *epl, const double *ecl, const ZSeg &g,
                                              uint32_t fl, int cw, const LineDev *__restrict__ L,
                                              const double *__restrict__ dnu_l, const double *__restrict__ star_l,
                                              const double *__restrict__ velo, int jb, int jmax, int cmin,
                                              double starfract, unsigned realbits, unsigned &xtra) {
  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1);
  const double nu0 = __ldg(&L->nu0), k_aa = __ldg(&L->k_aa), inv_nu0 = __ldg(&L->inv_nu0);
  const double ds = g.ds;
  const int init = (fl & (kFlagInit | kFlagStar | kFlagZero)) ? 1 : 0;
  unsigned mb = 0;
  for (int c = 0; c < cw; c++) {
    const int j = min(jb + c * 32, jmax);
    const int ch = j ? cmin + j - 1 : 0;
    const double dnu = __ldg(dnu_l + ch);
    const double ep = epl[c], ec = ecl[c];
    double inten = Ic[c];
    if (fl & kFlagZero) inten = 0.0;
    if (fl & kFlagStar) inten = (1.0 - starfract) * inten + starfract * __ldg(star_l + ch);
    double srcl0 = g.v0.cN * (g.nrm0 * ep), alpl0 = g.v0.kk * (g.nrm0 * ep);  // carried state (line.F:4613-4615)
    bool need = false, ms = false;
    double sleft = 0.0, sright = 0.0;
    if (fl & kFlagSub) {  // line.F:4706-4745
      const double q = fabs((g.dv1 - g.dv0) / (g.lwav / 2.99792458e5));
      const double s_c = ds * (dnu * inv_nu0 - g.dv0) / (g.dv1 - g.dv0);
      const double dls3 = 3.0 * (ds / q);
      sright = s_c + dls3;
      sleft = s_c - dls3;
      need = sright > 0.0 && sleft < ds;
    }
    if (!need) {
      if (init) {  // line.F:4559-4586: the start point with this segment's width
        const double vel = __ldg(velo + ch);
        const double e0 = gauss_tab(fma(vel, g.ian, -(g.dv0 * g.ian)), T1, 0);
        srcl0 = g.v0.cN * (g.nrm1 * e0);
        alpl0 = g.v0.kk * (g.nrm1 * e0);
      }
      const double src0 = g.v0.sd + srcl0, alp0 = g.v0.ad + alpl0;
      const double alpl1 = g.v1.kk * (g.nrm1 * ec);
      const double src1 = fma(g.v1.cN, g.nrm1 * ec, g.v1.sd), alp1 = g.v1.ad + alpl1;
      const double hds = 0.5 * ds;
      const double dtau = hds * (alp0 + alp1), theomax = hds * (src0 + src1);
      double x, q;
      step_onediv(alp0, src0, alp1, src1, dtau, theomax, T1, x, q);
      inten = fma(inten, x, q);
      ms = alpl1 * ds < (double)(-0.01f);
    }
    else {
      const int n = subgrid_tile(nu0, k_aa, dnu, inten, ds, sleft, sright, g.v0.sd, g.v0.ad, g.v0.cN, g.v0.kk, g.dv0,
                                 g.v1.sd, g.v1.ad, g.v1.cN, g.v1.kk, g.dv1, g.lwav, srcl0, alpl0, init);
      ms = alpl0 * ds < (double)(-0.01f);
      if ((realbits >> c) & 1u) xtra += (unsigned)(n - 1);
    }
    if (ms) mb |= 1u << c;
    Ic[c] = inten;
  }
  return mb;
}

#ifndef RL_CCW
#define RL_CCW 3
#endif
constexpr int kCCw = RL_CCW;  // blocks of 32 channels per chan_kernel thread
#ifndef RL_CMINB
#define RL_CMINB 24  // resident blocks (= warps) per SM the register budget is set for
#endif
template <int CW>
__global__ void __launch_bounds__(32, RL_CMINB) chan_kernel(const __grid_constant__ RenderParams P) {
  __shared__ CSeg s_seg[32];
  __shared__ CNode s_cn[32];
  const int lane = threadIdx.x;
  if (lane < kTabN) s_T1[lane] = exp2((double)lane * (1.0 / kTabN));
  __syncwarp();
  const unsigned tix = blockIdx.x;
  if (tix >= P.nztile) return;
  const ZTile t = P.ztiles[tix];
  const int ray = t.ray;
  const int l = (int)P.zlines[t.loff];
  const int nchk = t.nchk, j0 = t.j0, cmin = t.cmin;
  const int cw = (nchk + 31) >> 5;  // channel blocks in use
  const int jmax = j0 + nchk - 1, jb = j0 + lane;
  // channel of my slot c: position min(jb + 32 c, jmax) of the list {0, cmin, cmin + 1, ...}
  auto chan_of = [&](int c) {
    const int j = min(jb + c * 32, jmax);
    return j ? cmin + j - 1 : 0;
  };
  const int4 rg = P.rng[(long long)ray * P.nl + l];
  unsigned realbits = 0;  // slots that are channels the reference integrates for this line
#pragma unroll
  for (int c = 0; c < CW; c++) {
    if (c * 32 + lane < nchk) {
      const int ch = chan_of(c);
      if (ch == 0 || (ch >= rg.x && ch <= rg.y) || ch == rg.z) realbits |= 1u << c;
    }
  }
  const LineDev *Lp = P.lines + l;
  const double c_src = __ldg(&Lp->c_src);
  const double cb_du = __ldg(&Lp->c_alp) * __ldg(&Lp->bdu), cb_ud = __ldg(&Lp->c_alp) * __ldg(&Lp->bud);
  const double knorm = 0.56419583546 / __ldg(&Lp->k_aa);
  double I[CW], ep[CW], vel[CW];
#pragma unroll
  for (int c = 0; c < CW; c++) {
    const int ch = chan_of(c);
    I[c] = (P.out_itype == 3) ? __ldg(&P.isrf_line[(size_t)l * P.nfr + ch]) : __ldg(&Lp->i_outer);
    vel[c] = __ldg(P.velz + ch);
    ep[c] = 0.0;
  }
  const uint32_t T1 = (uint32_t)__cvta_generic_to_shared(s_T1);
  const long long n0 = P.node_off[ray];
  const int N = (int)(P.node_off[ray + 1] - n0);
  const NodeRec *__restrict__ rec = P.nodes.rec + n0;
  const double4 *__restrict__ cl = P.cellL + l;
  const size_t nl = (size_t)P.nl;
  // first segment to integrate: everything before it lies behind an opaque dust wall (wall_kernel)
  const int ns = P.nstart ? max(1, min(__ldg(&P.nstart[ray]), N - 1)) : 1;
  if (ns > 1) {
#pragma unroll
    for (int c = 0; c < CW; c++) I[c] = 0.0;
  }
  unsigned mbits = 0, xtra = 0;

  for (int c0 = ns; c0 < N; c0 += 31) {
    const int cnt = min(31, N - c0);
    __syncwarp();  // the previous batch is consumed
    {  // lanes as nodes c0-1 .. c0+cnt-1
      ZVal v;
      v.sd = v.ad = v.cN = v.kk = 0.0;
      double nrm = 0.0, hds = 0.0, ian = 0.0, dvi = 0.0, lw = 0.0, dvmu = 0.0, ds = 0.0;
      uint32_t fl = 0;
      if (lane <= cnt) {
        const int node = c0 - 1 + lane;
        const double2 *p = reinterpret_cast<const double2 *>(rec + node);
        const double2 a = __ldg(p), b = __ldg(p + 1), w = __ldg(p + 2);
        const int4 cells = __ldg(reinterpret_cast<const int4 *>(p + 3));
        const uint32_t icr = ((uint32_t)cells.x >> kCellFlagShift) & kFlagIcrMask;
        const double4 ca = ldg4(cl + (size_t)(cells.x & kCellMask) * nl);
        const double4 cb = ldg4(cl + (size_t)(icr == 2 ? cells.z : cells.y) * nl);
        double4 iv;
        if (icr == 3) {
          const double4 cc = ldg4(cl + (size_t)cells.z * nl), cd = ldg4(cl + (size_t)cells.w * nl);
          iv = interp4(ca, cb, cc, cd, w.x, w.y);
        } else {
          iv = interp2(ca, cb, icr == 2 ? w.x : w.y);
        }
        v = zvals(iv, c_src, cb_du, cb_ud);
        fl = ((uint32_t)cells.x >> kCellFlagShift) & ~kFlagIcrMask;
        if (!P.subgrid) fl &= ~kFlagSub;
        if (node == 1) fl |= kFlagInit;  // first segment of the ray: nothing carried yet
        ds = a.x;
        dvmu = a.y;
        lw = b.x;
        hds = 0.5 * ds;
        nrm = knorm * b.y;
        ian = b.y * kIanScale;
        dvi = dvmu * ian;
      }
      ZVal v0;
      v0.sd = __shfl_up_sync(0xffffffffu, v.sd, 1);
      v0.ad = __shfl_up_sync(0xffffffffu, v.ad, 1);
      v0.cN = __shfl_up_sync(0xffffffffu, v.cN, 1);
      v0.kk = __shfl_up_sync(0xffffffffu, v.kk, 1);
      const double nrm0 = __shfl_up_sync(0xffffffffu, nrm, 1);
      if (lane <= cnt) {
        CNode n;
        n.sd = v.sd; n.ad = v.ad; n.A1 = v.cN * nrm; n.K1 = v.kk * nrm;
        n.cN = v.cN; n.kk = v.kk; n.nrm = nrm; n.lw = lw;
        n.dvmu = dvmu; n.ds = ds;
        s_cn[lane] = n;
        CSeg g;
        const double hn0 = hds * nrm0, hn1 = hds * nrm;
        g.D = hds * (v0.ad + v.ad);
        g.Th = hds * (v0.sd + v.sd);
        g.Pq = hn0 * v0.kk;
        g.Q = hn1 * v.kk;
        g.R = hn0 * v0.cN;
        g.S = hn1 * v.cN;
        g.ian = ian;
        g.dvi = dvi;
        step_onediv(v0.ad, v0.sd, v.ad, v.sd, g.D, g.Th, T1, g.xd, g.qd);
        g.fl = fl;
        g.neg = v.kk < 0.0 ? 1u : 0u;  // inverted populations: full path, which carries the maser test
        g.thin = (!g.neg && !gt_thin(g.D + fabs(g.Pq) + fabs(g.Q))) ? 1u : 0u;
        // every opacity of the segment is positive whatever the profile (it is >= 0): dust opacity positive
        // at both nodes, no inversion at either
        g.upos = (!g.neg && v0.ad > kAlpMin && v.ad > kAlpMin && !(v0.kk * nrm0 < 0.0)) ? 1u : 0u;
        s_seg[lane] = g;
      }
    }
    __syncwarp();
    if (c0 == ns) {  // profile at the node the first integrated segment starts from (with its own width)
      const double ian = s_seg[0].ian, dvi = s_seg[0].dvi;
#pragma unroll
      for (int c = 0; c < CW; c++) ep[c] = gauss_tab(fma(vel[c], ian, -dvi), T1, 0);
    }
    for (int s = 1; s <= cnt; s++) {
      const CSeg *g = s_seg + s;
      const uint32_t fl = g->fl;
      const double ian = g->ian, dvi = g->dvi;
      if (fl == 0) {
#pragma unroll
        for (int c = 0; c < CW; c++) {
          if (c < cw) {
            const double u = fma(vel[c], ian, -dvi);
            const bool nolin = (((unsigned)__double2hiint(u) & 0x7fffffffu) > kHiUmax) & (ep[c] == 0.0);
            if (__all_sync(0xffffffffu, nolin)) {
              I[c] = fma(I[c], g->xd, g->qd);  // (the profile stays 0)
            } else {
              const double ec = gauss_tab(u, T1, 0);
              const double dtau = fma(g->Pq, ep[c], fma(g->Q, ec, g->D));
              const double theo = fma(g->R, ep[c], fma(g->S, ec, g->Th));
              if (g->thin || !__any_sync(0xffffffffu, gt_thin(dtau))) {
                // transfer.F:1522-1524,1545: Q = theomax, xp = 1 - dtau
                I[c] = fma(I[c], 1.0 - dtau, theo);
              } else {
                const CNode *n0p = s_cn + (s - 1), *n1p = s_cn + s;
                const double alp1 = fma(n1p->K1, ec, n1p->ad), src1 = fma(n1p->A1, ec, n1p->sd);
                const double alp0 = fma(n0p->K1, ep[c], n0p->ad), src0 = fma(n0p->A1, ep[c], n0p->sd);
                double x, q;
                if (g->upos) {
                  // step_onediv with both opacities known to be positive
                  const double xpe = expneg_tab(dtau, T1, 0);
                  const double e0 = 1.0 - xpe, e1 = dtau - e0;
                  const bool thick = gt_mid(dtau);
                  const double hb = 0.5 * dtau;
                  const double ca = thick ? fma(e0, dtau, -e1) : hb, cb = thick ? e1 : hb, dd = thick ? dtau : 1.0;
                  const double den = dd * (alp0 * alp1);
                  const double num = fma(ca, src0 * alp1, cb * (src1 * alp0));
                  x = thick ? xpe : (1.0 - dtau);
                  q = div_fast(num, den);
                  q = gt_thin(dtau) ? min_sel(q, theo) : theo;
                } else {
                  step_onediv(alp0, src0, alp1, src1, dtau, theo, T1, x, q);
                  if (g->neg && (n1p->K1 * ec) * n1p->ds < (double)(-0.01f)) mbits |= 1u << c;  // telescope.F:4295
                }
                I[c] = fma(I[c], x, q);
              }
              ep[c] = ec;
            }
          }
        }
      } else {
        double tmp[CW], epl[CW], ecl[CW];
#pragma unroll
        for (int c = 0; c < CW; c++) {
          tmp[c] = I[c];
          epl[c] = ep[c];
          ecl[c] = gauss_tab(fma(vel[c], ian, -dvi), T1, 0);
        }
        const CNode *n0p = s_cn + (s - 1), *n1p = s_cn + s;
        ZSeg sg;
        sg.ds = n1p->ds;
        sg.lwav = 0.5 * (n0p->lw + n1p->lw);
        sg.dv0 = n0p->dvmu;
        sg.dv1 = n1p->dvmu;
        sg.ian = ian;
        sg.nrm0 = n0p->nrm;
        sg.nrm1 = n1p->nrm;
        sg.v0.sd = n0p->sd; sg.v0.ad = n0p->ad; sg.v0.cN = n0p->cN; sg.v0.kk = n0p->kk;
        sg.v1.sd = n1p->sd; sg.v1.ad = n1p->ad; sg.v1.cN = n1p->cN; sg.v1.kk = n1p->kk;
        mbits |= chan_flagged(tmp, epl, ecl, sg, fl, cw, Lp, P.line_dnu + (size_t)l * P.nfr,
                              P.star_line + (size_t)l * P.nfr, P.velz, jb, jmax, cmin, P.starfract, realbits, xtra);
#pragma unroll
        for (int c = 0; c < CW; c++) {
          I[c] = tmp[c];
          ep[c] = ecl[c];
        }
      }
    }
  }
  // results: only the slots that are channels the reference integrates for this line
  unsigned long long r = 0;
  const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
#pragma unroll
  for (int c = 0; c < CW; c++) {
    if ((realbits >> c) & 1u) {
      const int ch = chan_of(c);
      P.img[row * P.nfr + ch] = I[c];
      if (P.sparse && I[c] == 0.0) P.dense[(long long)ray * P.nl + l] = 2;  // see fill_sparse_kernel
      if (P.integ) {
        const bool masked = P.nonredundant ? (ch != rg.z) : (ch == 0);  // telescope.F:548,575
        P.integ[row * P.nfr + ch] = masked ? 1 : 2;
      }
      r++;
    }
  }
  if (mbits & realbits) atomicOr(&P.maser[l], 1);
  unsigned long long sct = r * (unsigned long long)(N > 0 ? N - 1 : 0), e = sct + xtra;
  unsigned long long ex = r * (unsigned long long)(N > ns ? N - ns : 0) + xtra;
  for (int o = 16; o; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    sct += __shfl_xor_sync(0xffffffffu, sct, o);
    ex += __shfl_xor_sync(0xffffffffu, ex, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if (lane == 0 && r) {
    atomicAdd(&P.counters[0], r);
    atomicAdd(&P.counters[1], e);
    atomicAdd(&P.counters[2], sct);
    atomicAdd(&P.counters[3], ex);
  }
}

// tiles of ztile_kernel, one thread per ray.  The lines that carry a channel window on this ray (in index
// order, compacted into zlines) are cut into groups of zlw; the channel list of a group is
// {0} + [min lo, max hi] over its lines (a superset of every line's own item channels -- the kernel
// stores only those), cut into chunks when a thread would get more than kZCw channels.  Lines that only
// need channel 0 (and possibly their first skipped channel) follow, 32 per tile.  FILL = false counts the
// tiles of every ray, FILL = true (after the scan) writes them.
template <bool FILL>
__global__ void zplan_kernel(RenderParams P) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray > P.nray) return;
  unsigned n = 0, ncont = 0;
  if (ray < P.nray) {
    const int4 *rg = P.rng + (size_t)ray * P.nl;
    const unsigned *ni = P.nitems + (size_t)ray * P.nl;
    unsigned short *zl = P.zlines + (size_t)ray * P.nl;
    ZTile *out = FILL ? (P.ztiles_in ? P.ztiles_in : P.ztiles) + P.cta_off[ray] : nullptr;
    ZTile *outc = FILL ? P.ztiles + P.cta_off[P.nray + 1 + ray] : nullptr;
    unsigned *outk = (FILL && P.zkeys) ? P.zkeys + P.cta_off[ray] : nullptr;
    unsigned steps = 0;  // segments the ray's tiles walk (behind the opaque-wall start)
    if (FILL && P.zkeys) {
      const int N = (int)(P.node_off[ray + 1] - P.node_off[ray]);
      const int ns = P.nstart ? max(1, min(P.nstart[ray], N - 1)) : 1;
      steps = (unsigned)max(0, N - ns);
    }
    auto emit = [&](int start, int cnt, int cmin, int cmax, bool cont) {
      if (cmax < cmin) { cmin = 1; cmax = 0; }
      const int nch = 1 + (cmax - cmin + 1);
      int lws = 0;
      while ((1 << lws) < cnt) lws++;
      const int maxch = (32 >> lws) * (P.zlw == 1 ? kCCw : kZCw);
      const int nchunk = (nch + maxch - 1) / maxch, per = (nch + nchunk - 1) / nchunk;
      for (int j0 = 0; j0 < nch; j0 += per) {
        if (FILL) {
          ZTile t;
          t.ray = ray;
          t.loff = (unsigned)((size_t)ray * P.nl + start);
          t.cmin = (unsigned short)cmin;
          t.j0 = (unsigned short)j0;
          t.nchk = (unsigned short)min(per, nch - j0);
          t.nlt = (unsigned char)cnt;
          t.lwshift = (unsigned char)lws;
          if (cont) outc[ncont] = t;
          else {
            out[n] = t;
            if (outk) {  // key: 4095 - work / 4, work = segments x (per-node part + channels per thread)
              const unsigned cwt = ((unsigned)t.nchk + (32u >> lws) - 1u) >> (5 - lws);
              outk[n] = 4095u - min(4095u, (steps * (2u + cwt)) >> 2);
            }
          }
        }
        if (cont) ncont++;
        else n++;
      }
    };
    // a ray whose lines all fit kZCw channels (narrow lines, coarse channels) gets 32 lines per tile, one channel
    // group: the per-(node, line) part of ztile_kernel is then shared by twice as many channels
    int gsz0 = P.zlw;
    if (P.zlw > 1 && P.zlw < 32) {
      int cmn = 0x7fffffff, cmx = -1;
      for (int l = 0; l < P.nl; l++) {
        if (!ni[l]) continue;
        const int4 r = rg[l];
        if (r.y < r.x) continue;
        cmn = min(cmn, r.x);
        cmx = max(cmx, r.y);
        if (r.z >= 0) { cmn = min(cmn, r.z); cmx = max(cmx, r.z); }
      }
      if (cmx >= cmn && 2 + (cmx - cmn) <= kZCw) gsz0 = 32;
    }
    int na = 0;
    // pass 0: lines with a channel window; pass 1: lines that need channel 0 only (zcont_kernel's tiles);
    // pass 2: lines without a window whose channel 0 lies inside the ray's velocity span, so that their
    // first skipped channel is integrated as well (telescope.F:557-612; rare)
    for (int pass = 0; pass < 3; pass++) {
      const int gsz = pass == 0 ? gsz0 : (pass == 1 || P.zlw > 1) ? 32 : 1;  // (zlw = 1, chan_kernel: one line per tile)
      int gstart = na, cmin = 0x7fffffff, cmax = -1;
      for (int l = 0; l < P.nl; l++) {
        if (!ni[l]) continue;
        const int4 r = rg[l];
        const int cls = (r.y >= r.x) ? 0 : (r.z < 0 ? 1 : 2);
        if (cls != pass) continue;
        if (FILL) zl[na] = (unsigned short)l;
        if (r.y >= r.x) { cmin = min(cmin, r.x); cmax = max(cmax, r.y); }
        if (r.z >= 0) { cmin = min(cmin, r.z); cmax = max(cmax, r.z); }
        na++;
        if (na - gstart == gsz) {
          emit(gstart, gsz, cmin, cmax, pass == 1);
          gstart = na; cmin = 0x7fffffff; cmax = -1;
        }
      }
      if (na > gstart) emit(gstart, na - gstart, cmin, cmax, pass == 1);
    }
  }
  if (!FILL) {
    P.ncta[ray] = n;
    P.ncta[P.nray + 1 + ray] = ncont;
  }
}

// ------------------------------------------------------------------------------------------
// Opaque-wall start.  What a ray has accumulated before it enters an optically thick dust layer leaves that
// layer multiplied by exp(-tau); once that is far below the rounding of what the layer itself emits towards
// the observer, the segments behind it need not be integrated.  wall_kernel finds, per ray and for all lines
// of the batch at once, the first segment n from which on this provably holds:
//   dropped  <=  Smax exp(-tau_n)          Smax: the largest source function of any cell and line of the batch
//                                          (and the outer boundary intensity); qdr_src_2 never raises the
//                                          intensity above max(I, S1, S2) (transfer.F:1498-1571, a + b = 1 - xp)
//   kept     >=  Smin_n (1 - exp(-tau_n))  Smin_n: the smallest source function, over lines and over the line
//                                          profile (0 .. its peak value), of the cells the nodes n-1 .. N-1
//                                          interpolate from; every step adds at least (1 - xp) min(S1, S2)
//   tau_n    =   dust optical depth of segments n .. N-1, from the smallest dust opacity any line of the batch
//                has in a cell, interpolated with the nodes' own weights (all in [0, 1]); line opacity only adds
// and cuts where  tau_n - ln(Smax / Smin_n) > wall_tau  (default 64: the neglected part is below e^-64 ~ 1e-28
// of the result).  A batch with an inverted level pair or a negative dust opacity anywhere is never shortened,
// and a ray whose front crosses a cell without emission (T_dust = 0, or no dust at all) keeps its full
// length: Smin = 0 there.  The reference integrates every segment (telescope.F:4079-4300 has no such cut); its
// work counters R, E, S are reported unchanged (wallcount_kernel adds the sub-grid steps of the skipped
// segments), the executed element integrations separately.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_pos(double *addr, double v) {  // v >= 0: bit patterns order like values
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}
// wstat: {Smax (bit pattern of a non-negative double), inverted flag}
__global__ void __launch_bounds__(256) wallprep_kernel(RenderParams P, double lw_min, double *admin, double *smin,
                                                       unsigned long long *wstat) {
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= P.ncell) return;
  double mn = 1.0e300, lo = 1.0e300, hi = 0.0;
  bool inv = false;
  for (int l = 0; l < P.nl; l++) {
    const double4 v = ldg4(P.cellL + (size_t)cell * P.nl + l);
    const LineDev *L = P.lines + l;
    mn = fmin(mn, v.y);
    const double kk = __ldg(&L->c_alp) * (v.w * __ldg(&L->bdu) - v.z * __ldg(&L->bud));
    const double cN = __ldg(&L->c_src) * v.z;
    inv = inv || (kk < 0.0) || !(v.y >= 0.0) || !(v.x >= 0.0);
    // source function over the profile range 0 .. phimax: monotone in phi, so its extremes sit at the ends
    const double phimax = 0.56419583546 / (__ldg(&L->k_aa) * lw_min);
    const double s0 = v.y > 0.0 ? v.x / v.y : (v.x > 0.0 ? 1.0e300 : 0.0);
    const double a1 = v.y + kk * phimax;
    const double s1 = a1 > 0.0 ? (v.x + cN * phimax) / a1 : s0;
    lo = fmin(lo, v.y > 0.0 ? fmin(s0, s1) : 0.0);  // a cell without dust opacity gives no bound
    hi = fmax(hi, fmax(s0, s1));
    if (cell == 0) {  // outer boundary intensity of the line (telescope.F:3989-4028)
      double ib = __ldg(&L->i_outer);
      if (P.out_itype == 3)
        for (int k = 0; k < P.nfr; k++) ib = fmax(ib, __ldg(&P.isrf_line[(size_t)l * P.nfr + k]));
      hi = fmax(hi, ib);
    }
  }
  admin[cell] = mn;
  smin[cell] = lo;
  if (!(hi >= 0.0) || hi > 1.0e300) hi = 1.0e300;
  atomic_max_pos(reinterpret_cast<double *>(&wstat[0]), hi);
  if (inv) atomicOr(&wstat[1], 1ull);
}

// one warp per ray: 32 segments at a time from the observer's end, the running optical depth and the running
// smallest source function as warp scans
__global__ void __launch_bounds__(128) wall_kernel(RenderParams P, const double *__restrict__ admin,
                                                   const double *__restrict__ smin,
                                                   const unsigned long long *__restrict__ wstat, int *nstart) {
  const int ray = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (ray >= P.nray) return;
  int ns = 1;
  const long long n0 = P.node_off[ray];
  const int N = (int)(P.node_off[ray + 1] - n0);
  const double smax = __longlong_as_double((long long)__ldg(&wstat[0]));
  if (ray > 0 && N > 2 && P.wall_tau > 0.0 && !__ldg(&wstat[1]) && smax < 1.0e299) {
    const NodeRec *rec = P.nodes.rec + n0;
    const double lsmax = smax > 0.0 ? log(smax) : -1.0e300;
    // lower bounds at a node: dust opacity (interpolated) and source function (smallest of the stencil cells)
    auto at = [&](int n, double &amin, double &sm) {
      const Node nd = load_node(P.nodes.rec, n0 + n);
      const int icr = nd.flags & kFlagIcrMask;
      const double a = __ldg(&admin[nd.cells.x]);
      sm = __ldg(&smin[nd.cells.x]);
      if (icr == 1) {
        amin = (1.0 - nd.wt) * a + nd.wt * __ldg(&admin[nd.cells.y]);
        sm = fmin(sm, __ldg(&smin[nd.cells.y]));
      } else if (icr == 2) {
        amin = (1.0 - nd.wr) * a + nd.wr * __ldg(&admin[nd.cells.z]);
        sm = fmin(sm, __ldg(&smin[nd.cells.z]));
      } else {
        amin = (1.0 - nd.wr) * ((1.0 - nd.wt) * a + nd.wt * __ldg(&admin[nd.cells.y])) +
               nd.wr * ((1.0 - nd.wt) * __ldg(&admin[nd.cells.z]) + nd.wt * __ldg(&admin[nd.cells.w]));
        sm = fmin(fmin(sm, __ldg(&smin[nd.cells.y])), fmin(__ldg(&smin[nd.cells.z]), __ldg(&smin[nd.cells.w])));
      }
    };
    double cum0 = 0.0, sfront0;  // optical depth / smallest source function of everything in front of the chunk
    {
      double a;
      at(N - 1, a, sfront0);
    }
    for (int top = N - 1; top >= 2; top -= 32) {  // lane j: segment n = top - j, joining nodes n-1 and n
      const int n = top - lane;
      double dtau = 0.0, sm = 1.0e300;
      if (n >= 2) {
        double a0, a1, s1;
        at(n, a1, s1);
        at(n - 1, a0, sm);
        dtau = 0.5 * __ldg(&rec[n].ds) * (a0 + a1);
      }
      double cum = dtau, sf = sm;
      for (int o = 1; o < 32; o <<= 1) {  // inclusive scans over the lanes (= towards the far end of the ray)
        const double c = __shfl_up_sync(0xffffffffu, cum, o), m = __shfl_up_sync(0xffffffffu, sf, o);
        if (lane >= o) {
          cum += c;
          sf = fmin(sf, m);
        }
      }
      cum += cum0;
      sf = fmin(sf, sfront0);
      const bool dark = n >= 2 && !(sf > 0.0);  // no emission bound in front: the ray keeps its full length
      const bool hit = n >= 2 && !dark && cum > P.wall_tau && cum - (lsmax - log(sf)) > P.wall_tau;
      const unsigned stop = __ballot_sync(0xffffffffu, dark || hit);
      if (stop) {
        const int j = __ffs(stop) - 1;
        if ((__ballot_sync(0xffffffffu, hit) >> j) & 1u) ns = top - j;
        break;
      }
      cum0 = __shfl_sync(0xffffffffu, cum, 31);
      sfront0 = __shfl_sync(0xffffffffu, sf, 31);
    }
  }
  if (lane == 0) nstart[ray] = ns;
}

// The element integrations of the sub-gridded segments that the opaque-wall start skips, added to the E
// counter so that it stays the reference's count (line.F:4745-4833: a sub-gridded segment takes one
// integration per sub-point strictly inside it plus one).  One block per ray: threads first act as channels
// (the number of sub-points is a function of segment and channel alone -- the velocity grid is the same for
// every line), then as lines (sum over the channels the reference integrates for that line).
__global__ void __launch_bounds__(128) wallcount_kernel(RenderParams P, const int *__restrict__ nstart, int tile_shift) {
  extern __shared__ unsigned s_x[];  // [nfr + 1] extra integrations per channel, then their prefix sums
  const int ray = blockIdx.x;
  const long long n0 = P.node_off[ray];
  const int N = (int)(P.node_off[ray + 1] - n0);
  int first = max(1, min(nstart[ray], N - 1));             // first segment ztile_kernel / zcont_kernel integrate
  if (tile_shift) first = max(0, min(nstart[ray], N - 1) - 2) + 1;  // tile_kernel starts one segment earlier
  if (first <= 1 || !P.subgrid) return;
  // the skipped segments that are sub-grid candidates (6 q > 1, flagged by the geometry): usually a handful
  __shared__ int s_seg[256];
  __shared__ int s_nseg;
  if (threadIdx.x == 0) s_nseg = 0;
  __syncthreads();
  for (int n = 1 + threadIdx.x; n < first; n += blockDim.x) {
    const int cx = __ldg(reinterpret_cast<const int *>(P.nodes.rec + n0 + n) + 12);  // cells.x: flags in the top bits
    if (((uint32_t)cx >> kCellFlagShift) & kFlagSub) {
      const int k = atomicAdd(&s_nseg, 1);
      if (k < 256) s_seg[k] = n;
    }
  }
  __syncthreads();
  const int nseg = s_nseg;
  if (nseg == 0) return;
  for (int ch = threadIdx.x; ch < P.nfr; ch += blockDim.x) {
    const double vel = __ldg(&P.velz[ch]);
    unsigned cnt = 0;
    auto count_seg = [&](int n) {
      const Node p = load_node(P.nodes.rec, n0 + n - 1), c = load_node(P.nodes.rec, n0 + n);
      const double ds = c.ds, lwav = 0.5 * (p.lw + c.lw);
      const double q = fabs((c.dvmu - p.dvmu) / (lwav / 2.99792458e5));
      const double s_c = ds * (vel - p.dvmu) / (c.dvmu - p.dvmu);
      const double dls3 = 3.0 * (ds / q);
      const double sright = s_c + dls3, sleft = s_c - dls3;
      if (sright > 0.0 && sleft < ds) {
        const double lg_ds = (sright - sleft) / (kLgNrMax - 1.0);
        for (int j = 1; j <= kLgNrMax; j++) {
          const double sj = sleft + (j - 1) * lg_ds;
          if (sj > 0.0 && sj < ds) cnt++;
        }
      }
    };
    if (nseg <= 256) {
      for (int k = 0; k < nseg; k++) count_seg(s_seg[k]);
    } else {  // (list overflow: walk all skipped segments)
      for (int n = 1; n < first; n++)
        if (load_node(P.nodes.rec, n0 + n).flags & kFlagSub) count_seg(n);
    }
    s_x[ch] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // exclusive prefix over channels 1..nfr-1 (s_x[0] stays the count of channel 0)
    unsigned run = 0;
    for (int ch = 1; ch < P.nfr; ch++) {
      const unsigned v = s_x[ch];
      s_x[ch] = run;
      run += v;
    }
    s_x[P.nfr] = run;
  }
  __syncthreads();
  unsigned long long e = 0;
  for (int l = threadIdx.x; l < P.nl; l += blockDim.x) {
    const long long task = (long long)ray * P.nl + l;
    if (!P.nitems[task]) continue;
    const int4 rg = P.rng[task];
    e += s_x[0];
    if (rg.y >= rg.x) e += s_x[rg.y + 1] - s_x[rg.x];
    if (rg.z >= 1) e += s_x[rg.z + 1] - s_x[rg.z];
  }
  for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0 && e) atomicAdd(&P.counters[1], e);
}

// Work estimate of a plan per camera ring (rl_plan_costs: cost-weighted ring blocks for sharded renders), in
// units of warp instructions as the profiles show them: a ztile_kernel node step costs about 120 + 40 per
// channel slot of the thread, a zcont_kernel one about 100; building and scanning a ray's nodes (geometry,
// channel selection, wall) about 170 per node.
__global__ void __launch_bounds__(256) plan_cost_kernel(RenderParams P, unsigned n_main, unsigned n_all, double *cost) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_all) {
    int ray;
    double per;
    if (P.use_z) {
      const ZTile t = P.ztiles[i];
      const int GW = 32 >> t.lwshift, cw = (t.nchk + GW - 1) / GW, cw3 = 3 * ((cw + 2) / 3);
      ray = t.ray;
      per = i < n_main ? 120.0 + 40.0 * cw3 : 100.0;
    } else {  // tile_kernel: four consumer warps and a staging warp per 128 items
      const TileDesc t = P.tiles[i];
      ray = t.ray;
      per = 100.0 + 3.5 * (double)(t.g1 - t.g0);
    }
    const int N = (int)(P.node_off[ray + 1] - P.node_off[ray]);
    const int ns = P.nstart ? max(1, min(P.nstart[ray], N - 1)) : 1;
    const int ring = ray == 0 ? 0 : 1 + (ray - 1) / P.nphi;
    atomicAdd(&cost[ring], per * (double)max(0, N - ns));
  }
  if (i < (unsigned)P.nray) {
    const int N = (int)(P.node_off[i + 1] - P.node_off[i]);
    const int ring = i == 0 ? 0 : 1 + ((int)i - 1) / P.nphi;
    atomicAdd(&cost[ring], 170.0 * (double)N + (i == 0 ? 400.0 * (double)N * P.nl * P.nfr / 32.0 : 0.0));
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
    atomicAdd(&P.counters[3], e);
  }
}

// ------------------------------------------------------------------------------------------
// Rectangular imager / position-velocity cube (telescope.F:2061-2227 make_freq_image_rectang, driven per line
// by calc_write_line_posvel :1828): every pixel of an nx x ny Cartesian camera at every channel, no
// NONREDUNDANT shortcut, with the line-of-sight optical depth char_tau next to the intensity.  The reference
// marks this mode obsolete (line_params.ini: "image=1 -- do not use"); it is served by the reference-ordered
// scalar walk (integrate_ray_channel), one thread per (line, pixel, channel).  Ray 0 is the central ray that
// carries the unresolved star, rays 1 .. nx ny the pixels (ix outer, iy inner: telescope.F:2124-2125).
// image, tau: [nl][nx][ny][nfr] = imrec_int(inu,ix,iy), imrec_tau(inu,ix,iy).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rect_kernel(RenderParams P, int npix, double *image, double *tau_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)P.nl * npix * P.nfr;
  const int lane = threadIdx.x & 31;
  unsigned long long e = 0, sg = 0, r = 0;
  if (i < n) {
    const int ch = (int)(i % P.nfr);
    const int pix = (int)((i / P.nfr) % npix);
    const int l = (int)(i / ((long long)P.nfr * npix));
    const int ray = 1 + pix;
    const long long nn = P.node_off[ray + 1] - P.node_off[ray];
    double inten, tau = 0.0;
    if (nn > 0) {
      unsigned ne;
      int maser = 0;
      inten = integrate_ray_channel(P, l, ray, ch, tau, ne, maser);
      if (maser) atomicOr(&P.maser[l], 1);
      e = ne;
      sg = (unsigned long long)(nn - 1);
      r = 1;
    } else {  // the pixel misses the model (rp_b >= 0.999 R_nr, telescope.F:2127-2148)
      inten = (P.out_itype == 3) ? P.isrf_line[(size_t)l * P.nfr + ch] : 0.0;
    }
    image[i] = inten;
    if (tau_out) tau_out[i] = tau;
  }
  for (int o = 16; o; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  if (lane == 0 && r) {
    atomicAdd(&P.counters[0], r);
    atomicAdd(&P.counters[1], e);
    atomicAdd(&P.counters[2], sg);
    atomicAdd(&P.counters[3], e);
  }
}
// the unresolved star smeared over the four central pixels (telescope.F:2153-2200)
__global__ void __launch_bounds__(128) rect_star_kernel(RenderParams P, int nx, int ny, double srat, double *image) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nl * P.nfr) return;
  const int l = i / P.nfr, ch = i % P.nfr;
  double tau;
  unsigned ne;
  int maser = 0;
  double dummy = integrate_ray_channel(P, l, 0, ch, tau, ne, maser);
  if (maser) atomicOr(&P.maser[l], 1);
  dummy = dummy * srat;
  const double srat1 = 1.0 - srat;
  double *im = image + (size_t)l * nx * ny * P.nfr + ch;
  for (int dx = 0; dx < 2; dx++)
    for (int dy = 0; dy < 2; dy++) {
      double *p = im + ((size_t)(nx / 2 - 1 + dx) * ny + (size_t)(ny / 2 - 1 + dy)) * P.nfr;
      *p = dummy + srat1 * *p;
    }
  atomicAdd(&P.counters[0], 1ull);
  atomicAdd(&P.counters[1], (unsigned long long)ne);
  atomicAdd(&P.counters[2], (unsigned long long)(P.node_off[1] - P.node_off[0] - 1));
  atomicAdd(&P.counters[3], (unsigned long long)ne);
}

// tiles of a ray: its item list (lines in index order) is cut greedily into runs of at most
// tile_threads items spanning at most tile_max_lines lines; the item budget is evened out over the
// ray first.  FILL = false counts the tiles of every ray, FILL = true (after the scan) writes them.
template <bool FILL>
__global__ void plan_kernel(RenderParams P) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray > P.nray) return;
  unsigned n = 0;
  if (ray < P.nray) {
    const unsigned *off = P.item_off + (size_t)ray * P.nl;
    const unsigned base = off[0], M = off[P.nl] - base;
    if (M) {
      const unsigned T0 = (unsigned)P.tile_threads;
      const unsigned T = (M + (M + T0 - 1) / T0 - 1) / ((M + T0 - 1) / T0);  // ceil(M / ceil(M / T0))
      const unsigned Lmax = (unsigned)P.tile_max_lines;
      unsigned items = 0, lines = 0, g0 = 0;
      int l0 = 0;
      TileDesc *out = FILL ? P.tiles + P.cta_off[ray] : nullptr;
      for (int l = 0; l < P.nl; l++) {
        unsigned left = off[l + 1] - off[l];
        while (left) {
          if (items == T || lines == Lmax) {
            if (FILL) {
              TileDesc d;
              d.ray = ray; d.g0 = g0; d.g1 = g0 + items; d.l0 = (unsigned short)l0; d.nlc = (unsigned short)lines;
              out[n] = d;
            }
            n++;
            g0 += items;
            items = 0;
            lines = 0;
          }
          if (!lines) l0 = l;
          const unsigned take = min(left, T - items);
          items += take;
          lines++;
          left -= take;
        }
      }
      if (items) {
        if (FILL) {
          TileDesc d;
          d.ray = ray; d.g0 = g0; d.g1 = g0 + items; d.l0 = (unsigned short)l0; d.nlc = (unsigned short)lines;
          out[n] = d;
        }
        n++;
      }
    }
  }
  if (!FILL) P.ncta[ray] = n;
}

// the rare case of telescope.F:583: the row's continuum is exactly zero, so the reference integrates
// every skipped channel until one is non-zero.  Sequential; returns the counters of the extra work.
__device__ __noinline__ void fill_zero_continuum(const RenderParams &P, long long task, int l, int ray,
                                                 const int4 rg, double *I, double cont) {
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
    atomicAdd(&P.counters[3], e);
  }
}

// continuum copy for the channels the reference skips (telescope.F:557-612); one warp per task
// (cube / mask mode: every channel of the image is materialised)
__global__ void __launch_bounds__(256) fill_kernel(const __grid_constant__ RenderParams P) {
  const int lane = threadIdx.x & 31;
  const long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long ntask = (long long)P.nl * P.nray;
  if (task >= ntask) return;
  const int ray = (int)(task / P.nl), l = (int)(task % P.nl);
  if (ray == 0) return;
  const int4 rg = P.rng[task];
  if (rg.w == 3) return;  // ring of another rank
  const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
  double *I = P.img + row * P.nfr;
  // continuum known after channel 0 (if out of range) or after the pre-integrated channel c0
  const double cont = (rg.w == 0) ? I[0] : (rg.z >= 0 ? I[rg.z] : 0.0);
  if (cont != 0.0) {
    for (int c = 1 + lane; c < P.nfr; c += 32) {
      const bool in = (c >= rg.x && c <= rg.y);
      if (!in && c != rg.z) I[c] = cont;
    }
  } else if (lane == 0) {
    fill_zero_continuum(P, task, l, ray, rg, I, cont);
  }
}

// spectrum mode (no cube, no mask requested): the continuum copies are not materialised -- the ring
// sum synthesises the skipped channels from the row's continuum channel.  Only rows on which
// tile_kernel saw an exactly zero intensity (dense[task] == 2) are examined here; if their continuum
// is zero they are completed like in cube mode and stay flagged dense (1) for the ring sum.
__global__ void __launch_bounds__(256) fill_sparse_kernel(const __grid_constant__ RenderParams P) {
  const long long task = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ntask = (long long)P.nl * P.nray;
  if (task >= ntask || P.dense[task] == 0) return;
  const int ray = (int)(task / P.nl), l = (int)(task % P.nl);
  const int4 rg = P.rng[task];
  const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
  double *I = P.img + row * P.nfr;
  const double cont = (rg.w == 0) ? I[0] : (rg.z >= 0 ? I[rg.z] : 0.0);
  if (cont != 0.0) {
    P.dense[task] = 0;
    return;
  }
  fill_zero_continuum(P, task, l, ray, rg, I, cont);
  P.dense[task] = 1;
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

// telescope.F:1388-1423: per-ring flux contribution.  Row 0 = the central beam (pi ri(1)^2 I_centre,
// telescope.F:1393-1396); rows 1..nrr = mean over phi of one ring times the ring area, phi in index
// order.  Rings outside [ring_lo, ring_hi] (another rank's block) give 0.
__global__ void __launch_bounds__(128) ringsum_kernel(RenderParams P, const double *surf, double *ring) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)P.nl * (P.nrr + 1) * P.nfr;
  if (i >= n) return;
  const int c = (int)(i % P.nfr);
  const int ir = (int)((i / P.nfr) % (P.nrr + 1));
  const int l = (int)(i / ((long long)P.nfr * (P.nrr + 1)));
  if (ir < P.ring_lo || ir > P.ring_hi) {
    ring[i] = 0.0;
    return;
  }
  const double *I = P.img + ((size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)ir * P.nphi) * P.nfr + c;
  if (ir == 0) {
    ring[i] = surf[0] * I[0];
    return;
  }
  double dslum = 0.0;
  if (!P.sparse || !P.nonredundant) {
    for (int ip = 0; ip < P.nphi; ip++) dslum = dslum + I[(size_t)ip * P.nfr];
  } else {
    // skipped channels carry the row's continuum (telescope.F:582-612), never written in this mode
    // (the additions stay in phi order; the loads of kU pixels -- channel range, then the intensity it selects --
    // are in flight together: the loop is bound by their latency, not by their number)
    constexpr int kU = 6;
    const long long task0 = ((long long)(1 + (ir - 1) * P.nphi)) * P.nl + l;
    for (int ip0 = 0; ip0 < P.nphi; ip0 += kU) {
      int4 rg[kU];
      unsigned char dn[kU];
      double v[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int ip = min(ip0 + u, P.nphi - 1);
        const long long task = task0 + (long long)ip * P.nl;
        rg[u] = __ldg(&P.rng[task]);
        dn[u] = P.dense[task];
      }
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int ip = min(ip0 + u, P.nphi - 1);
        const double *row = I - c + (size_t)ip * P.nfr;
        const bool own = (c == 0) || (c >= rg[u].x && c <= rg[u].y) || (c == rg[u].z) || dn[u];
        const int src = own ? c : ((rg[u].w == 0) ? 0 : rg[u].z);
        v[u] = row[src];
      }
#pragma unroll
      for (int u = 0; u < kU; u++)
        if (ip0 + u < P.nphi) dslum = dslum + v[u];
    }
  }
  dslum = dslum / (1.0 * P.nphi);
  dslum = dslum * surf[ir];
  ring[i] = dslum;
}

// telescope.F:1388-1433: sum of the contributions in index order (centre, rings 1..nrr), divided by
// distance^2.  ring: [nl][nrr+1][nfr]
__global__ void __launch_bounds__(128) flux_kernel(int nl, int nrr, int nfr, const double *ring, double dist2,
                                                   double *flux) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nl * nfr) return;
  const int c = i % nfr, l = i / nfr;
  double slum = 0.0;
  const double *rg = ring + (size_t)l * (nrr + 1) * nfr + c;
  for (int ir = 0; ir <= nrr; ir++) slum = slum + rg[(size_t)ir * nfr];
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
  if (P.nonredundant) mask_kernel<<<(unsigned)((P.ncell * 4 + 255) / 256), 256, 0, st>>>(P);
  span_kernel<<<(unsigned)P.nray, kSpanThreads, 0, st>>>(P);
}
void launch_wall(const RenderParams &P, double lw_min, double *admin, double *smin, unsigned long long *wstat,
                 int *nstart, cudaStream_t st) {
  wallprep_kernel<<<(unsigned)((P.ncell + 255) / 256), 256, 0, st>>>(P, lw_min, admin, smin, wstat);
  wall_kernel<<<(P.nray + 3) / 4, 128, 0, st>>>(P, admin, smin, wstat, nstart);
  if (P.subgrid)
    wallcount_kernel<<<P.nray, 128, (P.nfr + 1) * sizeof(unsigned), st>>>(P, nstart, P.use_z ? 0 : 1);
}
void launch_plan(const RenderParams &P, bool fill, cudaStream_t st) {
  if (P.use_z) {
    if (fill) zplan_kernel<true><<<(P.nray + 1 + 127) / 128, 128, 0, st>>>(P);
    else zplan_kernel<false><<<(P.nray + 1 + 127) / 128, 128, 0, st>>>(P);
    return;
  }
  if (fill) plan_kernel<true><<<(P.nray + 1 + 127) / 128, 128, 0, st>>>(P);
  else plan_kernel<false><<<(P.nray + 1 + 127) / 128, 128, 0, st>>>(P);
}
// dynamic shared memory per tile_kernel block: 4 blocks of 128+32 threads (8 of 64+32) per SM, next
// to the 12 KB of exp table, node / line scratch and item metadata each block carries
int tile_smem_limit(int threads) {
  static int done = 0;
  const int want128 = 44 * 1024, want64 = 15 * 1024;
  if (!done) {
    cudaFuncSetAttribute(tile_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, want128);
    cudaFuncSetAttribute(tile_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, want64);
    done = 1;
  }
  return threads == 128 ? want128 : want64;
}
// most lines a tile may span so that kTileBufs buffers of three slots (+ the phantom) fit the budget
int tile_max_lines(int threads) {
  const int per_buf = tile_smem_limit(threads) / kTileBufs;
  // 3 slots of (kPairBytes nlc + kSlotBytes) + phantom (sizeof(HotLine) nlc + sizeof(HotNode))
  return min(kMaxTileLines, (per_buf - 3 * kSlotBytes - (int)sizeof(HotNode)) / (3 * kPairBytes + (int)sizeof(HotLine)));
}
void launch_plan_cost(const RenderParams &P, unsigned n_main, unsigned n_all, double *ring_cost, cudaStream_t st) {
  const unsigned n = max(n_all, (unsigned)P.nray);
  plan_cost_kernel<<<(n + 255) / 256, 256, 0, st>>>(P, n_main, n_all, ring_cost);
}
void launch_rect(const RenderParams &P, int nx, int ny, double *image, double *tau, double srat, bool star,
                 cudaStream_t st) {
  const long long n = (long long)P.nl * nx * ny * P.nfr;
  rect_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, nx * ny, image, tau);
  if (star) rect_star_kernel<<<(P.nl * P.nfr + 127) / 128, 128, 0, st>>>(P, nx, ny, srat, image);
}
void launch_center(const RenderParams &P, cudaStream_t st) {
  if (P.ring_lo <= 0) center_kernel<<<(P.nl * P.nfr + 127) / 128, 128, 0, st>>>(P);
}
// the continuum-only tiles [tile0, tile0 + ntile) of the ztile plan
void launch_zcont(const RenderParams &P, unsigned tile0, unsigned ntile, cudaStream_t st) {
  if (ntile) zcont_kernel<<<(ntile + kZcWarps - 1) / kZcWarps, 32 * kZcWarps, 0, st>>>(P, tile0, ntile);
}
void launch_integrate(const RenderParams &P, unsigned total_ctas, cudaStream_t st) {
  if (!total_ctas) return;
  if (P.use_z) {
    if (P.zlw == 1) chan_kernel<kCCw><<<total_ctas, 32, 0, st>>>(P);
    else ztile_kernel<kZCw><<<(total_ctas + kZWarps - 1) / kZWarps, 32 * kZWarps, 0, st>>>(P);
    return;
  }
  if (P.tile_threads == 128) tile_kernel<128><<<total_ctas, 128 + kProducerThreads, P.smem_budget, st>>>(P);
  else tile_kernel<64><<<total_ctas, 64 + kProducerThreads, P.smem_budget, st>>>(P);
}
void launch_fill(const RenderParams &P, cudaStream_t st) {
  if (!P.nonredundant) return;  // every channel was integrated
  const long long ntask = (long long)P.nl * P.nray;
  if (P.sparse) {
    fill_sparse_kernel<<<(unsigned)((ntask + 255) / 256), 256, 0, st>>>(P);
  } else {
    const long long threads = ntask * 32;
    fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P);
  }
}
void launch_center_replicate(const RenderParams &P, cudaStream_t st) {
  const long long n = (long long)P.nl * (P.nphi - 1) * P.nfr;
  if (n <= 0 || P.ring_lo > 0) return;
  center_replicate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P);
}
void launch_ringsum(const RenderParams &P, const double *surf, double *ring, cudaStream_t st) {
  const long long n = (long long)P.nl * (P.nrr + 1) * P.nfr;
  ringsum_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, surf, ring);
}
void launch_flux(int nl, int nrr, int nfr, const double *ring, double dist2, double *flux, cudaStream_t st) {
  flux_kernel<<<(nl * nfr + 127) / 128, 128, 0, st>>>(nl, nrr, nfr, ring, dist2, flux);
}
void launch_dfma_peak(double *sink, int iters, int blocks, int threads, cudaStream_t st) {
  dfma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters);
}
void launch_cmask(const RenderParams &P, unsigned char *accum, int *out, cudaStream_t st) {
  const long long per = (long long)(P.nrr + 1) * P.nphi * P.nfr;
  cmask_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(P, accum, out);
}

}  // namespace rl

// rl_render.cu -- per-line preparation, ray integration and flux reduction kernels.
//
//   prep_cells_kernel   line.F:3502-3608 (dust source term at line centre), setup.F:937 (bplanck),
//                       line.F:4069-4088 (N_up, N_down per cell)
//   span_kernel         telescope.F:4265-4270 (minvel/maxvel of a ray) + the NONREDUNDANT channel
//                       selection of telescope.F:544-612 -> work items
//   integrate_kernel    telescope.F:3889-4312 (charintline), line.F:4636-4848
//                       (clever_integrate_element_linedust), line.F:4515-4624
//                       (integrate_element_linedust), line.F:2280 (voigt_profile),
//                       transfer.F:1498 (qdr_src_2): one thread per (line, ray, channel) item
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
  const int l = (int)(task / P.nray), ray = (int)(task % P.nray);
  if (ray == 0 || !P.nonredundant) {
    if (lane == 0) {
      P.rng[task] = make_int4(1, P.nfr - 1, -1, ray == 0 ? 2 : 1);
      P.nitems[task] = P.nfr;
    }
    return;
  }
  const long long n0 = P.node_off[ray], n1 = P.node_off[ray + 1];
  const double4 *cellL = P.cellL + (size_t)l * P.ncell;
  double vmin = 2.0, vmax = -2.0;
  // telescope.F:4265-4270: start node of every segment, i.e. all nodes but the last
  for (long long i = n0 + lane; i < n1 - 1; i += 32) {
    const uint32_t fl = P.nodes.flag[i];
    const double4 v = gather_line(cellL, P.nodes.cell[i], P.nodes.wr[i], P.nodes.wt[i], fl & kFlagIcrMask);
    if (v.z + v.w > P.levthres) {
      const double dv = P.nodes.dvmu[i];
      vmin = fmin(vmin, dv);
      vmax = fmax(vmax, dv);
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

// ------------------------------------------------------------------------------------------
// the formal solution of one ray at one channel
// ------------------------------------------------------------------------------------------
struct Carry {
  double phiprof0, srcl0, alpl0;
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
__device__ __forceinline__ void integrate_element(const LineDev &L, double dnu_ch, double &inten,
                                                  double ds, double srcd0, double srcd1,
                                                  double alpd0, double alpd1, double lw0, double lw1,
                                                  double dvmu0, double dvmu1, double nup0,
                                                  double nup1, double ndown0, double ndown1,
                                                  Carry &k, double &tau) {
  const double lwav = 0.5 * (lw0 + lw1);
  const double aa = 3.33567e-6 * L.nu0 * lwav;
  const double norm = 0.56419583546 / aa;
  if (k.init) {
    const double dnu0 = dnu_ch - L.nu0 * dvmu0;
    const double u0 = dnu0 / aa;
    k.phiprof0 = norm * exp(-(u0 * u0));
    k.srcl0 = 5.27296241956e-28 * L.nu0 * nup0 * L.aud * k.phiprof0;
    k.alpl0 = 5.27296241956e-28 * L.nu0 * k.phiprof0 * (ndown0 * L.bdu - nup0 * L.bud);
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
  k.phiprof0 = phiprof1;
  k.srcl0 = srcl1;
  k.alpl0 = alpl1;
  k.init = 0;
}

// telescope.F:3889-4312 for ray `ray`, channel `ch` (0-based) of line slot `l`.
// Returns the intensity; counts element integrations in nelem.
__device__ double integrate_ray_channel(const RenderParams &P, int l, int ray, int ch, double &tau,
                                        unsigned &nelem, int &maser) {
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
  k.phiprof0 = k.srcl0 = k.alpl0 = 0.0;
  uint32_t fl = P.nodes.flag[n0];
  double4 v0 = gather_line(cellL, P.nodes.cell[n0], P.nodes.wr[n0], P.nodes.wt[n0], fl & kFlagIcrMask);
  double dvmu0 = P.nodes.dvmu[n0], lw0 = P.nodes.lw[n0];
  for (long long i = n0 + 1; i < n1; i++) {
    fl = P.nodes.flag[i];
    const double ds = P.nodes.ds[i];
    const double dvmu1 = P.nodes.dvmu[i], lw1 = P.nodes.lw[i];
    const double4 v1 = gather_line(cellL, P.nodes.cell[i], P.nodes.wr[i], P.nodes.wt[i], fl & kFlagIcrMask);
    if (fl & (kFlagInit | kFlagStar | kFlagZero)) {
      if (fl & kFlagZero) inten = 0.0;
      if (fl & kFlagStar)
        inten = (1.0 - P.starfract) * inten + P.starfract * P.star_line[(size_t)l * P.nfr + ch];
      k.init = 1;
    }
    // line.F:4636-4848
    bool done = false;
    if (P.subgrid) {
      const double lw = 0.5 * (lw0 + lw1);
      const double ds_over_deltal_s = fabs((dvmu1 - dvmu0) / (lw / 2.99792458e5));
      if (2.0 * 3.0 * ds_over_deltal_s > 1.0) {
        const double s_c = ds * (velo_ch - dvmu0) / (dvmu1 - dvmu0);
        const double dls = ds / ds_over_deltal_s;
        const double sright = s_c + 3.0 * dls;
        const double sleft = s_c - 3.0 * dls;
        if (sright > 0.0 && sleft < ds) {
          const double lg_ds = (sright - sleft) / (kLgNrMax - 1.0);
          double sp = 0.0, nup_p = v0.z, ndn_p = v0.w, dv_p = dvmu0, sd_p = v0.x, ad_p = v0.y;
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
            integrate_element(L, dnu_ch, inten, s - sp, sd_p, sd_c, ad_p, ad_c, lw, lw, dv_p, dv_c,
                              nup_p, nup_c, ndn_p, ndn_c, k, tau);
            nelem++;
            sp = s; nup_p = nup_c; ndn_p = ndn_c; dv_p = dv_c; sd_p = sd_c; ad_p = ad_c;
          }
          done = true;
        }
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

// one thread per (line, ray, channel) item
__global__ void __launch_bounds__(128) integrate_kernel(RenderParams P, unsigned total_items) {
  const unsigned item = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const long long ntask = (long long)P.nl * P.nray;
  const bool active = item < total_items;
  // the warp's first item -> task by binary search; later lanes walk forward
  unsigned first = __shfl_sync(0xffffffffu, item, 0);
  long long task = 0;
  if (lane == 0 && first < total_items) {
    long long lo = 0, hi = ntask;  // largest t with item_off[t] <= first
    while (hi - lo > 1) {
      long long mid = (lo + hi) >> 1;
      if (P.item_off[mid] <= first) lo = mid;
      else hi = mid;
    }
    task = lo;
  }
  task = __shfl_sync(0xffffffffu, task, 0);
  unsigned nelem = 0, nseg = 0;
  int maser = 0;
  if (active) {
    while (P.item_off[task + 1] <= item) task++;
    const int l = (int)(task / P.nray), ray = (int)(task % P.nray);
    const unsigned j = item - P.item_off[task];
    const int4 rg = P.rng[task];
    const int nin = (rg.y >= rg.x) ? (rg.y - rg.x + 1) : 0;
    int ch;
    bool masked;  // reference sets imcir_cmask=1 for this channel
    if (rg.w == 2) { ch = (int)j; masked = false; }  // centre ray: the reference never sets its mask
    else if (!P.nonredundant) { ch = (int)j; masked = (j == 0); }  // telescope.F:548 only
    else if (j == 0) { ch = 0; masked = true; }
    else if ((int)j <= nin) { ch = rg.x + (int)j - 1; masked = true; }
    else { ch = rg.z; masked = false; }
    double tau;
    const double inten = integrate_ray_channel(P, l, ray, ch, tau, nelem, maser);
    const size_t row = (size_t)l * (size_t)(P.nrr + 1) * P.nphi + (size_t)img_row(P, ray);
    P.img[row * P.nfr + ch] = inten;
    if (P.integ) P.integ[row * P.nfr + ch] = masked ? 1 : 2;
    if (ray == 0 && ch == P.nfr - 1) P.tau_center[l] = tau;
    if (maser) atomicOr(&P.maser[l], 1);
    nseg = (unsigned)(P.node_off[ray + 1] - P.node_off[ray]);
    nseg = nseg > 0 ? nseg - 1 : 0;
  }
  // work counters
  unsigned long long e = nelem, s = nseg, r = active ? 1 : 0;
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

// continuum copy for the channels the reference skips (telescope.F:557-612); one warp per task
__global__ void __launch_bounds__(256) fill_kernel(RenderParams P) {
  const int lane = threadIdx.x & 31;
  const long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long ntask = (long long)P.nl * P.nray;
  if (task >= ntask) return;
  const int l = (int)(task / P.nray), ray = (int)(task % P.nray);
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
void launch_integrate(const RenderParams &P, unsigned total_items, cudaStream_t st) {
  if (!total_items) return;
  integrate_kernel<<<(total_items + 127) / 128, 128, 0, st>>>(P, total_items);
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

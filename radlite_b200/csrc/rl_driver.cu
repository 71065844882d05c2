// rl_driver.cu -- the driver-side ends of the path on the device, so that a many-line
// spectrum never passes through ASCII intermediates:
//   rl_set_lines_lte         LTE level populations from level energies, degeneracies, the gas temperature and
//                            the tabulated partition sum, written straight into the population table the
//                            ray tracer reads (pyradlite radlite.py:1111-1119, 1147-1153; PRO/make_levelpop.pro)
//                            -- replaces levelpop_<mol>.dat (4 GB at BASELINE configs[4]) and its upload
//   rl_synthesize_spectrum   per-line spectra -> one spectrum on a wavelength grid: continuum subtraction,
//                            interpolation of every line onto the common grid, continuum interpolation,
//                            Gaussian convolution, resampling (radlite.py:3001-3184 _process_spectrum;
//                            PRO/genspec.pro)
// Compiled with --fmad=false: the arithmetic follows the numpy expressions operation by operation (the
// checker is oracle/driver_np.py); what differs is libm's exp against CUDA's (<= 1 ulp).
#include "../../include/radlite_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

extern "C" int rl_internal_set_lines_device(rl_ctx *ctx, int nlines, int nlevels, const int *lev_up,
                                            const int *lev_down, const double *linefreq, const double *aud,
                                            const double *gdeg, const double *d_popul_src);
extern "C" int rl_internal_grid_cells(const rl_ctx *ctx);
extern "C" void rl_internal_count_launches(rl_ctx *ctx, int n);
extern "C" int rl_internal_fail(rl_ctx *ctx, int code, const char *msg);

namespace {

// radlite.py:27-41 (the astropy-free constant set)
constexpr double c0 = 2.99792458E10, h0 = 6.6262000E-27, kB0 = 1.3807E-16;
constexpr double cinmu0 = c0 * 1.0E4, cinkm0 = c0 / 1.0E5;

// scipy interp1d(kind='linear'): idx = searchsorted(x, xnew) clipped to [1, n-1]; slope (xnew - x_lo) + y_lo
__host__ __device__ inline int search_left(const double *x, int n, double v) {  // first i with x[i] >= v
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (x[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
__host__ __device__ inline double interp_lin(const double *x, const double *y, int n, double v) {
  int idx = search_left(x, n, v);
  idx = idx < 1 ? 1 : (idx > n - 1 ? n - 1 : idx);
  const double slope = (y[idx] - y[idx - 1]) / (x[idx] - x[idx - 1]);
  return slope * (v - x[idx - 1]) + y[idx - 1];
}

// one thread per (cell, level): popul[cell][lev] = g exp(-E_K / T) / Q(T), flushed to 0 below 1e-99
__global__ void __launch_bounds__(256) lte_kernel(long long ncell, int nlev, const double *__restrict__ ek,
                                                  const double *__restrict__ g, const double *__restrict__ tgas,
                                                  int npsum, const double *__restrict__ pt,
                                                  const double *__restrict__ ps, double *__restrict__ popul) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell * nlev) return;
  const long long cell = i / nlev;
  const int lev = (int)(i - cell * nlev);
  const double t = tgas[cell];
  const double q = interp_lin(pt, ps, npsum, t);
  double v = (g[lev] * exp(-1.0 * ek[lev] / t)) / 1.0 / q;
  if (v < 1E-99) v = 0.0;
  popul[i] = v;
}

// full-resolution emission spectrum: one thread per grid point, lines added in index order
// (radlite.py:3091-3110).  mu_old [nl][nfr+2], em_old [nl][nfr+2]
__global__ void __launch_bounds__(128) fullem_kernel(int nfull, const double *__restrict__ fullmu, int nl, int np,
                                                     const double *__restrict__ mu_old,
                                                     const double *__restrict__ em_old, double scale,
                                                     double *__restrict__ fullem) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfull) return;
  const double mu = fullmu[i];
  double acc = 0.0;
  for (int a = 0; a < nl; a++) {
    const double *x = mu_old + (size_t)a * np;
    if (mu <= x[np - 1] && mu >= x[0]) acc = acc + interp_lin(x, em_old + (size_t)a * np, np, mu);
  }
  fullem[i] = acc * scale;
}
// continuum between the line centres (radlite.py:3115-3127, 3133)
__global__ void __launch_bounds__(128) fullcont_kernel(int nfull, const double *__restrict__ fullmu, int n,
                                                       const double *__restrict__ mus, const double *__restrict__ cs,
                                                       double scale, const double *__restrict__ fullem,
                                                       double *__restrict__ fullcont, double *__restrict__ fully) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nfull) return;
  const double c = interp_lin(mus, cs, n, fullmu[i]) * scale;
  fullcont[i] = c;
  fully[i] = c + fullem[i];
}
// scipy.ndimage.convolve(mode='reflect') / sum(kernel) (radlite.py:3148-3151)
__global__ void __launch_bounds__(128) convolve_kernel(int L, const double *__restrict__ a, int n,
                                                       const double *__restrict__ k, double ksum,
                                                       double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  double acc = 0.0;
  for (int j = 0; j < n; j++) {
    int src = i + (n / 2) - j;
    if (src < 0) src = -src - 1;
    if (src >= L) src = 2 * L - 1 - src;
    src = src < 0 ? 0 : (src >= L ? L - 1 : src);  // (kernels longer than the spectrum: clamp)
    acc = acc + k[j] * a[src];
  }
  out[i] = acc / 1.0 / ksum;
}
// resampling onto the output grid (radlite.py:3158-3165)
__global__ void __launch_bounds__(128) resample_kernel(int nout, const double *__restrict__ outmu, int nfull,
                                                       const double *__restrict__ fullmu,
                                                       const double *__restrict__ a, const double *__restrict__ b,
                                                       const double *__restrict__ c, double *__restrict__ oa,
                                                       double *__restrict__ ob, double *__restrict__ oc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nout) return;
  const double mu = outmu[i];
  oa[i] = interp_lin(fullmu, a, nfull, mu);
  ob[i] = interp_lin(fullmu, b, nfull, mu);
  oc[i] = interp_lin(fullmu, c, nfull, mu);
}

template <typename T>
struct Dev {
  T *p = nullptr;
  ~Dev() {
    if (p) cudaFree(p);
  }
  cudaError_t put(const T *h, size_t n) {
    cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return h ? cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
  }
};

struct Grids {
  double boxwidth, vres, mumin, mumax;
  std::vector<double> fullmu, outmu;
};
// radlite.py:3031-3067
Grids make_grids(int nl, int nfr, const double *vel, const double *freq, double obsres, double vsampling) {
  Grids G;
  double maxvspan = 0.0;
  for (int a = 0; a < nl; a++) maxvspan = std::max(maxvspan, std::fabs(vel[(size_t)a * nfr]));
  G.boxwidth = std::max(3 * maxvspan, 3 * obsres);
  G.vres = vel[1] - vel[0];
  double mn = 1e300, mx = -1e300;
  for (int a = 0; a < nl; a++) {
    if (freq[a] == 0) continue;
    const double mu = cinmu0 / 1.0 / freq[a];
    mn = std::min(mn, mu);
    mx = std::max(mx, mu);
  }
  G.mumin = mn - (G.boxwidth * 1.0E9 * mn / cinmu0);
  G.mumax = mx + (G.boxwidth * 1.0E9 * mx / cinmu0);
  auto grid = [&](double res) {
    const double growth = (1 + (res / 1.0 / cinkm0));
    const int len = (int)(std::floor(std::log(G.mumax / 1.0 / G.mumin) / std::log(growth)) + 1);
    std::vector<double> mu((size_t)std::max(len, 0));
    for (int a = 0; a < len; a++) mu[a] = G.mumin * std::pow(growth, (double)a);
    return mu;
  };
  G.fullmu = grid(G.vres);
  G.outmu = grid(vsampling);
  return G;
}

}  // namespace

extern "C" {

int rl_set_lines_lte(rl_ctx *ctx, int nlines, int nlevels, const int *lev_up, const int *lev_down,
                     const double *linefreq, const double *aud, const double *gdeg, const double *energy_cm,
                     const double *tgas, int npsum, const double *psum_temp, const double *psum) {
  if (!ctx) return 13;
  const long long ncell = rl_internal_grid_cells(ctx);
  if (ncell <= 0) return rl_internal_fail(ctx, 13, "set_lines_lte: call set_grid first");
  if (nlevels < 2 || npsum < 2 || !energy_cm || !tgas || !psum_temp || !psum)
    return rl_internal_fail(ctx, 13, "set_lines_lte: bad arguments");
  std::vector<double> ek(nlevels);
  for (int a = 0; a < nlevels; a++) ek[a] = energy_cm[a] * h0 * c0 / 1.0 / kB0;  // radlite.py:1111
  Dev<double> d_ek, d_g, d_t, d_pt, d_ps, d_pop;
  if (d_ek.put(ek.data(), nlevels) || d_g.put(gdeg, nlevels) || d_t.put(tgas, (size_t)ncell) ||
      d_pt.put(psum_temp, npsum) || d_ps.put(psum, npsum) || d_pop.put(nullptr, (size_t)ncell * nlevels))
    return rl_internal_fail(ctx, 13, "set_lines_lte: device allocation failed");
  const long long n = ncell * nlevels;
  lte_kernel<<<(unsigned)((n + 255) / 256), 256>>>(ncell, nlevels, d_ek.p, d_g.p, d_t.p, npsum, d_pt.p, d_ps.p, d_pop.p);
  rl_internal_count_launches(ctx, 1);
  if (cudaDeviceSynchronize() != cudaSuccess) return rl_internal_fail(ctx, 13, "set_lines_lte: kernel failed");
  return rl_internal_set_lines_device(ctx, nlines, nlevels, lev_up, lev_down, linefreq, aud, gdeg, d_pop.p);
}

int rl_synthesis_size(int nl, int nfr, const double *vel, const double *freq, double obsres, double vsampling,
                      int *nout, int *nfull) {
  if (nl < 1 || nfr < 2 || !vel || !freq) return 13;
  const Grids G = make_grids(nl, nfr, vel, freq, obsres, vsampling);
  if (nout) *nout = (int)G.outmu.size();
  if (nfull) *nfull = (int)G.fullmu.size();
  return 0;
}

int rl_synthesize_spectrum(rl_ctx *ctx, int nl, int nfr, const double *vel, const double *flux, const double *freq,
                           double dist_pc, double obsres, double vsampling, double *wavelength, double *spectrum,
                           double *emission, double *continuum) {
  if (!ctx) return 13;
  if (nl < 1 || nfr < 2 || !vel || !flux || !freq || !wavelength || !spectrum || !emission || !continuum)
    return rl_internal_fail(ctx, 13, "synthesize_spectrum: bad arguments");
  const Grids G = make_grids(nl, nfr, vel, freq, obsres, vsampling);
  const int nfull = (int)G.fullmu.size(), nout = (int)G.outmu.size(), np = nfr + 2;
  if (nfull < 2 || nout < 1) return rl_internal_fail(ctx, 13, "synthesize_spectrum: empty wavelength grid");
  // per line (radlite.py:3076-3098): continuum = straight line between the edge channels, emission = flux -
  // continuum, centre continuum at v = 0; old wavelength axis widened to the box
  std::vector<double> mu_old((size_t)nl * np), em_old((size_t)nl * np), contcen(nl), mus(nl);
  for (int a = 0; a < nl; a++) {
    const double *v = vel + (size_t)a * nfr, *f = flux + (size_t)a * nfr;
    const double xs[2] = {v[0], v[nfr - 1]}, ys[2] = {f[0], f[nfr - 1]};
    double *mo = mu_old.data() + (size_t)a * np, *eo = em_old.data() + (size_t)a * np;
    for (int k = 0; k < nfr; k++) eo[k + 1] = f[k] - interp_lin(xs, ys, 2, v[k]);
    eo[0] = eo[1];
    eo[np - 1] = eo[np - 2];
    contcen[a] = interp_lin(xs, ys, 2, 0.0);
    const double shift = cinmu0 / 1.0 / freq[a];
    mo[0] = ((-1 * G.boxwidth) * 1.0E9 / freq[a]) + shift;
    for (int k = 0; k < nfr; k++) mo[k + 1] = (v[k] * 1.0E9 / freq[a]) + shift;
    mo[np - 1] = (G.boxwidth * 1.0E9 / freq[a]) + shift;
    mus[a] = shift;
  }
  std::vector<int> order(nl);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int i, int j) { return mus[i] < mus[j]; });
  std::vector<double> musort(nl + 2), csort(nl + 2);
  for (int k = 0; k < nl; k++) {
    musort[k + 1] = mus[order[k]];
    csort[k + 1] = contcen[order[k]];
  }
  musort[0] = G.mumin;
  musort[nl + 1] = G.mumax;
  csort[0] = csort[1];
  csort[nl + 1] = csort[nl];
  // Gaussian kernel (radlite.py:3141-3144)
  const int ng = (int)std::ceil(3.0 * obsres / G.vres);
  std::vector<double> kern((size_t)std::max(ng, 1));
  double ksum = 0.0;
  {
    const double bot = ((obsres / 1.0 / G.vres) * (obsres / 1.0 / G.vres)) * std::log(2.0);
    for (int j = 0; j < ng; j++) {
      const double d = j - ((ng - 1) / 2.0);
      kern[j] = std::exp((-1 * 2 * (d * d)) / bot);
    }
    for (int j = 0; j < ng; j++) ksum += kern[j];  // np.sum: pairwise for long arrays, sequential below 8 terms
  }
  if (ng < 1) return rl_internal_fail(ctx, 13, "synthesize_spectrum: observing resolution below the sampling");
  const double scale = 1.0E23 / (dist_pc * dist_pc);
  Dev<double> d_fullmu, d_outmu, d_mo, d_eo, d_mus, d_cs, d_k, d_em, d_cont, d_y, d_rem, d_ry, d_oa, d_ob, d_oc;
  if (d_fullmu.put(G.fullmu.data(), nfull) || d_outmu.put(G.outmu.data(), nout) || d_mo.put(mu_old.data(), mu_old.size()) ||
      d_eo.put(em_old.data(), em_old.size()) || d_mus.put(musort.data(), nl + 2) || d_cs.put(csort.data(), nl + 2) ||
      d_k.put(kern.data(), ng) || d_em.put(nullptr, nfull) || d_cont.put(nullptr, nfull) || d_y.put(nullptr, nfull) ||
      d_rem.put(nullptr, nfull) || d_ry.put(nullptr, nfull) || d_oa.put(nullptr, nout) || d_ob.put(nullptr, nout) ||
      d_oc.put(nullptr, nout))
    return rl_internal_fail(ctx, 13, "synthesize_spectrum: device allocation failed");
  const unsigned gf = (unsigned)((nfull + 127) / 128), go = (unsigned)((nout + 127) / 128);
  fullem_kernel<<<gf, 128>>>(nfull, d_fullmu.p, nl, np, d_mo.p, d_eo.p, scale, d_em.p);
  fullcont_kernel<<<gf, 128>>>(nfull, d_fullmu.p, nl + 2, d_mus.p, d_cs.p, scale, d_em.p, d_cont.p, d_y.p);
  convolve_kernel<<<gf, 128>>>(nfull, d_em.p, ng, d_k.p, ksum, d_rem.p);
  convolve_kernel<<<gf, 128>>>(nfull, d_y.p, ng, d_k.p, ksum, d_ry.p);
  resample_kernel<<<go, 128>>>(nout, d_outmu.p, nfull, d_fullmu.p, d_rem.p, d_ry.p, d_cont.p, d_oa.p, d_ob.p, d_oc.p);
  rl_internal_count_launches(ctx, 5);
  if (cudaMemcpy(emission, d_oa.p, (size_t)nout * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(spectrum, d_ob.p, (size_t)nout * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(continuum, d_oc.p, (size_t)nout * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
    return rl_internal_fail(ctx, 13, "synthesize_spectrum: kernels failed");
  std::copy(G.outmu.begin(), G.outmu.end(), wavelength);
  return 0;
}

}  // extern "C"

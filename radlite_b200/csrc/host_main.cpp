// host_main.cpp -- radlite_b200_host: stand-alone replacement for the RADlite / RADlite_imcir executables
// the drivers spawn in a working directory (pyradlite radlite.py:593, PRO/line_run.pro:173).
//
// Reads the same input files in the same order as main.F:79-1043 (rl_io.cpp), hands the model to the
// CUDA library through the C ABI of include/radlite_b200.h -- the only way this program reaches the
// ray tracer -- and writes linespectrum_<mol>.dat (telescope.F:1662-1803), optionally
// lineposvelcirc_<mol>_<iline>.dat (SAVE_IMCIR build, telescope.F:1346-1371, 1595-1621) and
// radlite.success (main.F:68-70).  There is no CPU path: without a B200 rl_create fails and the program
// exits non-zero without radlite.success, which is how the drivers detect a failed run.
//
//   radlite_b200_host [--dir D] [--device N] [--imcir] [--parse-only] [--dump FILE] [--replay FILE]
// --imcir (or a program name containing "imcir", or RADLITE_B200_IMCIR=1) = the RADlite_imcir build.
// Developer options used by the CPU tests of the file formats (no GPU, nothing is computed):
//   --parse-only / --dump FILE   read and check the inputs, write the parsed model
//   --replay FILE                write the output files from spectra / cubes stored in FILE
//   --fmt e|f|i|g W D VALUE      print one number the way the Fortran edit descriptor would
#include "../../include/radlite_b200.h"
#include "rl_io.h"

#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// records of a --dump / --replay file: {name[32], dtype, ndim, dims, data}
struct Rec {
  std::vector<long long> dims;
  std::vector<double> d;
  std::vector<int> i;
};
static bool load_records(const std::string &file, std::vector<std::pair<std::string, Rec>> &out) {
  FILE *f = fopen(file.c_str(), "rb");
  if (!f) return false;
  char nm[32], dt;
  while (fread(nm, 1, 32, f) == 32) {
    Rec r;
    int nd = 0;
    if (fread(&dt, 1, 1, f) != 1 || fread(&nd, sizeof nd, 1, f) != 1) break;
    size_t n = 1;
    r.dims.resize(nd);
    for (int k = 0; k < nd; k++) {
      if (fread(&r.dims[k], sizeof(long long), 1, f) != 1) break;
      n *= (size_t)r.dims[k];
    }
    if (dt == 'd') {
      r.d.resize(n);
      if (fread(r.d.data(), sizeof(double), n, f) != n) break;
    } else {
      r.i.resize(n);
      if (fread(r.i.data(), sizeof(int), n, f) != n) break;
    }
    out.emplace_back(std::string(nm), std::move(r));
  }
  fclose(f);
  return true;
}
static const Rec *find_rec(const std::vector<std::pair<std::string, Rec>> &v, const char *name) {
  for (const auto &p : v)
    if (p.first == name) return &p.second;
  return nullptr;
}

static int fail_rl(rl_ctx *ctx, int code) {
  fprintf(stderr, " radlite_b200: %s\n", rl_last_error(ctx));
  return code > 0 ? (code & 255 ? code & 255 : 1) : 1;
}

#define RL(call)                                  \
  do {                                            \
    const int rc_ = (call);                       \
    if (rc_ != 0) return fail_rl(ctx, rc_);       \
  } while (0)

int main(int argc, char **argv) {
  std::string dir, dump, replay;
  int device = 0;
  bool imcir = std::strstr(argv[0], "imcir") != nullptr, parse_only = false;
  if (const char *e = getenv("RADLITE_B200_DEVICE")) device = atoi(e);
  if (const char *e = getenv("RADLITE_B200_IMCIR")) imcir = imcir || atoi(e) != 0;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "--dir" && i + 1 < argc) dir = argv[++i];
    else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
    else if (a == "--dump" && i + 1 < argc) dump = argv[++i];
    else if (a == "--replay" && i + 1 < argc) replay = argv[++i];
    else if (a == "--fmt" && i + 4 < argc) {
      const char k = argv[i + 1][0];
      const int fw = atoi(argv[i + 2]), fd = atoi(argv[i + 3]);
      const double v = atof(argv[i + 4]);
      std::string o = k == 'e' ? rlio::fmt_e(v, fw, fd) : k == 's' ? rlio::fmt_es(v, fw, fd) : k == 'f' ? rlio::fmt_f(v, fw, fd)
                    : k == 'i' ? rlio::fmt_i((long)v, fw) : rlio::fmt_list_real(v);
      printf("[%s]\n", o.c_str());
      return 0;
    }
    else if (a == "--imcir") imcir = true;
    else if (a == "--parse-only") parse_only = true;
    else {
      fprintf(stderr, "usage: %s [--dir D] [--device N] [--imcir] [--parse-only] [--dump FILE]\n", argv[0]);
      return 2;
    }
  }
  if (!dir.empty() && chdir(dir.c_str()) != 0) {
    fprintf(stderr, " cannot enter %s\n", dir.c_str());
    return 13;
  }
  printf(" ==============================================\n");
  printf("   RADLite line ray tracer -- radlite_b200_host \n");
  printf("   (B200 / sm_100a CUDA path behind the C ABI)  \n");
  printf(" ==============================================\n");
  unlink("radlite.success");
  rlio::WorkDir w;
  try {
    w = rlio::read_workdir();
    if (!dump.empty()) rlio::dump_workdir(w, dump);
  } catch (const rlio::Stop &s) {
    fprintf(stderr, " %s\n stop %d\n", s.msg.c_str(), s.code);
    return s.code & 255 ? s.code & 255 : 1;
  }
  const int nr = (int)w.r.size(), nth = (int)w.theta.size(), ncf = (int)w.cont_freq.size();
  printf(" grid %d x %d (x2 mirrored), %d continuum frequencies, %d dust species\n", nr, nth, ncf, w.nspec);
  printf(" molecule %s: %d levels, %d lines; rendering lines %d..%d, %d channels, passband %g km/s\n",
         w.molname.c_str(), w.nlevels, w.nlines, w.ilinestart, w.ilinestart + w.nlines_render - 1, w.nfr,
         w.passband);
  if (parse_only) return 0;
  if (!replay.empty()) {  // format test: write the outputs from stored results, no GPU involved
    std::vector<std::pair<std::string, Rec>> recs;
    if (!load_records(replay, recs)) return 13;
    const Rec *fl = find_rec(recs, "flux"), *ve = find_rec(recs, "velo"), *im = find_rec(recs, "image"),
              *cm = find_rec(recs, "cmask"), *rr = find_rec(recs, "rays_r"), *ri = find_rec(recs, "imcir_ri");
    if (!fl || !ve || fl->dims.size() != 2) return 13;
    const int nl = (int)fl->dims[0], nfr = (int)fl->dims[1];
    const std::string specfile = "linespectrum_" + w.molname + ".dat";
    try {
      rlio::write_spectrum_header(specfile, w.molname, "./" + w.molfile, nl, nfr, w.dist_cm, w.radvelo,
                                  w.incl_deg * 0.0174532925199, w.style);
      for (int l = 0; l < nl; l++) {
        const int il = w.ilinestart + l;
        if (im && cm && rr && ri) {
          const int nrr = (int)im->dims[1] - 1, nphi = (int)im->dims[2];
          const size_t per = (size_t)(nrr + 1) * nphi * nfr;
          rlio::write_imcir(rlio::imcir_filename(w.molname, il), nfr, w.linefreq[il - 1], nphi, nrr, ri->d.data(),
                            rr->d.data(), ve->d.data() + (size_t)l * nfr, im->d.data() + l * per,
                            cm->i.data() + l * per);
        }
        rlio::append_line_spectrum(specfile, w.lev_up[il - 1], w.lev_down[il - 1], w.linefreq[il - 1], nfr,
                                   ve->d.data() + (size_t)l * nfr, fl->d.data() + (size_t)l * nfr, w.radvelo);
      }
    } catch (const rlio::Stop &s) {
      fprintf(stderr, " %s\n", s.msg.c_str());
      return 13;
    }
    return 0;
  }

  const auto t0 = std::chrono::steady_clock::now();
  rl_ctx *ctx = nullptr;
  const int rc = rl_create(&ctx, device);
  if (rc != 0) {
    fprintf(stderr, " radlite_b200: rl_create(device=%d) failed with %d: no usable sm_100 GPU (no CPU fallback)\n",
            device, rc);
    return 1;
  }
  RL(rl_set_grid(ctx, nr, nth, w.r.data(), w.theta.data()));
  RL(rl_set_medium(ctx, w.rho.data(), w.abund.data(), w.vel.data(), w.linewidth.data(), w.umass_av));
  RL(rl_set_lines(ctx, w.nlines, w.nlevels, w.lev_up.data(), w.lev_down.data(), w.linefreq.data(), w.aud.data(),
                  w.gdeg.data(), w.popul.data()));
  RL(rl_set_dust(ctx, w.nspec, w.nsize.data(), ncf, w.cont_freq.data(), w.kappa_abs.data(), w.kappa_scat.data(),
                 w.dust_rho.data(), w.dust_temp.data(), w.scat.empty() ? nullptr : w.scat.data()));
  // main.F:792 setup_rays_circular(1,nr,1,main_anginf,main_nrphiinf,main_nrrayextra,telesc_dbdr,rstar,1,10)
  const double anginf = w.incl_deg * 0.0174532925199;  // telescope.F:188
  RL(rl_set_camera(ctx, anginf, w.nphi, w.nrext, w.dbdr, w.rstar, 1, 10));
  RL(rl_set_bc(ctx, w.in_itype, w.out_itype, ncf, w.cont_freq.data(), w.starspec.data(),
               w.isrf.empty() ? nullptr : w.isrf.data()));
  RL(rl_set_options(ctx, 1, 1, 1.e-3, -1.0));  // configure.h: SUBGRID, NONREDUNDANT, LEVTHRES

  if (w.command == 2) {
    // linespectrum.inp command 2 -> telesc_command 6: main.F:800 setup_rays_rectang, main.F:1062-1075 the line
    // loop over calc_write_line_posvel (telescope.F:1828)
    RL(rl_set_camera_rect(ctx, anginf, w.imr_nx, w.imr_ny, w.imr_spx, w.imr_spy, w.imr_phioff, w.imr_xoff, w.imr_yoff,
                          w.rstar, w.imrec_addstar));
    const int nfr = w.nfr;
    const size_t per = (size_t)w.imr_nx * w.imr_ny * nfr;
    std::vector<double> image(per), tauim(per), velo(nfr);
    for (int l = 0; l < w.nlines_render; l++) {
      const int il = w.ilinestart + l;
      int maser = 0;
      printf(" Rendering position-velocity diagram of line %12d\n", il);
      RL(rl_render_rect(ctx, il, 1, nfr, w.passband, image.data(), tauim.data(), &maser));
      const double nu0 = std::fabs(w.linefreq[il - 1]), passb = 3.33567e-6 * nu0 * w.passband;
      for (int k = 0; k < nfr; k++) velo[k] = ((0.0 - passb) + k * (2.0 * passb / (nfr - 1.0))) / w.linefreq[il - 1];
      try {
        rlio::write_posvel(rlio::posvel_filename(w.molname, il), w.molname, "./" + w.molfile, w.dist_cm, w.radvelo,
                           anginf, w.lev_up[il - 1], w.lev_down[il - 1], w.linefreq[il - 1], nfr, w.imr_nx, w.imr_ny,
                           w.imr_spx, w.imr_spy, w.imr_phioff, w.imr_xoff, w.imr_yoff, velo.data(), image.data(),
                           tauim.data());
      } catch (const rlio::Stop &s) {
        fprintf(stderr, " %s\n", s.msg.c_str());
        return s.code & 255 ? s.code & 255 : 13;
      }
      if (maser) printf(" WARNING: Masing detected!\n Will only warn once for this transition.\n");
    }
    rl_destroy(ctx);
    FILE *fs = fopen("radlite.success", "w");
    if (fs) {
      fprintf(fs, " 1\n");
      fclose(fs);
    }
    return 0;
  }

  int nrr = 0, nphi = 0, nray = 0;
  RL(rl_get_camera_dims(ctx, &nrr, &nphi, &nray));
  const int nl = w.nlines_render, nfr = w.nfr;
  const std::string specfile = "linespectrum_" + w.molname + ".dat";
  // main.F:1037 header_line_spectrum
  try {
    rlio::write_spectrum_header(specfile, w.molname, "./" + w.molfile, nl, nfr, w.dist_cm, w.radvelo, anginf,
                                w.style);
  } catch (const rlio::Stop &s) {
    fprintf(stderr, " %s\n", s.msg.c_str());
    return 13;
  }
  if (w.style != 1) {
    fprintf(stderr, " Outputting temperature instead of Fnu only possible in convolution mode\n stop 52987\n");
    return 52987 & 255;
  }
  std::vector<double> rays_r(nrr + 1), ri(nrr + 2);
  RL(rl_get_rings(ctx, rays_r.data(), ri.data()));
  // the line loop main.F:1043-1049; cube mode goes line by line to bound host memory
  const size_t per = (size_t)(nrr + 1) * nphi * nfr;
  const int chunk = imcir ? 1 : nl;
  std::vector<double> flux((size_t)chunk * nfr), velo((size_t)chunk * nfr), tau(chunk), cube;
  std::vector<int> mask, maser(chunk);
  if (imcir) {
    cube.resize(per);
    mask.resize(per);
  }
  for (int l0 = 0; l0 < nl; l0 += chunk) {
    const int n = std::min(chunk, nl - l0);
    RL(rl_render(ctx, w.ilinestart + l0, n, nfr, w.passband, w.dist_cm, flux.data(), imcir ? cube.data() : nullptr,
                 imcir ? mask.data() : nullptr, tau.data(), maser.data(), velo.data()));
    for (int l = 0; l < n; l++) {
      const int il = w.ilinestart + l0 + l;  // 1-based
      try {
        if (imcir)
          rlio::write_imcir(rlio::imcir_filename(w.molname, il), nfr, w.linefreq[il - 1], nphi, nrr, ri.data(),
                            rays_r.data(), velo.data() + (size_t)l * nfr, cube.data(), mask.data());
        rlio::append_line_spectrum(specfile, w.lev_up[il - 1], w.lev_down[il - 1], w.linefreq[il - 1], nfr,
                                   velo.data() + (size_t)l * nfr, flux.data() + (size_t)l * nfr, w.radvelo);
      } catch (const rlio::Stop &s) {
        fprintf(stderr, " %s\n", s.msg.c_str());
        return 13;
      }
      if (maser[l]) printf(" WARNING: Masing detected! (line %d)\n", il);
      printf(" Rendered spectrum of line %12d\n", il);
    }
  }
  double R = 0, E = 0, S = 0;
  rl_get_counters(ctx, &R, &E, &S);
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf(" radlite_b200: %d lines, %d rays, %d channels: %.3g ray-channel integrations in %.3f s (%lld kernel launches)\n",
         nl, nray, nfr, R, dt, rl_launch_count(ctx));
  rl_destroy(ctx);
  FILE *f = fopen("radlite.success", "w");
  if (f) {
    fprintf(f, " 1\n");
    fclose(f);
  }
  return 0;
}

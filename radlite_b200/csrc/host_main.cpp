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

static void append_record(FILE *f, const std::string &name, char dt, const std::vector<long long> &dims,
                          const void *data) {
  char nm[32] = {0};
  snprintf(nm, sizeof nm, "%s", name.c_str());
  const int nd = (int)dims.size();
  size_t n = 1;
  for (long long d : dims) n *= (size_t)d;
  fwrite(nm, 1, 32, f);
  fwrite(&dt, 1, 1, f);
  fwrite(&nd, sizeof nd, 1, f);
  fwrite(dims.data(), sizeof(long long), dims.size(), f);
  fwrite(data, dt == 'd' ? sizeof(double) : sizeof(int), n, f);
}

// --assemble: the part files of a ring-sharded run (one process per GPU, --rings lo:hi --part FILE) -> the
// output files of the unsharded run, byte for byte: the ring sums of disjoint blocks add exactly, the flux is
// the reference's index-ordered sum (telescope.F:1388-1433), cube and mask rows are placed where they belong.
static int assemble(const rlio::WorkDir &w, const std::vector<std::string> &parts, bool imcir) {
  const int nl = w.nlines_render, nfr = w.nfr;
  int nrr = -1, nphi = 0;
  std::vector<std::vector<std::pair<std::string, Rec>>> recs(parts.size());
  std::vector<int> covered;
  for (size_t k = 0; k < parts.size(); k++) {
    if (!load_records(parts[k], recs[k])) {
      fprintf(stderr, " cannot read part file %s\n", parts[k].c_str());
      return 13;
    }
    const Rec *h = find_rec(recs[k], "part");
    if (!h || h->i.size() < 7 || h->i[2] != nl || h->i[3] != nfr || (imcir && !h->i[6])) {
      fprintf(stderr, " part file %s does not belong to this run\n", parts[k].c_str());
      return 13;
    }
    if (nrr < 0) {
      nrr = h->i[4];
      nphi = h->i[5];
      covered.assign(nrr + 1, 0);
    }
    for (int ir = h->i[0]; ir <= h->i[1]; ir++) covered[ir]++;
  }
  for (int ir = 0; ir <= nrr; ir++)
    if (covered[ir] != 1) {
      fprintf(stderr, " ring %d is covered %d times by the part files\n", ir, covered[ir]);
      return 13;
    }
  const Rec *rr = find_rec(recs[0], "rays_r"), *ri = find_rec(recs[0], "imcir_ri");
  const std::string specfile = "linespectrum_" + w.molname + ".dat";
  const double anginf = w.incl_deg * 0.0174532925199;
  try {
    rlio::write_spectrum_header(specfile, w.molname, "./" + w.molfile, nl, nfr, w.dist_cm, w.radvelo, anginf, w.style);
    const size_t per = (size_t)(nrr + 1) * nphi * nfr;
    std::vector<double> ring((size_t)(nrr + 1) * nfr), flux(nfr), velo(nfr), cube;
    std::vector<int> mask;
    if (imcir) {
      cube.resize(per);
      mask.resize(per);
    }
    for (int l = 0; l < nl; l++) {
      const int il = w.ilinestart + l;
      std::fill(ring.begin(), ring.end(), 0.0);
      int maser = 0;
      for (size_t k = 0; k < parts.size(); k++) {
        const Rec *h = find_rec(recs[k], "part");
        const int lo = h->i[0], hi = h->i[1];
        const Rec *rs = find_rec(recs[k], ("ringsum_" + std::to_string(l)).c_str());
        const Rec *ms = find_rec(recs[k], ("maser_" + std::to_string(l)).c_str());
        if (!rs) return 13;
        for (size_t i = 0; i < ring.size(); i++) ring[i] = ring[i] + rs->d[i];  // x + 0 outside the block
        if (ms && ms->i[0]) maser = 1;
        if (imcir) {
          const Rec *cb = find_rec(recs[k], ("cube_" + std::to_string(l)).c_str());
          const Rec *mk = find_rec(recs[k], ("mask_" + std::to_string(l)).c_str());
          if (!cb || !mk) return 13;
          const size_t o = (size_t)lo * nphi * nfr, n = (size_t)(hi - lo + 1) * nphi * nfr;
          std::copy(cb->d.begin(), cb->d.begin() + n, cube.begin() + o);
          std::copy(mk->i.begin(), mk->i.begin() + n, mask.begin() + o);
        }
      }
      const double dist2 = w.dist_cm * w.dist_cm;
      for (int c = 0; c < nfr; c++) {  // telescope.F:1388-1433: centre, then rings 1..nrr, in index order
        double slum = 0.0;
        for (int ir = 0; ir <= nrr; ir++) slum = slum + ring[(size_t)ir * nfr + c];
        flux[c] = slum / dist2;
      }
      const double nu0 = std::fabs(w.linefreq[il - 1]), passb = 3.33567e-6 * nu0 * w.passband;
      const double nu1 = 0.0 - passb, dnu = 2.0 * passb / (nfr - 1.0);
      for (int k = 1; k <= nfr; k++) velo[k - 1] = (nu1 + (k - 1) * dnu) / w.linefreq[il - 1];
      if (imcir)
        rlio::write_imcir(rlio::imcir_filename(w.molname, il), nfr, w.linefreq[il - 1], nphi, nrr, ri->d.data(),
                          rr->d.data(), velo.data(), cube.data(), mask.data());
      rlio::append_line_spectrum(specfile, w.lev_up[il - 1], w.lev_down[il - 1], w.linefreq[il - 1], nfr, velo.data(),
                                 flux.data(), w.radvelo);
      if (maser) printf(" WARNING: Masing detected! (line %d)\n", il);
      printf(" Rendered spectrum of line %12d\n", il);
    }
  } catch (const rlio::Stop &s) {
    fprintf(stderr, " %s\n", s.msg.c_str());
    return 13;
  }
  FILE *f = fopen("radlite.success", "w");
  if (f) {
    fprintf(f, " 1\n");
    fclose(f);
  }
  return 0;
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
  std::string dir, dump, replay, part;
  std::vector<std::string> assemble_parts;
  int ring_lo = -1, ring_hi = -1;
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
    else if (a == "--rings" && i + 1 < argc) {  // camera-ring block lo:hi of a sharded run (0 = central beam)
      if (sscanf(argv[++i], "%d:%d", &ring_lo, &ring_hi) != 2) return 2;
    }
    else if (a == "--part" && i + 1 < argc) part = argv[++i];
    else if (a == "--assemble") {
      while (i + 1 < argc && argv[i + 1][0] != '-') assemble_parts.push_back(argv[++i]);
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
  if (!assemble_parts.empty()) return assemble(w, assemble_parts, imcir);
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
  if (ring_lo >= 0) {
    // one rank of a ring-sharded run: render the block, leave its ring sums (and cube / mask rows) in a part
    // file; `--assemble` of all part files writes the output files of the unsharded run
    if (part.empty() || ring_hi < ring_lo || ring_hi > nrr) {
      fprintf(stderr, " --rings lo:hi needs --part FILE and 0 <= lo <= hi <= %d\n", nrr);
      return 2;
    }
    FILE *pf = fopen(part.c_str(), "wb");
    if (!pf) return 13;
    std::vector<double> rays_r(nrr + 1), ri(nrr + 2);
    RL(rl_get_rings(ctx, rays_r.data(), ri.data()));
    const int hdr[7] = {ring_lo, ring_hi, nl, nfr, nrr, nphi, imcir ? 1 : 0};
    append_record(pf, "part", 'i', {7}, hdr);
    append_record(pf, "rays_r", 'd', {nrr + 1}, rays_r.data());
    append_record(pf, "imcir_ri", 'd', {nrr + 2}, ri.data());
    const size_t per = (size_t)(nrr + 1) * nphi * nfr, slab = (size_t)(ring_hi - ring_lo + 1) * nphi * nfr;
    std::vector<double> ring((size_t)(nrr + 1) * nfr), cube;
    std::vector<int> mask;
    if (imcir) {
      cube.resize(per);
      mask.resize(per);
    }
    for (int l = 0; l < nl; l++) {
      double tau = 0.0;
      int maser = 0;
      RL(rl_render_rings_cube(ctx, w.ilinestart + l, 1, nfr, w.passband, w.dist_cm, ring_lo, ring_hi, ring.data(),
                              imcir ? cube.data() : nullptr, imcir ? mask.data() : nullptr, &tau, &maser));
      append_record(pf, "ringsum_" + std::to_string(l), 'd', {nrr + 1, nfr}, ring.data());
      append_record(pf, "maser_" + std::to_string(l), 'i', {1}, &maser);
      if (imcir) {
        append_record(pf, "cube_" + std::to_string(l), 'd', {(long long)slab}, cube.data() + (size_t)ring_lo * nphi * nfr);
        append_record(pf, "mask_" + std::to_string(l), 'i', {(long long)slab}, mask.data() + (size_t)ring_lo * nphi * nfr);
      }
    }
    fclose(pf);
    rl_destroy(ctx);
    return 0;
  }
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

// rl_io.cpp -- RADLite file formats for the stand-alone host program (see rl_io.h).
#include "rl_io.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rlio {

namespace {

bool file_exists(const std::string &p) {
  std::ifstream f(p.c_str());
  return f.good();
}

[[noreturn]] void stop(int code, const std::string &msg) { throw Stop{code, msg}; }

// Fortran real literal -> double ("1.d-3", "1.0D+00", "2.5e3", "7")
bool parse_real(const std::string &t, double &v) {
  if (t.empty()) return false;
  std::string s = t;
  for (char &c : s)
    if (c == 'd' || c == 'D') c = 'e';
  char *end = nullptr;
  v = std::strtod(s.c_str(), &end);
  return end && *end == '\0';
}

}  // namespace

// ---- FortranReader ------------------------------------------------------------------------------
FortranReader::FortranReader(const std::string &path) : path_(path) {
  if (path.compare(0, 7, "string:") == 0) {  // in-memory "file" (strip_comments output, tools.F:6)
    buf_ = path.substr(7);
    path_ = "temp.inp";
    ok_ = true;
    return;
  }
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f.good()) return;
  std::stringstream ss;
  ss << f.rdbuf();
  buf_ = ss.str();
  ok_ = true;
}

bool FortranReader::next_token(std::string &tok) {
  if (repeat_ > 0) {
    repeat_--;
    tok = repeat_tok_;
    return true;
  }
  const size_t n = buf_.size();
  while (pos_ < n && (buf_[pos_] == ' ' || buf_[pos_] == '\t' || buf_[pos_] == '\r' || buf_[pos_] == '\n' ||
                      buf_[pos_] == ','))
    pos_++;
  if (pos_ >= n) return false;
  tok.clear();
  if (buf_[pos_] == '\'' || buf_[pos_] == '"') {
    const char q = buf_[pos_++];
    while (pos_ < n && buf_[pos_] != q && buf_[pos_] != '\n') tok += buf_[pos_++];
    if (pos_ < n && buf_[pos_] == q) pos_++;
    return true;
  }
  while (pos_ < n && !(buf_[pos_] == ' ' || buf_[pos_] == '\t' || buf_[pos_] == '\r' || buf_[pos_] == '\n' ||
                       buf_[pos_] == ','))
    tok += buf_[pos_++];
  const size_t star = tok.find('*');
  if (star != std::string::npos && star > 0) {  // r*c
    bool digits = true;
    for (size_t i = 0; i < star; i++) digits = digits && std::isdigit((unsigned char)tok[i]);
    if (digits) {
      repeat_ = std::atol(tok.substr(0, star).c_str()) - 1;
      repeat_tok_ = tok.substr(star + 1);
      tok = repeat_tok_;
    }
  }
  return true;
}

void FortranReader::end_statement() {
  repeat_ = 0;
  const size_t n = buf_.size();
  while (pos_ < n && buf_[pos_] != '\n') pos_++;
  if (pos_ < n) pos_++;
}

std::vector<double> FortranReader::reals(size_t n) {
  std::vector<double> out(n);
  std::string t;
  for (size_t i = 0; i < n; i++) {
    if (!next_token(t)) stop(13, "Prematurely reached end of file " + path_);
    if (!parse_real(t, out[i])) stop(13, "Could not understand contents of " + path_ + " ('" + t + "')");
  }
  end_statement();
  return out;
}

std::vector<long> FortranReader::integers(size_t n) {
  std::vector<double> v = reals(n);
  std::vector<long> out(n);
  for (size_t i = 0; i < n; i++) out[i] = (long)v[i];
  return out;
}

long FortranReader::integer() { return integers(1)[0]; }

std::string FortranReader::word() {
  std::string t;
  if (!next_token(t)) stop(13, "Prematurely reached end of file " + path_);
  end_statement();
  return t;
}

std::string FortranReader::record() {
  const size_t n = buf_.size();
  if (pos_ >= n) stop(13, "Prematurely reached end of file " + path_);
  size_t e = pos_;
  while (e < n && buf_[e] != '\n') e++;
  std::string rec = buf_.substr(pos_, e - pos_);
  if (!rec.empty() && rec.back() == '\r') rec.pop_back();
  pos_ = (e < n) ? e + 1 : e;
  return rec;
}

static std::string field(const std::string &rec, int col, int width) {
  if ((size_t)(col - 1) >= rec.size()) return "";
  std::string f = rec.substr(col - 1, width);
  std::string out;
  for (char c : f)
    if (c != ' ' && c != '\t') out += c;  // blanks inside numeric fields are ignored (BN)
  return out;
}
double field_real(const std::string &rec, int col, int width) {
  const std::string f = field(rec, col, width);
  double v = 0.0;
  if (f.empty()) return 0.0;
  if (!parse_real(f, v)) stop(13, "bad numeric field '" + f + "'");
  return v;
}
long field_int(const std::string &rec, int col, int width) {
  const std::string f = field(rec, col, width);
  if (f.empty()) return 0;
  return std::atol(f.c_str());
}

// ---- readers ---------------------------------------------------------------------------------------
namespace {

// tools.F:6-16 strip_comments: drop lines starting with # ; % = and keep the first 16 characters
std::string strip_comments(const std::string &path) {
  std::ifstream f(path.c_str());
  std::string line, out;
  while (std::getline(f, line)) {
    if (!line.empty() && (line[0] == '#' || line[0] == ';' || line[0] == '%' || line[0] == '=')) continue;
    out += line.substr(0, 16);
    out += '\n';
  }
  return out;
}

// setup.F:1584-1690 read_scalar_field (mirror flag required, MIRROR_THETA): returns [nr][nth]
std::vector<double> read_scalar_field(const std::string &file, int nr_grid, int nth_grid) {
  FortranReader f(file);
  if (!f.is_open()) stop(13, "Could not open file " + file);
  std::vector<long> h = f.integers(3);
  const int nr = (int)h[0], nt = (int)h[1], imirt = (int)h[2];
  if (imirt == 0) stop(13, "ERROR: RADICAL is compiled with MIRROR_THETA (" + file + ")");
  if (nr != nr_grid || nt != nth_grid) stop(13, "dimensions of " + file + " are inconsistent with radius.inp / theta.inp");
  std::vector<double> v((size_t)nr * nt);
  for (size_t i = 0; i < v.size(); i++) v[i] = f.real();
  return v;
}

void read_radlite_inp(WorkDir &w) {  // main.F:410-686, setup.F:1117-1137
  FortranReader f("string:" + strip_comments("radlite.inp"));
  w.input_format = (int)f.integer();
  if (w.input_format < 100) stop(13, "ERROR while reading radlite.inp: This format version is not supported.");
  try {
    f.integer();                               // max nr of iterations
    f.integer();                               // method of iteration
    if (w.input_format >= 103) f.integer();    // flux conservation tricks
    f.real();                                  // convergence tolerance
    f.integer();                               // convergence style
    f.integer();                               // initial guess style
    f.integer();                               // anggrid_frsizemu
    f.integer();                               // anggrid_frsizephi
    f.integer();                               // anggrid_mu_type
    f.real();                                  // anggrid_muerr_max
    f.real();                                  // dmuvdr
    f.integer();                               // iextrmu
    w.out_itype = (int)f.integer();            // iradbnd_out_itype
    w.in_itype = (int)f.integer();             // iradbnd_in_itype
    f.integer();                               // equatorial boundary
    f.real();                                  // inclination (overridden by linespectrum.inp, main.F:212)
    w.nphi = (int)f.integer();                 // main_nrphiinf
    w.dbdr = (int)f.integer();                 // telesc_dbdr
    w.nrext = (int)f.integer();                // main_nrrayextra
    for (int i = 0; i < 4; i++) f.integer();   // isave_scatnonlte, intens_inu, source, intstore
    if (w.input_format >= 101) f.integer();    // tempstore
    if (w.input_format >= 102) f.integer();    // isave_alioper
    f.integer();                               // isave_physvar
    if (w.input_format >= 104) f.integer();    // isave_fluxcons
  } catch (const Stop &) {
    stop(133, "ERROR in reading input file radlite.inp");
  }
  if (w.nrext == 0) stop(1, "Must set nr of rays inward of Rin to a value > 0");
  const long wp_type = f.integer();
  if (wp_type != -551) stop(1, "Dont know setup type (only -551 is supported)");
  w.in_itype = 2;  // setup.F:1119
  w.do_dust = (int)f.integer();
  w.do_lines = (int)f.integer();
  w.dust_in_lines = (int)f.integer();
  w.star_pump = (int)f.integer();
  if (w.do_dust != 0 && w.do_lines != 0) stop(1, "ERROR: Cannot do dust and lines together...");
  if (w.do_lines == 0) stop(13, "Line transfer is not active");
}

void read_grids(WorkDir &w) {  // grid.F:711-727, 1098-1146, 1451-1470
  {
    FortranReader f("radius.inp");
    if (!f.is_open()) stop(13, "Could not open radius.inp");
    const long n = f.integer();
    w.r.resize(n);
    for (long i = 0; i < n; i++) w.r[i] = f.real();
  }
  {
    FortranReader f("theta.inp");
    if (!f.is_open()) stop(13, "Could not open theta.inp");
    std::vector<long> h = f.integers(2);
    if (h[1] == 0) stop(13, "theta.inp: compiled with MIRROR_THETA, file must have mirror flag 1");
    w.theta.resize(h[0]);
    for (long i = 0; i < h[0]; i++) {
      w.theta[i] = f.real();
      if (w.theta[i] > 1.5707963268 || w.theta[i] < 0.0) stop(13, "theta.inp: theta out of range");
    }
  }
  {
    FortranReader f("frequency.inp");
    if (f.is_open()) {
      const long n = f.integer();
      if (n > 1) {
        w.cont_freq.resize(n);
        for (long i = 0; i < n; i++) w.cont_freq[i] = f.real();
      }
    }
  }
}

void read_dust(WorkDir &w) {  // setup.F:681-735, 159-246, 295-423; source.F:494-529; dust.F:92-292, 317-518
  const int nr = (int)w.r.size(), nth = (int)w.theta.size(), ncf = (int)w.cont_freq.size();
  std::string file = "dustdensity.inp";
  if (!file_exists(file)) file = "dustdens.inp";
  if (!file_exists(file)) stop(1, "ERROR: dustdensity.inp nor dustdens.inp found");
  {
    FortranReader f(file);
    std::vector<long> h = f.integers(4);  // nv nr nt imirt
    if (h[3] == 0) stop(13, "dust density file must have the mirror flag set");
    if (h[1] != nr || h[2] != nth) stop(13, "ERROR: The dimensions of dustdensity.inp are inconsistent");
    w.nspec = (int)h[0];
    w.dust_rho.assign((size_t)nr * nth * w.nspec, 0.0);
    for (int is = 0; is < w.nspec; is++)  // isep = 1: species by species
      for (int ir = 0; ir < nr; ir++)
        for (int it = 0; it < nth; it++) w.dust_rho[((size_t)ir * nth + it) * w.nspec + is] = f.real();
  }
  if (file_exists("scatsource.dat") || file_exists("quantsource.dat")) {
    w.scat.assign((size_t)nr * nth * ncf, 0.0);
    for (const char *name : {"scatsource.dat", "quantsource.dat"}) {
      if (!file_exists(name)) continue;
      FortranReader f(name);
      f.integers(4);
      for (int k = 0; k < ncf; k++)
        for (int it = 0; it < nth; it++)
          for (int ir = 0; ir < nr; ir++) w.scat[((size_t)ir * nth + it) * ncf + k] += f.real();
    }
  }
  // dust temperatures: dusttemp.info -> file (setup.F:186-205)
  if (!file_exists("dusttemp.info")) stop(13, "Could not find dusttemp.info");
  std::string tfile;
  {
    FortranReader f("dusttemp.info");
    const long flnr = f.integer();
    if (flnr == -2) tfile = "dusttemp_final.dat";
    else tfile = "dusttemp_" + std::to_string(flnr) + ".dat";
  }
  std::vector<std::vector<double>> temps;  // [spec][size][ir][it]
  {
    FortranReader f(tfile);
    if (!f.is_open()) stop(13, "Could not open file " + tfile);
    std::vector<long> h = f.integers(4);
    if (h[3] == 0) stop(13, "dust temperature file must have the mirror flag set");
    if (h[1] != nr || h[2] != nth) stop(13, "ERROR: The dimensions of dusttemp are inconsistent");
    if (h[0] != w.nspec) stop(99, "dust temperature file: wrong number of species");
    w.nsize.assign(w.nspec, 0);
    temps.resize(w.nspec);
    for (int is = 0; is < w.nspec; is++) {
      w.nsize[is] = (int)f.integer();
      temps[is].resize((size_t)w.nsize[is] * nr * nth);
      for (size_t i = 0; i < temps[is].size(); i++) temps[is][i] = f.real();
    }
  }
  // dustopac.inp + dustopac_<n>.inp
  {
    FortranReader f("dustopac.inp");
    if (!f.is_open()) stop(13, "Could not open dustopac.inp");
    const long iformat = f.integer();
    const long nsp = f.integer();
    f.word();
    if (nsp != w.nspec) stop(99, "dustopac.inp and the dust density file disagree on the number of species");
    w.maxsize = 0;
    std::vector<std::vector<double>> ka(nsp), ks(nsp);
    std::vector<int> nsz(nsp);
    for (long is = 0; is < nsp; is++) {
      const long idum = f.integer();
      if (iformat >= 2) {
        const long idum2 = f.integer();
        if (idum2 == 2) f.real();
        else if (idum2 == 3) { f.real(); f.real(); }
      }
      const long idustfile = f.integer();
      if (idum == -2) { f.real(); f.real(); }
      else if (idum == -3) { f.real(); f.real(); f.real(); f.real(); }
      else if (idum != -1) stop(13, "dustopac.inp: input modes other than -1, -2, -3 not implemented");
      const std::string ofile = "dustopac_" + std::to_string(idustfile) + ".inp";
      FortranReader g(ofile);
      if (!g.is_open()) stop(13, "Could not open " + ofile);
      std::vector<long> h = g.integers(2);
      if (h[0] == -1) stop(13, ofile + ": temperature-dependent opacities need DUST_OPAC_TEMPDEP");
      if (h[0] != ncf) stop(13, ofile + ": number of frequencies differs from frequency.inp");
      nsz[is] = (int)h[1];
      ka[is].resize((size_t)ncf * nsz[is]);
      ks[is].resize((size_t)ncf * nsz[is]);
      for (int k = 0; k < ncf; k++)
        for (int iz = 0; iz < nsz[is]; iz++) ka[is][(size_t)iz * ncf + k] = g.real();
      for (int k = 0; k < ncf; k++)
        for (int iz = 0; iz < nsz[is]; iz++) ks[is][(size_t)iz * ncf + k] = g.real();
      if (nsz[is] != w.nsize[is]) stop(13, ofile + ": number of grain sizes differs from the dust temperature file");
      w.maxsize = std::max(w.maxsize, nsz[is]);
      f.word();
    }
    w.kappa_abs.assign((size_t)nsp * w.maxsize * ncf, 0.0);
    w.kappa_scat.assign((size_t)nsp * w.maxsize * ncf, 0.0);
    for (long is = 0; is < nsp; is++)
      for (int iz = 0; iz < nsz[is]; iz++)
        for (int k = 0; k < ncf; k++) {
          w.kappa_abs[((size_t)is * w.maxsize + iz) * ncf + k] = ka[is][(size_t)iz * ncf + k];
          w.kappa_scat[((size_t)is * w.maxsize + iz) * ncf + k] = ks[is][(size_t)iz * ncf + k];
        }
  }
  w.dust_temp.assign((size_t)nr * nth * w.nspec * w.maxsize, 0.0);
  for (int is = 0; is < w.nspec; is++)
    for (int iz = 0; iz < w.nsize[is]; iz++)
      for (int ir = 0; ir < nr; ir++)
        for (int it = 0; it < nth; it++)
          w.dust_temp[(((size_t)ir * nth + it) * w.nspec + is) * w.maxsize + iz] =
              temps[is][((size_t)iz * nr + ir) * nth + it];
}

void read_gas(WorkDir &w) {  // line.F:3918; setup.F:1453, 864, 754, 803
  const int nr = (int)w.r.size(), nth = (int)w.theta.size();
  const size_t nc = (size_t)nr * nth;
  {
    FortranReader f("line.inp");
    if (!f.is_open()) stop(13, "ERROR: Could not find line.inp");
    f.integer();
    f.integer();
    w.umass_av = f.real();
  }
  if (!file_exists("density.inp")) stop(1, "ERROR: density.inp not found");
  w.rho = read_scalar_field("density.inp", nr, nth);
  w.abund.assign(nc, 0.0);
  if (file_exists("abundance.inp")) {
    FortranReader f("abundance.inp");
    std::vector<long> h = f.integers(2);
    if (h[0] != nr || h[1] != nth) stop(325, "abundance.inp: wrong dimensions");
    for (size_t i = 0; i < nc; i++) w.abund[i] = f.reals(2)[0];
  }
  w.vel.assign(3 * nc, 0.0);
  if (file_exists("velocity.inp")) {
    FortranReader f("velocity.inp");
    std::vector<long> h = f.integers(2);
    if (h[0] != nr || h[1] != nth) stop(325, "velocity.inp: wrong dimensions");
    for (size_t i = 0; i < nc; i++) {
      std::vector<double> v = f.reals(3);
      w.vel[3 * i] = v[0];
      w.vel[3 * i + 1] = v[1];
      w.vel[3 * i + 2] = v[2];
    }
  }
  w.linewidth.assign(nc, 0.0);
  if (file_exists("turbulence.inp")) {
    FortranReader f("turbulence.inp");
    f.integer();
    std::vector<long> h = f.integers(2);
    if (h[0] != nr || h[1] != nth) stop(325, "turbulence.inp: wrong dimensions");
    for (size_t i = 0; i < nc; i++) w.linewidth[i] = f.real();
  }
}

void read_molecule(WorkDir &w) {  // line.F:1826-1985 read_species_lambda
  FortranReader f("./" + w.molfile);
  if (!f.is_open()) stop(13, "Could not open the molecular data file " + w.molfile);
  f.word();
  f.word();
  f.word();
  w.umass_molec = field_real(f.record(), 1, 4);  // (F4.1)
  f.word();
  w.nlev_orig = (int)field_int(f.record(), 1, 6);  // (I6)
  w.nlevels = w.nlev_orig;
  f.word();
  if (w.nlev_orig < 2) stop(13, "Minimum of 2 levels!");
  w.ener_cm.resize(w.nlev_orig);
  w.gdeg.resize(w.nlev_orig);
  for (int i = 0; i < w.nlev_orig; i++) {  // (I5,F12.4,F7.1)
    const std::string rec = f.record();
    w.ener_cm[i] = field_real(rec, 6, 12);
    w.gdeg[i] = field_real(rec, 18, 7);
  }
  f.word();
  w.nlines = (int)field_int(f.record(), 1, 6);
  if (w.nlines < 1) stop(13, "Minimum of 1 line!");
  f.word();
  w.lev_up.resize(w.nlines);
  w.lev_down.resize(w.nlines);
  w.aud.resize(w.nlines);
  w.linefreq.resize(w.nlines);
  for (int i = 0; i < w.nlines; i++) {  // (I5,I5,I5,E12.3)
    const std::string rec = f.record();
    w.lev_up[i] = (int)field_int(rec, 6, 5);
    w.lev_down[i] = (int)field_int(rec, 11, 5);
    w.aud[i] = field_real(rec, 16, 12);
    if (w.lev_up[i] < 1 || w.lev_up[i] > w.nlev_orig || w.lev_down[i] < 1 || w.lev_down[i] > w.nlev_orig)
      stop(13, "line levels should be within the range 1..nlevels");
  }
  for (int i = 0; i < w.nlines; i++) {  // line.F:1903, 1981
    const double eu = 1.986468498e-16 * w.ener_cm[w.lev_up[i] - 1];
    const double ed = 1.986468498e-16 * w.ener_cm[w.lev_down[i] - 1];
    w.linefreq[i] = 1.509160e26 * (eu - ed);
  }
}

void read_linespectrum_inp(WorkDir &w) {  // telescope.F:86-238
  FortranReader f("linespectrum.inp");
  if (!f.is_open()) stop(13, "Could not find linespectrum.inp");
  const long iformat = f.integer();
  w.style = 0;
  if (iformat == 1) w.style = (int)f.integer();
  if (iformat < 0 || iformat > 1) stop(13, "linespectrum.inp: unknown format");
  f.word();
  const long iformtel = f.integer();
  if (iformtel != 2) stop(13, "linespectrum.inp: unknown telescope format");
  f.word();
  w.vmax = f.real();
  w.dv = f.real();
  f.word();
  const long iformobj = f.integer();
  if (iformobj == 1) {
    stop(13, "linespectrum.inp object format 1 carries no molecular data file");
  } else if (iformobj == 2) {
    const std::string s = f.word();
    const size_t p = s.rfind(".dat");
    if (p == std::string::npos) stop(13, "linespectrum.inp: molecular data file must end in .dat");
    w.molfile = s.substr(0, p + 4);
    w.molname = s.substr(0, p);
    read_molecule(w);
    w.command = (int)f.integer();
    f.real();  // distance: forced to 1 pc (main.F:206)
    w.incl_deg = f.real();
    w.radvelo = f.real();
    w.nlines_render = (int)f.integer();
    w.ilinestart = (int)f.integer();
    if (w.command == 1) stop(1, "frequency-integrated images are not available");
    if (w.command != 0 && w.command != 2) stop(13, "linespectrum.inp: unknown command");
    if (w.command == 2) {  // telescope.F:203-228
      w.imr_nx = (int)f.integer();
      w.imr_ny = (int)f.integer();
      const long isizespecifier = f.integer();
      double szimx = f.real(), szimy = f.real();
      szimx = szimx * 0.5;
      szimy = szimy * 0.5;
      w.imr_spx = szimx / (w.imr_nx * 0.5);
      w.imr_spy = szimy / (w.imr_ny * 0.5);
      w.imr_phioff = f.real();
      w.imr_xoff = f.real();
      w.imr_yoff = f.real();
      w.imrec_addstar = (int)f.integer();
      if (isizespecifier != 0 && isizespecifier != 1) stop(13, "linespectrum.inp: unknown size specifier");
    }
  } else {
    stop(13, "linespectrum.inp: unknown object format");
  }
  if (w.ilinestart + w.nlines_render - 1 > w.nlines || w.ilinestart < 1)
    stop(13, "linespectrum.inp: line range exceeds the lines of the molecular data file");
}

void read_line_inp(WorkDir &w) {  // line.F:142-260: only line_rangewidth(1) can matter (telescope.F:232)
  FortranReader f("line.inp");
  const long iformat = f.integer();
  if (iformat == 2) f.integer();
  else if (iformat != 1) stop(23223, "line.inp: unknown format");
  w.umass_av = f.real();
  f.word();
  f.word();
  const long iinfo = f.integer();
  if (iinfo != 1) stop(13, "line.inp: the first line must carry its information");
  if (f.integer() != 0) stop(13, "line.inp: line symmetry not supported");
  f.integer();
  w.rangewidth1 = f.real();
}

void read_star(WorkDir &w) {  // star.F:576-617, 449-528, 675-711
  if (file_exists("starinfo.inp")) {
    FortranReader f("starinfo.inp");
    f.integer();
    w.rstar = f.real();
  } else {
    w.rstar = w.r[0] * 1.e-3;
  }
  const int ncf = (int)w.cont_freq.size();
  w.starspec.assign(ncf, 0.0);
  if (w.star_pump != 0) {
    const bool fold = file_exists("starspec.inp"), fnew = file_exists("starspectrum.inp");
    if (fold && fnew) stop(13, "PROBLEM: both starspec.inp and starspectrum.inp are present");
    if (fold) {
      FortranReader f("starspec.inp");
      if (f.integer() != ncf) stop(13, "starspec.inp: number of frequencies differs from frequency.inp");
      for (int i = 0; i < ncf; i++) w.starspec[i] = f.real();
    } else if (fnew) {
      FortranReader f("starspectrum.inp");
      if (f.integer() != ncf) stop(13, "starspectrum.inp: number of frequencies differs from frequency.inp");
      for (int i = 0; i < ncf; i++) {
        std::vector<double> v = f.reals(2);
        if (ncf > 1 && std::fabs(v[0] - w.cont_freq[i]) / (v[0] + w.cont_freq[i]) > 1.e-3)
          stop(13, "PROBLEM: Frequency grid of stellar spectrum unequal to frequency.inp");
        w.starspec[i] = 3.0308410e36 * v[1] / (w.rstar * w.rstar);
      }
    } else {
      stop(13, "PROBLEM: neither starspec.inp nor starspectrum.inp found");
    }
  }
  if (w.out_itype == 3) {
    FortranReader f("interstellfield.inp");
    if (!f.is_open()) stop(13, "ERROR: outer BC type 3 but no interstellfield.inp");
    if (f.integer() != ncf) stop(13, "interstellfield.inp: inconsistent nr of frequencies");
    w.isrf.resize(ncf);
    for (int i = 0; i < ncf; i++) w.isrf[i] = f.real();
  }
}

void read_levelpop(WorkDir &w) {  // line.F:811-950 (filenr < -5: levelpop.info decides)
  const int nr = (int)w.r.size(), nth = (int)w.theta.size();
  FortranReader info("levelpop.info");
  if (!info.is_open()) stop(13, "Could not open levelpop.info");
  const long filenr = info.integer();
  std::string file;
  if (filenr == -3) {
    const std::string s = info.word();
    if (s.compare(0, 9, "levelpop_") != 0 || s.find(".dat") == std::string::npos)
      stop(13, "levelpop.info: file name must be levelpop_<molecule>.dat");
    file = "levelpop_" + w.molname + ".dat";
  } else if (filenr == -1) {
    file = "levelpop.dat";
  } else if (filenr == -2) {
    file = "levelpop_final.dat";
  } else {
    file = "levelpop_" + std::to_string(filenr) + ".dat";
  }
  FortranReader f(file);
  if (!f.is_open()) stop(13, "Could not open " + file);
  std::vector<long> h = f.integers(4);
  if (h[0] != nr || h[1] != nth || h[3] != 1) stop(13, file + ": grid dimensions differ from radius.inp / theta.inp");
  w.nlevels = (int)h[2];
  if (w.nlevels > w.nlev_orig || w.nlevels < 2) stop(13, file + ": number of levels differs from the molecular data file");
  f.reals(w.nlevels);
  f.reals(w.nlevels);
  w.popul.resize((size_t)nr * nth * w.nlevels);
  for (size_t c = 0; c < (size_t)nr * nth; c++) {
    std::vector<double> v = f.reals(w.nlevels);
    std::copy(v.begin(), v.end(), w.popul.begin() + c * w.nlevels);
  }
}

}  // namespace

WorkDir read_workdir() {
  WorkDir w;
  if (!file_exists("radlite.inp")) stop(13, "ERROR: Could not find the main input file: radlite.inp");
  read_radlite_inp(w);      // main.F:112-116, 410
  read_grids(w);            // main.F:276 create_all_grids
  read_gas(w);              // setup.F:1068 read_compute_medium_lean (-551)
  if (w.dust_in_lines != 0) read_dust(w);
  else stop(13, "line transfer without dust (iradproc_line_dust = 0) is not supported by this host");
  read_linespectrum_inp(w); // main.F:202 read_telesc_linespec (+ read_species_lambda)
  read_line_inp(w);         // main.F:203 read_linedata(0)
  // main.F:204-213
  w.ilinestart = 1;
  w.nfr = (int)(2.0 * w.vmax / w.dv + 1);
  if (w.vmax <= 0.0) w.vmax = w.rangewidth1;  // telescope.F:231-233
  w.nfr = (int)(2.0 * w.vmax / w.dv + 1);
  w.vmax = 0.5 * (w.nfr - 1) * w.dv;
  w.passband = w.vmax;
  w.dist_cm = 1.0 * 3.08572e18;
  read_star(w);             // main.F:926, 937
  read_levelpop(w);         // main.F:970 read_nonlte_scat -> read_levelpopul
  for (int i = 0; i < w.nlines; i++)
    if (w.lev_up[i] > w.nlevels || w.lev_down[i] > w.nlevels)
      stop(13, "a line refers to a level that the level-population file does not hold");
  return w;
}

// ---- gfortran-compatible formatting ------------------------------------------------------------
static std::string pad_left(const std::string &s, int w) {
  if ((int)s.size() >= w) return s;
  return std::string(w - s.size(), ' ') + s;
}

std::string fmt_i(long v, int w) {
  std::string s = std::to_string(v);
  if ((int)s.size() > w) return std::string(w, '*');
  return pad_left(s, w);
}

std::string fmt_f(double v, int w, int d) {
  char b[512];
  snprintf(b, sizeof b, "%.*f", d, v);
  std::string s = b;
  if ((int)s.size() > w) {  // the optional leading zero goes first
    if (s.compare(0, 2, "0.") == 0) s = s.substr(1);
    else if (s.compare(0, 3, "-0.") == 0) s = "-" + s.substr(2);
  }
  if ((int)s.size() > w) return std::string(w, '*');
  return pad_left(s, w);
}

// digits and decimal exponent of |v| in the Fortran normalisation 0.d1d2..dn x 10^e
static void e_digits(double v, int d, std::string &digits, int &e) {
  if (v == 0.0) {
    digits.assign(d, '0');
    e = 0;
    return;
  }
  char b[64];
  snprintf(b, sizeof b, "%.*e", d - 1, std::fabs(v));
  digits.clear();
  const char *p = b;
  for (; *p && *p != 'e'; p++)
    if (std::isdigit((unsigned char)*p)) digits += *p;
  e = std::atoi(p + 1) + 1;
}

static std::string e_exponent(int e, int ewidth) {
  char b[16];
  const char sg = e < 0 ? '-' : '+';
  const int a = std::abs(e);
  if (ewidth == 2) {
    if (a <= 99) snprintf(b, sizeof b, "E%c%02d", sg, a);
    else snprintf(b, sizeof b, "%c%03d", sg, a);  // Ew.d drops the letter for three-digit exponents
  } else {
    snprintf(b, sizeof b, "E%c%0*d", sg, ewidth, a);
  }
  return b;
}

std::string fmt_e(double v, int w, int d) {
  std::string digits;
  int e;
  e_digits(v, d, digits, e);
  const std::string body = "." + digits + e_exponent(e, 2);
  const std::string sign = (std::signbit(v) && v != 0.0) ? "-" : "";
  std::string s = sign + "0" + body;
  if ((int)s.size() > w) s = sign + body;  // drop the optional leading zero
  if ((int)s.size() > w) return std::string(w, '*');
  return pad_left(s, w);
}

// ESw.d: one non-zero digit before the point; three-digit exponents drop the letter
std::string fmt_es(double v, int w, int d) {
  char b[64];
  snprintf(b, sizeof b, "%.*E", d, v);
  std::string s = b;
  const size_t e = s.find('E');
  if (e != std::string::npos && s.size() - e - 2 > 2) {  // E+123 -> +123 ; C pads the exponent to two digits only
    s = s.substr(0, e) + s.substr(e + 1);
  }
  if ((int)s.size() > w) return std::string(w, '*');
  return pad_left(s, w);
}

// list-directed doubleprecision: 17 significant digits; F layout inside 0.1 <= |x| < 1e16 (blank
// exponent field), otherwise 1PE form with a three-digit exponent; field width 25 + one leading blank
std::string fmt_list_real(double v) {
  char b[64];
  const double a = std::fabs(v);
  if (v == 0.0) {
    snprintf(b, sizeof b, "%.16f", 0.0);
    return "   " + std::string(b) + "     ";
  }
  if (a >= 0.1 && a < 1.e16) {
    int k = (int)std::floor(std::log10(a)) + 1;  // digits before the point
    if (k < 0) k = 0;
    snprintf(b, sizeof b, "%.*f", 17 - (k > 0 ? k : 1), v);
    std::string s = b;
    if ((int)s.size() < 20) s = std::string(20 - s.size(), ' ') + s;
    return "  " + s + "     ";
  }
  snprintf(b, sizeof b, "%.16E", v);
  std::string s = b;  // d.ddddE+XX -> three-digit exponent
  const size_t p = s.find('E');
  const int ex = std::atoi(s.c_str() + p + 1);
  char eb[16];
  snprintf(eb, sizeof eb, "E%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
  s = s.substr(0, p) + eb;
  return pad_left(s, 26);
}

std::string fmt_list_int(long v) { return pad_left(std::to_string(v), 12); }

// ---- writers -----------------------------------------------------------------------------------------
void write_spectrum_header(const std::string &file, const std::string &molname, const std::string &molfile,
                           int nlinespec, int nfrmax, double dist_cm, double radvelo, double anginf,
                           int iformout) {
  FILE *f = fopen(file.c_str(), "w");
  if (!f) stop(13, "cannot open " + file);
  auto a80 = [](const std::string &s) {
    std::string t = s.substr(0, 80);
    return t + std::string(80 - t.size(), ' ');
  };
  fprintf(f, "%s\n", fmt_i(1, 2).c_str());
  fprintf(f, "%s\n", fmt_i(iformout, 2).c_str());
  fprintf(f, "%s\n", a80(molname).c_str());
  fprintf(f, "%s\n", a80(molfile).c_str());
  fprintf(f, "%s\n%s\n", fmt_i(nlinespec, 10).c_str(), fmt_i(nfrmax, 10).c_str());
  fprintf(f, "%s%s%s\n", fmt_e(dist_cm / 3.08572e18, 12, 4).c_str(), fmt_e(radvelo, 12, 4).c_str(),
          fmt_f(anginf * 57.2957795132, 7, 3).c_str());
  fclose(f);
}

void append_line_spectrum(const std::string &file, int lev_up, int lev_down, double linefreq, int nfr,
                          const double *velo, const double *flux, double radvelo) {
  FILE *f = fopen(file.c_str(), "a");
  if (!f) stop(13, "cannot open " + file);
  fprintf(f, "\n");
  fprintf(f, "%s%s\n", fmt_i(lev_up, 5).c_str(), fmt_i(lev_down, 5).c_str());
  fprintf(f, "%s\n", fmt_e(linefreq, 14, 9).c_str());
  fprintf(f, "%s\n", fmt_e(0.0, 10, 5).c_str());  // linespec_beamsize: convolution is switched off
  fprintf(f, "%s\n", fmt_i(nfr, 5).c_str());
  fprintf(f, "  \n");
  std::string out;
  out.reserve((size_t)nfr * 29);
  for (int inu = nfr - 1; inu >= 0; inu--) {  // written inu = nfr..1 (ascending velocity)
    // telescope.F:1771-1775 with spec_freq = nu0 + line_dnu: velo = -c (spec_freq - nu0)/nu0 + v_lsr
    const double spec_freq = linefreq + velo[inu] * linefreq;
    double v = spec_freq - linefreq;
    v = -2.99792458e5 * v / linefreq;
    v = v + radvelo;
    out += fmt_e(v, 13, 6);
    out += ' ';
    out += fmt_e(flux[inu], 13, 6);
    out += '\n';
  }
  fwrite(out.data(), 1, out.size(), f);
  fclose(f);
}

std::string imcir_filename(const std::string &molname, int iline) {
  // telescope.F:1596-1613: only iline < 100 gets a defined name (the reference leaves it unset beyond)
  return "lineposvelcirc_" + molname + "_" + std::to_string(iline) + ".dat";
}

std::string posvel_filename(const std::string &molname, int iline) {
  if (iline >= 100) stop(177, "lineposvel file names exist for lines 1..99 only (telescope.F:1967)");
  return "lineposvel_" + molname + "_" + std::to_string(iline) + ".dat";
}

void write_posvel(const std::string &file, const std::string &molname, const std::string &molfile, double dist_cm,
                  double radvelo, double anginf, int lev_up, int lev_down, double linefreq, int nfr, int nx, int ny,
                  double spx, double spy, double phioff, double xoff, double yoff, const double *velo,
                  const double *image, const double *tau) {
  FILE *f = fopen(file.c_str(), "w");
  if (!f) stop(13, "cannot open " + file);
  auto a80 = [](const std::string &t) { return (t + std::string(80, ' ')).substr(0, 80); };
  std::string out;
  out.reserve((size_t)nfr * nx * ny * 27 + 4096);
  out += "\n";
  out += fmt_i(1, 2) + "\n";
  out += a80(molname) + "\n" + a80(molfile) + "\n";
  // (57.2957795132 is a REAL literal in the reference, telescope.F:1982)
  out += fmt_es(dist_cm / 3.08572e18, 12, 4) + fmt_es(radvelo, 12, 4) + fmt_f(anginf * (double)57.2957795132f, 7, 3) + "\n";
  out += fmt_i(lev_up, 5) + fmt_i(lev_down, 5) + "\n";
  out += fmt_es(linefreq, 12, 4) + "\n";
  out += fmt_i(nfr, 5) + "\n";
  out += fmt_i(nx, 5) + fmt_i(ny, 5) + fmt_es(spx, 12, 4) + fmt_es(spy, 12, 4) + fmt_f(phioff, 7, 3) +
         fmt_es(xoff, 12, 4) + fmt_es(yoff, 12, 4) + "\n";
  out += "  \n";
  for (int k = 0; k < nfr; k++) {  // telescope.F:2002-2007
    double v = (linefreq + velo[k] * linefreq) - linefreq;
    v = -2.99792458e5 * v / linefreq;
    v = v + radvelo;
    out += fmt_e(v, 10, 3) + " \n";
  }
  out += "  \n";
  const double conv = 3.25465503368e36 / (linefreq * linefreq);
  for (int k = 0; k < nfr; k++) {  // telescope.F:2017-2027: channel, then iy, then ix
    for (int iy = 0; iy < ny; iy++)
      for (int ix = 0; ix < nx; ix++) {
        const size_t i = ((size_t)ix * ny + iy) * nfr + k;
        double temp = conv;
        temp = temp * image[i];
        out += fmt_es(temp, 12, 5) + " " + fmt_es(tau[i], 12, 5) + " \n";
      }
    out += "  \n";
  }
  fwrite(out.data(), 1, out.size(), f);
  fclose(f);
}

void write_imcir(const std::string &file, int nfr, double nu0, int nphi, int nrr, const double *imcir_ri,
                 const double *rays_r, const double *velo, const double *image, const int *cmask) {
  FILE *f = fopen(file.c_str(), "w");
  if (!f) stop(13, "cannot open " + file);
  std::string out;
  out.reserve((size_t)nfr * nphi * nrr * 14 + 4096);
  out += fmt_list_int(nfr) + "\n";
  out += fmt_list_real(nu0) + "\n";
  out += fmt_list_int(nphi) + fmt_list_int(nrr) + "\n";
  for (int ir = 1; ir <= nrr + 1; ir++) out += fmt_e(imcir_ri[ir], 12, 6) + "\n";
  for (int ir = 1; ir <= nrr + 1; ir++) out += fmt_e(ir <= nrr ? rays_r[ir] : 0.0, 12, 6) + "\n";
  for (int inu = 0; inu < nfr; inu++) {
    out += "        \n";
    out += fmt_list_real(velo[inu] * 2.99792458e5) + "\n";
    out += "        \n";
    out += fmt_list_real(image[inu]) + "\n";  // imcir_int(inu,1,0): the central beam
    out += "        \n";
    for (int ip = 0; ip < nphi; ip++)
      for (int ir = 1; ir <= nrr; ir++) {
        const size_t k = ((size_t)ir * nphi + ip) * nfr + inu;
        out += fmt_e(image[k], 10, 4);
        out += "  ";
        out += (char)('0' + (cmask ? (cmask[k] ? 1 : 0) : 0));
        out += '\n';
      }
  }
  fwrite(out.data(), 1, out.size(), f);
  fclose(f);
}

// ---- dump for the tests --------------------------------------------------------------------------
namespace {
void put(FILE *f, const char *name, char dtype, const std::vector<long> &dims, const void *data, size_t elsize) {
  char nm[32] = {0};
  strncpy(nm, name, 31);
  fwrite(nm, 1, 32, f);
  fwrite(&dtype, 1, 1, f);
  const int nd = (int)dims.size();
  fwrite(&nd, sizeof nd, 1, f);
  size_t n = 1;
  for (long d : dims) {
    const long long dd = d;
    fwrite(&dd, sizeof dd, 1, f);
    n *= (size_t)d;
  }
  fwrite(data, elsize, n, f);
}
void putd(FILE *f, const char *name, const std::vector<double> &v, std::vector<long> dims = {}) {
  if (dims.empty()) dims = {(long)v.size()};
  put(f, name, 'd', dims, v.data(), sizeof(double));
}
void puti(FILE *f, const char *name, const std::vector<int> &v) { put(f, name, 'i', {(long)v.size()}, v.data(), sizeof(int)); }
}  // namespace

void dump_workdir(const WorkDir &w, const std::string &file) {
  FILE *f = fopen(file.c_str(), "wb");
  if (!f) stop(13, "cannot open " + file);
  const long nr = (long)w.r.size(), nth = (long)w.theta.size(), ncf = (long)w.cont_freq.size();
  putd(f, "r", w.r);
  putd(f, "theta", w.theta);
  putd(f, "cont_freq_nu", w.cont_freq);
  puti(f, "nsize", w.nsize);
  putd(f, "kappa_abs", w.kappa_abs, {w.nspec, w.maxsize, ncf});
  putd(f, "kappa_scat", w.kappa_scat, {w.nspec, w.maxsize, ncf});
  putd(f, "dust_rho", w.dust_rho, {nr, nth, w.nspec});
  putd(f, "dust_temp", w.dust_temp, {nr, nth, w.nspec, w.maxsize});
  if (!w.scat.empty()) putd(f, "scati_src", w.scat, {nr, nth, ncf});
  putd(f, "rho", w.rho, {nr, nth});
  putd(f, "abund", w.abund, {nr, nth});
  putd(f, "vel", w.vel, {nr, nth, 3});
  putd(f, "linewidth", w.linewidth, {nr, nth});
  putd(f, "starspec_cont", w.starspec);
  if (!w.isrf.empty()) putd(f, "isrf_cont", w.isrf);
  putd(f, "ener_cm", w.ener_cm);
  putd(f, "gdeg", std::vector<double>(w.gdeg.begin(), w.gdeg.begin() + w.nlevels));
  putd(f, "aud", w.aud);
  putd(f, "linefreq", w.linefreq);
  puti(f, "lev_up", w.lev_up);
  puti(f, "lev_down", w.lev_down);
  putd(f, "popul", w.popul, {nr, nth, w.nlevels});
  putd(f, "scalars",
       {w.umass_av, w.rstar, (double)w.out_itype, (double)w.in_itype, (double)w.nphi, (double)w.dbdr,
        (double)w.nrext, w.incl_deg, w.radvelo, (double)w.nfr, w.passband, w.dist_cm, (double)w.nlines_render,
        (double)w.ilinestart, (double)w.command, (double)w.style, w.umass_molec, w.vmax, w.dv});
  fclose(f);
}

}  // namespace rlio

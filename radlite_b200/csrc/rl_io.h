// rl_io.h -- host-side file formats of RADLite (plain C++, no CUDA): the readers the Fortran host runs
// before the line loop and the writers it runs after it, for the stand-alone host program
// radlite_b200_host.  Formats are frozen by the drivers (pyradlite, PRO/*.pro);
// every routine cites the reference reader / writer it mirrors.  Nothing here touches the hot path.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

namespace rlio {

// a reference `stop <code>` reached in host I/O
struct Stop {
  int code;
  std::string msg;
};

// Fortran sequential formatted file read the way the reference reads it: list-directed READs
// (`read(u,*) a,b,c`: blanks / tabs / commas separate, a statement starts on a fresh record, runs on
// over records until its list is satisfied and discards the rest of the last one; `d` exponents;
// r*c repeats) and fixed-column READs (`read(u,'(I5,F12.4,F7.1)')`).
class FortranReader {
 public:
  explicit FortranReader(const std::string &path);
  bool is_open() const { return ok_; }
  const std::string &path() const { return path_; }
  // one list-directed READ of n numeric items
  std::vector<double> reals(size_t n);
  double real() { return reals(1)[0]; }
  long integer();
  std::vector<long> integers(size_t n);
  // one list-directed READ of a character item (first token of the next non-blank record)
  std::string word();
  // one formatted READ: the next record as it is (for fixed-column fields)
  std::string record();

 private:
  bool next_token(std::string &tok);
  void end_statement();
  std::string path_, buf_;
  size_t pos_ = 0;
  bool ok_ = false;
  bool at_record_start_ = true;
  long repeat_ = 0;
  std::string repeat_tok_;
};

// fixed-column field of a formatted record (1-based column, width); blank = 0
double field_real(const std::string &rec, int col, int width);
long field_int(const std::string &rec, int col, int width);

// everything the Fortran host holds in COMMON before the line loop (main.F:79-1043)
struct WorkDir {
  // radlite.inp (main.F:488-686, setup.F:1117-1137)
  int input_format = 0, out_itype = 0, in_itype = 2, nphi = 0, dbdr = 1, nrext = 0;
  int do_dust = 0, do_lines = 1, dust_in_lines = 1, star_pump = 1;
  // grids (grid.F:711-727, 1098-1118, 1451-1470)
  std::vector<double> r, theta, cont_freq;
  // dust (setup.F:681, 159; dust.F:92, 317; source.F:494)
  int nspec = 0, maxsize = 0;
  std::vector<int> nsize;
  std::vector<double> kappa_abs, kappa_scat;  // [nspec][maxsize][ncf]
  std::vector<double> dust_rho;               // [nr][nth][nspec]
  std::vector<double> dust_temp;              // [nr][nth][nspec][maxsize]
  std::vector<double> scat;                   // [nr][nth][ncf] or empty
  // gas (setup.F:1453, 864, 754, 803; line.F:3918)
  std::vector<double> rho, abund, vel, linewidth;
  double umass_av = 0.0;
  // star / outer boundary (star.F:576, 449, 675)
  double rstar = 0.0;
  std::vector<double> starspec, isrf;  // surface intensity on cont_freq ; ISRF (out_itype 3)
  // linespectrum.inp (telescope.F:86-238) and main.F:202-213
  int style = 0, command = 0, nlines_render = 0, ilinestart = 1;
  // imager block of command 2 (telescope.F:204-216): pixel counts, pixel sizes [cm], rotation, offsets, star switch
  int imr_nx = 0, imr_ny = 0, imrec_addstar = 0;
  double imr_spx = 0.0, imr_spy = 0.0, imr_phioff = 0.0, imr_xoff = 0.0, imr_yoff = 0.0;
  double vmax = 0.0, dv = 0.0, incl_deg = 0.0, radvelo = 0.0;
  std::string molfile, molname;
  int nfr = 0;
  double passband = 0.0, dist_cm = 3.08572e18;
  // molecule (line.F:1826-1985) and level populations (line.F:811-950)
  double umass_molec = 0.0;
  int nlev_orig = 0, nlevels = 0, nlines = 0;
  std::vector<double> ener_cm, gdeg, aud, linefreq;
  std::vector<int> lev_up, lev_down;
  std::vector<double> popul;  // [nr][nth][nlevels]
  // line.inp (line.F:142-260)
  double rangewidth1 = 0.0;
};

// read a RADLite working directory (the current directory) in the reference's order; throws Stop
WorkDir read_workdir();

// ---- gfortran-compatible output ---------------------------------------------------------------
std::string fmt_e(double v, int w, int d);  // Ew.d
std::string fmt_f(double v, int w, int d);  // Fw.d
std::string fmt_i(long v, int w);           // Iw
std::string fmt_es(double v, int w, int d); // ESw.d
std::string fmt_list_real(double v);        // write(u,*) of a doubleprecision
std::string fmt_list_int(long v);           // write(u,*) of an integer

// telescope.F:1662-1681 header_line_spectrum
void write_spectrum_header(const std::string &file, const std::string &molname, const std::string &molfile,
                           int nlinespec, int nfrmax, double dist_cm, double radvelo, double anginf,
                           int iformout);
// telescope.F:1703-1803 write_line_spectrum (iformout = 1, linespec_flag1 = 1): appends one line
void append_line_spectrum(const std::string &file, int lev_up, int lev_down, double linefreq, int nfr,
                          const double *velo /*line_dnu/nu0*/, const double *flux, double radvelo);
// telescope.F:1595-1621 + 1346-1371 (SAVE_IMCIR): lineposvelcirc_<mol>_<iline>.dat
void write_imcir(const std::string &file, int nfr, double nu0, int nphi, int nrr, const double *imcir_ri,
                 const double *rays_r, const double *velo, const double *image /*[nrr+1][nphi][nfr]*/,
                 const int *cmask);
std::string imcir_filename(const std::string &molname, int iline);
// lineposvel_<mol>_<iline>.dat: telescope.F:1934-2040 (calc_write_line_posvel).  image, tau: [nx][ny][nfr];
// velo[nfr] = line_dnu / nu0
void write_posvel(const std::string &file, const std::string &molname, const std::string &molfile, double dist_cm,
                  double radvelo, double anginf, int lev_up, int lev_down, double linefreq, int nfr, int nx, int ny,
                  double spx, double spy, double phioff, double xoff, double yoff, const double *velo,
                  const double *image, const double *tau);
std::string posvel_filename(const std::string &molname, int iline);

// binary dump of the parsed model for the tests: records {name, dtype 'd'|'i', ndim, dims, data}
void dump_workdir(const WorkDir &w, const std::string &file);

}  // namespace rlio

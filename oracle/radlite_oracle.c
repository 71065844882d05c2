/*
 * radlite_oracle.c -- CPU restatement (plain C99) of RADLite's line ray-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see radlite_oracle.h).  PARITY UNPINNED: the Fortran reference
 * cannot be built in this image and ships no golden vectors; every routine below restates the
 * cited reference lines (paths relative to /root/reference/RADLITE), keeping the reference's
 * literals, branch order, single-precision quirks and redundancy (geometry rebuilt per line,
 * node interpolation redone per channel, Einstein B's redone per line).
 *
 * Compile with -O2 -ffp-contract=off (no FMA contraction: the reference Makefile:23 builds for
 * baseline x86-64, i.e. separate multiply and add).
 */
#include "radlite_oracle.h"

#include <math.h>
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- reference compile-time constants ------------------------------------------------- */
#define TELESC_EPS 1.0e-10               /* configure.h:46 */
#define PICONST 3.1415926535897932385    /* main.h:452 */
#define TEMPCMB 2.728                    /* main.h:453 */
#define RAYEXPT 100                      /* common_telescope.h:4 */
#define RAYADPT 4                        /* common_telescope.h:5 */
#define RAYRNPT 4                        /* common_telescope.h:6 */
#define LG_NRMAX 31                      /* line.F:4657-4661: 2*3.0*5.+1 */

struct orc_ctx {
  char err[256];
  jmp_buf jb;
  int jb_armed;
  /* grid (common_grid.h:17 rsi_x_c(-1:FRSIZE_MAX+2,1:2), ridx_it(-4:nt+4)) */
  int nr, nt, nth;
  double *rc, *tc; /* rc[i+1] = rsi_x_c(i,1), i=-1..nr+2 ; tc[i+1] = rsi_x_c(i,2), i=-1..nt+2 */
  int *ridx;       /* ridx[i+4] = ridx_it(i), i=-4..nt+4 */
  /* medium */
  double *rho, *abund, *vel, *lw;
  double umass_av;
  /* lines */
  int nlines, nlevels;
  int *lev_up, *lev_down;
  double *linefreq, *aud, *gdeg, *popul;
  double *bud, *bdu;
  /* per-line dust: [iline][ir][it] */
  double *ld_src, *ld_alp;
  int have_line_dust;
  /* dust inputs */
  int nspec, maxsize, ncf_d;
  int *nsize;
  double *cfreq_d, *kabs, *kscat, *drho, *dtemp, *scat;
  int have_dust;
  /* camera (common_telescope.h:12-33) */
  int cam_set;
  /* rectangular camera (telescope.F:2229-2475): rays 1..nrx*nry, then the optional central ray */
  int rect_set, rays_nrx, rays_nry, imrec_addstar, imrec_starunres;
  double rays_sizepix_x, rays_sizepix_y;
  double *rp_b;
  double anginf;
  int nphi, nrext, dbdr, imethod, nrref;
  double rstar;
  int rays_amount, rays_nrr, rays_nrphi, rp_nrrayextra, rp_nrref, rp_dbdr;
  double *rp_x0, *rp_z0, *rp_theta0, *rp_s0; /* index 0..rays_amount */
  double *rays_r;                            /* 0..nrr */
  double *imcir_r, *imcir_ri;                /* 0..nrr+1 */
  float *minvel, *maxvel;                    /* REAL*4, common_telescope.h:17 */
  int *cmask_persist;                        /* imcir_cmask is never cleared (telescope.F:548,575) */
  int cmask_nfr;
  /* boundary */
  int in_itype, out_itype, ncf_b;
  double *cfreq_b, *starspec_cont, *isrf_cont;
  int bc_set;
  /* options */
  int subgrid, nonredundant;
  double levthres, aksmax_opt;
  /* per-line state (common_lines.h) */
  double *line_dnu;     /* 1..nfr */
  double *freq_nu;      /* 1..nfr */
  double *starspec_line, *isrf_line;
  int nfr;
  double line_nu0;
  int maserflag;
  double char_tau, char_emis;
  /* trajectory (common_telescope.h /onetraject/) */
  int raysize;
  int tr_amount;
  double *tr_s, *tr_radius, *tr_theta, *tr_mu, *tr_phi;
  int *tr_icross, *tr_iradius, *tr_itheta;
  double tr_b;
  /* scratch for make_trajectory_c */
  double *th_radius, *th_theta, *th_s, *r_radius, *r_theta, *r_s, *sar1, *sar2;
  int *th_ir, *th_itheta, *r_ir, *r_itheta, *iyar;
  double *rrgrid, *ttgrid;
  /* counters */
  double cnt_R, cnt_E, cnt_S;
  /* bounded-sample benchmarking: only rings ring_lo..ring_hi are traced (0,0 = all) */
  int ring_lo, ring_hi, ring_stride;
};

#define RC(i) (c->rc[(i) + 1])
#define TC(i) (c->tc[(i) + 1])
#define RIDX(i) (c->ridx[(i) + 4])
#define CELL(it, ir) ((size_t)((ir)-1) * (size_t)c->nth + (size_t)((it)-1))

static void stop_(orc_ctx *c, int code, const char *msg) {
  snprintf(c->err, sizeof c->err, "stop %d: %s", code, msg);
  if (c->jb_armed) longjmp(c->jb, code ? code : 1);
}
#define STOP(code, msg) stop_(c, (code), (msg))

static void *xcalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) {
    fprintf(stderr, "oracle: out of memory\n");
    abort();
  }
  return p;
}
static double *dupd(const double *a, size_t n) {
  double *p = (double *)xcalloc(n, sizeof(double));
  if (a) memcpy(p, a, n * sizeof(double));
  return p;
}
static int *dupi(const int *a, size_t n) {
  int *p = (int *)xcalloc(n, sizeof(int));
  if (a) memcpy(p, a, n * sizeof(int));
  return p;
}

/* ---- nrecip.F:157-204 hunt ; xx is 1-based xx[1..n] ----------------------------------- */
static void hunt1(const double *xx, int n, double x, int *jlo_io) {
  int jlo = *jlo_io, jhi, jm, inc;
  int ascnd = xx[n] > xx[1];
  if (jlo <= 0 || jlo > n) {
    jlo = 0;
    jhi = n + 1;
    goto l3;
  }
  inc = 1;
  if ((x >= xx[jlo]) == ascnd) {
  l1:
    jhi = jlo + inc;
    if (jhi > n) {
      jhi = n + 1;
    } else if ((x >= xx[jhi]) == ascnd) {
      jlo = jhi;
      inc = inc + inc;
      goto l1;
    }
  } else {
    jhi = jlo;
  l2:
    jlo = jhi - inc;
    if (jlo < 1) {
      jlo = 0;
    } else if ((x < xx[jlo]) == ascnd) {
      jhi = jlo;
      inc = inc + inc;
      goto l2;
    }
  }
l3:
  while (jhi - jlo != 1) {
    jm = (jhi + jlo) / 2;
    if ((x > xx[jm]) == ascnd)
      jlo = jm;
    else
      jhi = jm;
  }
  *jlo_io = jlo;
}
void orc_hunt(const double *xx, int n, double x, int *jlo) { hunt1(xx - 1, n, x, jlo); }
/* nrecip.F:206-247 hunt0 is the same routine on an array declared xx(0:n); indices coincide */
#define hunt0 hunt1

/* ---- nrecip.F:752-831 indexx, 1-based arrays ------------------------------------------ */
static void indexx(int n, const double *arr, int *indx) {
  enum { M = 7, NSTACK = 50 };
  int i, indxt, ir, itemp, j, jstack, k, l, istack[NSTACK + 1];
  double a;
  for (j = 1; j <= n; j++) indx[j] = j;
  jstack = 0;
  l = 1;
  ir = n;
  for (;;) {
    if (ir - l < M) {
      for (j = l + 1; j <= ir; j++) {
        indxt = indx[j];
        a = arr[indxt];
        for (i = j - 1; i >= 1; i--) {
          if (arr[indx[i]] <= a) goto l2;
          indx[i + 1] = indx[i];
        }
        i = 0;
      l2:
        indx[i + 1] = indxt;
      }
      if (jstack == 0) return;
      ir = istack[jstack];
      l = istack[jstack - 1];
      jstack -= 2;
    } else {
      k = (l + ir) / 2;
      itemp = indx[k];
      indx[k] = indx[l + 1];
      indx[l + 1] = itemp;
      if (arr[indx[l + 1]] > arr[indx[ir]]) {
        itemp = indx[l + 1];
        indx[l + 1] = indx[ir];
        indx[ir] = itemp;
      }
      if (arr[indx[l]] > arr[indx[ir]]) {
        itemp = indx[l];
        indx[l] = indx[ir];
        indx[ir] = itemp;
      }
      if (arr[indx[l + 1]] > arr[indx[l]]) {
        itemp = indx[l + 1];
        indx[l + 1] = indx[l];
        indx[l] = itemp;
      }
      i = l + 1;
      j = ir;
      indxt = indx[l];
      a = arr[indxt];
      for (;;) {
        do i++;
        while (arr[indx[i]] < a);
        do j--;
        while (arr[indx[j]] > a);
        if (j < i) break;
        itemp = indx[i];
        indx[i] = indx[j];
        indx[j] = itemp;
      }
      indx[l] = indx[j];
      indx[j] = indxt;
      jstack += 2;
      if (ir - i + 1 >= j - l) {
        istack[jstack] = ir;
        istack[jstack - 1] = i;
        ir = j - 1;
      } else {
        istack[jstack] = j - 1;
        istack[jstack - 1] = l;
        l = i;
      }
    }
  }
}

/* ---- nrecip.F:833-870 ray_sort (1-based) ---------------------------------------------- */
static void ray_sort(int n, double *ra, double *rb, double *rc_, int *ird, int *ire) {
  double wksp[RAYEXPT + 1];
  int iwksp[RAYEXPT + 1], iiwksp[RAYEXPT + 1], j;
  indexx(n, ra, iwksp);
  for (j = 1; j <= n; j++) wksp[j] = ra[j];
  for (j = 1; j <= n; j++) ra[j] = wksp[iwksp[j]];
  for (j = 1; j <= n; j++) wksp[j] = rb[j];
  for (j = 1; j <= n; j++) rb[j] = wksp[iwksp[j]];
  for (j = 1; j <= n; j++) wksp[j] = rc_[j];
  for (j = 1; j <= n; j++) rc_[j] = wksp[iwksp[j]];
  for (j = 1; j <= n; j++) iiwksp[j] = ird[j];
  for (j = 1; j <= n; j++) ird[j] = iiwksp[iwksp[j]];
  for (j = 1; j <= n; j++) iiwksp[j] = ire[j];
  for (j = 1; j <= n; j++) ire[j] = iiwksp[iwksp[j]];
}

/* ---- setup.F:937-952 bplanck ----------------------------------------------------------- */
double orc_bplanck(double temp, double nu) {
  if (temp == 0.0) return 0.0;
  return 1.47455e-47 * nu * nu * nu / (exp(4.7989e-11 * nu / temp) - 1.0) + 1.e-290;
}

/* ---- transfer.F:1498-1571 qdr_src_2 ---------------------------------------------------- */
double orc_qdr_src_2(double inten, double js1, double alp1, double js2, double alp2, double ds) {
  double e0, e1, a, b, dtau1, xp, src1, src2, theomax, q;
  dtau1 = 0.5 * (alp1 + alp2) * ds;
  theomax = 0.5 * (js1 + js2) * ds;
  if (dtau1 > 1.e-6) {
    xp = exp(-dtau1);
    e0 = 1.0 - xp;
    e1 = dtau1 - e0;
    b = e1 / dtau1;
    a = e0 - b;
  } else {
    a = 0.5 * dtau1;
    b = 0.5 * dtau1;
    xp = 1.0 - dtau1;
  }
  if (alp1 > 0.0)
    src1 = js1 / alp1;
  else if (alp2 > 0.0)
    src1 = js2 / alp2;
  else
    src1 = 0.0;
  if (alp2 > 0.0)
    src2 = js2 / alp2;
  else if (alp1 > 0.0)
    src2 = js1 / alp1;
  else
    src2 = 0.0;
  if (dtau1 > (double)1e-9f) /* transfer.F:1542: REAL literal 1e-9 */
    q = a * src1 + b * src2;
  else
    q = 0.5 * (js1 + js2) * ds;
  q = (q < theomax) ? q : theomax; /* min(q,theomax), transfer.F:1559 */
  return inten * xp + q;
}

/* ======================================================================================= */
int orc_create(orc_ctx **out) {
  orc_ctx *c = (orc_ctx *)xcalloc(1, sizeof *c);
  c->subgrid = 1;
  c->nonredundant = 1;
  c->levthres = 1e-3;
  c->aksmax_opt = -1.0;
  c->in_itype = 2;
  c->out_itype = 0;
  *out = c;
  return 0;
}

static void free_traj(orc_ctx *c) {
  free(c->tr_s); free(c->tr_radius); free(c->tr_theta); free(c->tr_mu); free(c->tr_phi);
  free(c->tr_icross); free(c->tr_iradius); free(c->tr_itheta);
  free(c->th_radius); free(c->th_theta); free(c->th_s); free(c->r_radius); free(c->r_theta);
  free(c->r_s); free(c->sar1); free(c->sar2); free(c->th_ir); free(c->th_itheta);
  free(c->r_ir); free(c->r_itheta); free(c->iyar); free(c->rrgrid); free(c->ttgrid);
  c->tr_s = c->tr_radius = c->tr_theta = c->tr_mu = c->tr_phi = 0;
  c->tr_icross = c->tr_iradius = c->tr_itheta = 0;
  c->th_radius = c->th_theta = c->th_s = c->r_radius = c->r_theta = c->r_s = c->sar1 = c->sar2 = 0;
  c->th_ir = c->th_itheta = c->r_ir = c->r_itheta = c->iyar = 0;
  c->rrgrid = c->ttgrid = 0;
}
static void free_cam(orc_ctx *c) {
  free(c->rp_x0); free(c->rp_z0); free(c->rp_theta0); free(c->rp_s0); free(c->rays_r);
  free(c->imcir_r); free(c->imcir_ri); free(c->minvel); free(c->maxvel); free(c->cmask_persist);
  c->rp_x0 = c->rp_z0 = c->rp_theta0 = c->rp_s0 = c->rays_r = c->imcir_r = c->imcir_ri = 0;
  c->minvel = c->maxvel = 0;
  c->cmask_persist = 0; c->cam_set = 0;
}
void orc_destroy(orc_ctx *c) {
  if (!c) return;
  free(c->rc); free(c->tc); free(c->ridx);
  free(c->rho); free(c->abund); free(c->vel); free(c->lw);
  free(c->lev_up); free(c->lev_down); free(c->linefreq); free(c->aud); free(c->gdeg);
  free(c->popul); free(c->bud); free(c->bdu); free(c->ld_src); free(c->ld_alp);
  free(c->nsize); free(c->cfreq_d); free(c->kabs); free(c->kscat); free(c->drho);
  free(c->dtemp); free(c->scat);
  free(c->cfreq_b); free(c->starspec_cont); free(c->isrf_cont);
  free(c->line_dnu); free(c->freq_nu); free(c->starspec_line); free(c->isrf_line);
  free_cam(c);
  free_traj(c);
  free(c);
}
const char *orc_last_error(const orc_ctx *c) { return c ? c->err : "null ctx"; }

/* interpol.F:87-100 make_index -> ridx_it (MIRROR_THETA) */
static void make_index(orc_ctx *c) {
  int it, nt = c->nt;
  free(c->ridx);
  c->ridx = (int *)xcalloc((size_t)nt + 9, sizeof(int));
  for (it = -4; it <= nt + 4; it++) {
    int v = it;
    if (v < 1) v = 1 - v;
    if (v > nt) v = 2 * nt + 1 - v;
    if (v > nt / 2) v = nt + 1 - v;
    RIDX(it) = v;
  }
}

static void alloc_traj(orc_ctx *c) {
  /* common_telescope.h:7 RAYSIZE = 2*(2*FRSIZE_X+FRSIZE_Y+RAYEXPT), with run-time sizes */
  int n = 2 * (2 * c->nr + c->nt + RAYEXPT) + 8;
  free_traj(c);
  c->raysize = n;
#define AD(p) c->p = (double *)xcalloc((size_t)n + 2, sizeof(double))
#define AI(p) c->p = (int *)xcalloc((size_t)n + 2, sizeof(int))
  AD(tr_s); AD(tr_radius); AD(tr_theta); AD(tr_mu); AD(tr_phi);
  AI(tr_icross); AI(tr_iradius); AI(tr_itheta);
  AD(th_radius); AD(th_theta); AD(th_s); AD(r_radius); AD(r_theta); AD(r_s); AD(sar1); AD(sar2);
  AI(th_ir); AI(th_itheta); AI(r_ir); AI(r_itheta); AI(iyar);
#undef AD
#undef AI
  c->rrgrid = (double *)xcalloc((size_t)c->nr + 3, sizeof(double));
  c->ttgrid = (double *)xcalloc((size_t)c->nt + 3, sizeof(double));
}

int orc_set_grid_ghosted(orc_ctx *c, int nr, int nt, const double *rc_m1, const double *tc_m1) {
  if (nr < 2 || nt < 2 || (nt & 1)) {
    snprintf(c->err, sizeof c->err, "set_grid: bad sizes nr=%d nt=%d", nr, nt);
    return 13;
  }
  free(c->rc);
  free(c->tc);
  c->nr = nr;
  c->nt = nt;
  c->nth = nt / 2;
  c->rc = dupd(rc_m1, (size_t)nr + 4);
  c->tc = dupd(tc_m1, (size_t)nt + 4);
  make_index(c);
  alloc_traj(c);
  free_cam(c);
  return 0;
}

int orc_set_grid(orc_ctx *c, int nr, int nth, const double *r, const double *theta) {
  int nt = 2 * nth, i, rcode;
  double *rc = (double *)xcalloc((size_t)nr + 4, sizeof(double));
  double *tc = (double *)xcalloc((size_t)nt + 4, sizeof(double));
#define R_(i) rc[(i) + 1]
#define T_(i) tc[(i) + 1]
  for (i = 1; i <= nr; i++) R_(i) = r[i - 1];
  /* grid.F:772-777 */
  R_(0) = R_(1) * R_(1) / R_(2);
  R_(-1) = R_(0) * R_(0) / R_(1);
  R_(nr + 1) = R_(nr) * R_(nr) / R_(nr - 1);
  R_(nr + 2) = R_(nr + 1) * R_(nr + 1) / R_(nr);
  /* grid.F:1147-1176 */
  for (i = 1; i <= nth; i++) T_(i) = theta[i - 1];
  for (i = 1; i <= nth; i++) T_(2 * nth + 1 - i) = 3.14159265359 - T_(i);
  T_(0) = -T_(1);
  T_(-1) = -T_(2);
  {
    float twopi_f = 2 * 3.1415926e0f; /* grid.F:1173: REAL arithmetic */
    T_(nt + 1) = (double)twopi_f - T_(nt);
    T_(nt + 2) = (double)twopi_f - T_(nt - 1);
  }
#undef R_
#undef T_
  rcode = orc_set_grid_ghosted(c, nr, nt, rc, tc);
  free(rc);
  free(tc);
  return rcode;
}

int orc_set_medium(orc_ctx *c, const double *rho, const double *abund, const double *vel,
                   const double *linewidth, double umass_av) {
  size_t n = (size_t)c->nr * c->nth;
  free(c->rho); free(c->abund); free(c->vel); free(c->lw);
  c->rho = dupd(rho, n);
  c->abund = dupd(abund, n);
  c->vel = dupd(vel, 3 * n);
  c->lw = dupd(linewidth, n);
  c->umass_av = umass_av;
  return 0;
}

/* line.F:1708-1788 prepare_lines: Einstein B's */
static void prepare_lines(orc_ctx *c) {
  int i;
  for (i = 0; i < c->nlines; i++) {
    double gratio = c->gdeg[c->lev_up[i] - 1] / c->gdeg[c->lev_down[i] - 1];
    c->bud[i] = 6.78171833781e46 * c->aud[i] / (c->linefreq[i] * c->linefreq[i] * c->linefreq[i]);
    c->bdu[i] = c->bud[i] * gratio;
  }
}

int orc_set_lines(orc_ctx *c, int nlines, int nlevels, const int *lev_up, const int *lev_down,
                  const double *linefreq, const double *aud, const double *gdeg,
                  const double *popul) {
  int i;
  if (nlines < 1) { snprintf(c->err, sizeof c->err, "stop 13: Minimum of 1 line!"); return 13; }
  if (nlevels < 2) { snprintf(c->err, sizeof c->err, "stop 13: Minimum of 2 levels!"); return 13; }
  for (i = 0; i < nlines; i++) {
    if (lev_up[i] < 1 || lev_down[i] < 1 || lev_up[i] > nlevels || lev_down[i] > nlevels) {
      snprintf(c->err, sizeof c->err, "stop 13: line %d levels out of range", i + 1);
      return 13;
    }
    if (lev_up[i] <= lev_down[i]) {
      snprintf(c->err, sizeof c->err, "stop 13: line %d not upper->lower", i + 1);
      return 13;
    }
    if (linefreq[i] == 0.0) { snprintf(c->err, sizeof c->err, "stop 13: linefreq 0"); return 13; }
  }
  free(c->lev_up); free(c->lev_down); free(c->linefreq); free(c->aud); free(c->gdeg);
  free(c->popul); free(c->bud); free(c->bdu);
  c->nlines = nlines;
  c->nlevels = nlevels;
  c->lev_up = dupi(lev_up, nlines);
  c->lev_down = dupi(lev_down, nlines);
  c->linefreq = dupd(linefreq, nlines);
  c->aud = dupd(aud, nlines);
  c->gdeg = dupd(gdeg, nlevels);
  c->popul = dupd(popul, (size_t)c->nr * c->nth * nlevels);
  c->bud = (double *)xcalloc(nlines, sizeof(double));
  c->bdu = (double *)xcalloc(nlines, sizeof(double));
  prepare_lines(c);
  free(c->ld_src); free(c->ld_alp);
  c->ld_src = (double *)xcalloc((size_t)nlines * c->nr * c->nth, sizeof(double));
  c->ld_alp = (double *)xcalloc((size_t)nlines * c->nr * c->nth, sizeof(double));
  c->have_line_dust = 0;
  return 0;
}

/* line.F:3502-3608 line_dust_compute_src_alp + 3687-3743 global_prepare_line_dust */
static void global_prepare_line_dust(orc_ctx *c) {
  int ir, it, iline, ispec, isize, inudust = 0;
  int ncf = c->ncf_d;
  const double *cf = c->cfreq_d - 1; /* 1-based */
  for (ir = 1; ir <= c->nr; ir++)
    for (it = 1; it <= c->nth; it++)
      for (iline = 1; iline <= c->nlines; iline++) {
        double src, alp, freq = c->linefreq[iline - 1];
        size_t cell = CELL(it, ir);
        hunt1(cf, ncf, freq, &inudust);
        if (inudust == 0 || inudust == ncf) {
          src = 0.0;
          alp = 0.0;
        } else {
          double wgt = (freq - cf[inudust]) / (cf[inudust + 1] - cf[inudust]);
          src = 0.0;
          alp = 0.0;
          for (ispec = 0; ispec < c->nspec; ispec++)
            for (isize = 0; isize < c->nsize[ispec]; isize++) {
              const double *ka = c->kabs + ((size_t)ispec * c->maxsize + isize) * ncf - 1;
              double kappawgt = wgt * ka[inudust + 1] + (1.0 - wgt) * ka[inudust];
              double rho = c->drho[cell * c->nspec + ispec];
              double temp = c->dtemp[(cell * c->nspec + ispec) * c->maxsize + isize];
              src = src + rho * kappawgt * orc_bplanck(temp, freq);
              alp = alp + rho * kappawgt;
            }
          if (c->scat) {
            const double *sc = c->scat + cell * ncf - 1;
            src = src + wgt * sc[inudust + 1] + (1.0 - wgt) * sc[inudust];
          } else {
            src = src + wgt * 0.0 + (1.0 - wgt) * 0.0;
          }
          for (ispec = 0; ispec < c->nspec; ispec++)
            for (isize = 0; isize < c->nsize[ispec]; isize++) {
              const double *ks = c->kscat + ((size_t)ispec * c->maxsize + isize) * ncf - 1;
              double kappawgt = wgt * ks[inudust + 1] + (1.0 - wgt) * ks[inudust];
              double rho = c->drho[cell * c->nspec + ispec];
              alp = alp + rho * kappawgt;
            }
        }
        c->ld_src[(size_t)(iline - 1) * c->nr * c->nth + cell] = src;
        c->ld_alp[(size_t)(iline - 1) * c->nr * c->nth + cell] = alp;
      }
  c->have_line_dust = 1;
}

int orc_set_dust(orc_ctx *c, int nspec, const int *nsize, int ncf, const double *cont_freq_nu,
                 const double *kappa_abs, const double *kappa_scat, const double *dust_rho,
                 const double *dust_temp, const double *scati_src) {
  int i, maxsize = 0;
  size_t ncell = (size_t)c->nr * c->nth;
  if (!c->nlines) { snprintf(c->err, sizeof c->err, "set_dust: call set_lines first"); return 13; }
  for (i = 0; i < nspec; i++)
    if (nsize[i] > maxsize) maxsize = nsize[i];
  free(c->nsize); free(c->cfreq_d); free(c->kabs); free(c->kscat); free(c->drho);
  free(c->dtemp); free(c->scat);
  c->nspec = nspec;
  c->maxsize = maxsize;
  c->ncf_d = ncf;
  c->nsize = dupi(nsize, nspec);
  c->cfreq_d = dupd(cont_freq_nu, ncf);
  c->kabs = dupd(kappa_abs, (size_t)nspec * maxsize * ncf);
  c->kscat = dupd(kappa_scat, (size_t)nspec * maxsize * ncf);
  c->drho = dupd(dust_rho, ncell * nspec);
  c->dtemp = dupd(dust_temp, ncell * nspec * maxsize);
  c->scat = scati_src ? dupd(scati_src, ncell * ncf) : 0;
  c->have_dust = 1;
  global_prepare_line_dust(c);
  return 0;
}

int orc_set_line_dust(orc_ctx *c, const double *src, const double *alp) {
  size_t n = (size_t)c->nlines * c->nr * c->nth;
  if (!c->nlines) { snprintf(c->err, sizeof c->err, "set_line_dust: call set_lines first"); return 13; }
  memcpy(c->ld_src, src, n * sizeof(double));
  memcpy(c->ld_alp, alp, n * sizeof(double));
  c->have_line_dust = 1;
  return 0;
}

int orc_set_bc(orc_ctx *c, int in_itype, int out_itype, int ncf, const double *cont_freq_nu,
               const double *starspec_cont, const double *isrf_cont) {
  free(c->cfreq_b); free(c->starspec_cont); free(c->isrf_cont);
  c->in_itype = in_itype;
  c->out_itype = out_itype;
  c->ncf_b = ncf;
  c->cfreq_b = dupd(cont_freq_nu, ncf);
  c->starspec_cont = dupd(starspec_cont, ncf);
  c->isrf_cont = isrf_cont ? dupd(isrf_cont, ncf) : 0;
  c->bc_set = 1;
  return 0;
}

int orc_set_options(orc_ctx *c, int subgrid, int nonredundant, double levthres, double aksmax) {
  c->subgrid = subgrid;
  c->nonredundant = nonredundant;
  c->levthres = levthres;
  c->aksmax_opt = aksmax;
  return 0;
}

/* ---- telescope.F:715-1191 setup_rays_circular ------------------------------------------ */
static int setup_rays_circular(orc_ctx *c) {
  const double epsxyz = 1.0e2 * TELESC_EPS, epsrrr = 1.0e3 * TELESC_EPS;
  int irmin = 1, irmax = c->nr, nrphiinf = c->nphi, nrext = c->nrext, dbdr = c->dbdr;
  int imethod = c->imethod, nrref = c->nrref;
  int nrrextra, ix, iys, ir, iradius, iins;
  double anginf = c->anginf, rstar = c->rstar;
  double theta0, sinth0, dphi, phi, r_c = 0.0, zh02, dum, dr, refdum1, refdum2;
  double *zhat0, *xhat0;
  size_t maxrays, maxrings;

  if (imethod < 0) STOP(1, "Negative imethod not allowed.");
  nrrextra = nrext;
  if (nrrextra < 0) nrrextra = -nrrextra;
  if (imethod == 0) {
    if (nrext > 0) { nrrextra = nrext; imethod = -1; }
    else if (nrext < 0) { nrrextra = -nrext; imethod = -2; }
    else STOP(1, "Must have non-zero nrrextra");
  }
  if (irmax <= irmin) STOP(13, "ERROR Telescope: irmax.le.irmin");
  if (fabs(anginf) < 1.e-1) anginf = 0.1 * fabs(anginf) / anginf; /* telescope.F:818-827 */

  maxrings = (size_t)nrrextra + nrref + (size_t)(c->nr) * dbdr + 4;
  maxrays = 2 + maxrings * (size_t)nrphiinf;
  free_cam(c);
  c->rp_x0 = (double *)xcalloc(maxrays, sizeof(double));
  c->rp_z0 = (double *)xcalloc(maxrays, sizeof(double));
  c->rp_theta0 = (double *)xcalloc(maxrays, sizeof(double));
  c->rp_s0 = (double *)xcalloc(maxrays, sizeof(double));
  c->rays_r = (double *)xcalloc(maxrings + 2, sizeof(double));
  zhat0 = (double *)xcalloc((size_t)nrphiinf + 1, sizeof(double));
  xhat0 = (double *)xcalloc((size_t)nrphiinf + 1, sizeof(double));

  theta0 = anginf + 1.e-4; /* telescope.F:863 */
  sinth0 = sin(theta0);
  c->rays_nrphi = nrphiinf;
  dphi = 6.28318530718 / (1.0 * nrphiinf);
  phi = 0.5 * dphi;
  for (iys = 1; iys <= nrphiinf; iys++) { /* telescope.F:887-903 */
    zhat0[iys] = -sin(phi) / sinth0;
    zh02 = zhat0[iys] * zhat0[iys];
    dum = 1.0 - zh02 * sinth0 * sinth0;
    dum = dum + epsxyz;
    if (dum >= 0.0) {
      if (cos(phi) > 0.0) xhat0[iys] = sqrt(dum);
      else xhat0[iys] = -sqrt(dum);
    } else {
      free(zhat0); free(xhat0);
      STOP(13, "ERROR in setup_rays_circular");
    }
    phi = phi + dphi;
  }
  ir = 1;
  iradius = 0;
  c->rp_x0[ir] = 0.0;
  c->rp_z0[ir] = 0.0;
  c->rp_theta0[ir] = theta0;
  c->rp_s0[ir] = 1.e30;
  c->rays_r[iradius] = 0.0;
  ir = 2;
  iradius = 1;
#define ADD_RING()                                                      \
  do {                                                                  \
    c->rays_r[iradius] = r_c;                                           \
    for (iys = 1; iys <= nrphiinf; iys++) {                             \
      c->rp_x0[ir] = r_c * xhat0[iys];                                  \
      c->rp_z0[ir] = r_c * zhat0[iys];                                  \
      c->rp_theta0[ir] = theta0;                                        \
      c->rp_s0[ir] = 1.e30;                                             \
      ir = ir + 1;                                                      \
    }                                                                   \
    iradius = iradius + 1;                                              \
  } while (0)
  if (imethod < 0) { /* telescope.F:930-985 */
    for (ix = 1; ix <= nrrextra; ix++) {
      if (imethod == -2) {
        if (rstar > RC(1)) { free(zhat0); free(xhat0); STOP(83991, "rstar > R(1)"); }
        r_c = (ix * (RC(1) - rstar) / (nrrextra + 1.0)) + rstar;
      } else {
        r_c = ix * RC(1) / (nrrextra + 1.0);
      }
      ADD_RING();
    }
  } else { /* imethod > 0 : telescope.F:986-1056 */
    if (nrref <= 0) { free(zhat0); free(xhat0); STOP(1, "nrref must be set>0"); }
    refdum1 = 0.5;
    refdum2 = 0.0;
    for (ix = 1; ix <= nrrextra + nrref; ix++) {
      if (imethod == 1) {
        if (rstar > RC(1)) { free(zhat0); free(xhat0); STOP(91991, "rstar > R(1)"); }
        if (ix <= nrrextra) {
          r_c = ((ix - 1) * (RC(1) - rstar) / nrrextra) + rstar;
        } else {
          refdum2 = refdum2 + refdum1;
          refdum1 = refdum1 / 2;
          r_c = ((nrrextra + refdum2) * (RC(1) - rstar) / (nrrextra + 1.0)) + rstar;
        }
      } else {
        free(zhat0); free(xhat0);
        STOP(1, "Do not know imethod");
      }
      ADD_RING();
    }
  }
  /* telescope.F:1066-1138 */
  if (irmax < c->nr) irmax = c->nr;
  for (ix = irmin; ix <= irmax - 1; ix++) {
    r_c = RC(ix) * (1.0 + epsrrr);
    ADD_RING();
    if (ix < irmax && dbdr > 1) {
      dr = (RC(ix + 1) - RC(ix)) / (1.0 * dbdr);
      for (iins = 1; iins <= dbdr - 1; iins++) {
        r_c = RC(ix) + iins * dr;
        ADD_RING();
      }
    }
  }
#undef ADD_RING
  c->rays_amount = ir - 1;
  c->rays_nrr = iradius - 1;
  c->rp_nrrayextra = nrrextra;
  c->rp_nrref = nrref;
  c->rp_dbdr = dbdr;
  free(zhat0);
  free(xhat0);
  c->imcir_r = (double *)xcalloc((size_t)c->rays_nrr + 3, sizeof(double));
  c->imcir_ri = (double *)xcalloc((size_t)c->rays_nrr + 3, sizeof(double));
  c->minvel = (float *)xcalloc((size_t)c->rays_amount + 1, sizeof(float));
  c->maxvel = (float *)xcalloc((size_t)c->rays_amount + 1, sizeof(float));
  return 0;
}

/* telescope.F:443-488 ring radii and edges */
static int setup_ring_edges(orc_ctx *c) {
  int ir, nb = c->rp_dbdr * (c->nr - 1) + c->rp_nrrayextra + c->rp_nrref;
  c->imcir_r[0] = 0.0;
  for (ir = 1; ir <= c->rays_nrr; ir++) c->imcir_r[ir] = c->rays_r[ir];
  c->imcir_ri[0] = 0.0;
  for (ir = 1; ir <= nb; ir++) c->imcir_ri[ir] = 0.5 * (c->imcir_r[ir] + c->imcir_r[ir - 1]);
  c->imcir_ri[nb + 1] = RC(c->nr);
  for (ir = 1; ir <= nb - 1; ir++)
    if (c->imcir_r[ir] - c->imcir_ri[ir] > 2 * (c->imcir_ri[ir + 1] - c->imcir_r[ir]))
      c->imcir_ri[ir] = c->imcir_r[ir] - 2 * (c->imcir_ri[ir + 1] - c->imcir_r[ir]);
  if (c->imcir_ri[1] < c->rstar) {
    if (c->imcir_r[1] < c->rstar) STOP(1, "INTERNAL ERROR IN RAY-SETUP...");
    c->imcir_ri[1] = c->rstar;
  }
  return 0;
}

int orc_set_camera(orc_ctx *c, double anginf, int nphi, int nrext, int dbdr, double rstar,
                   int imethod, int nrref) {
  int code;
  if (!c->rc) { snprintf(c->err, sizeof c->err, "set_camera: call set_grid first"); return 13; }
  c->anginf = anginf; c->nphi = nphi; c->nrext = nrext; c->dbdr = dbdr; c->rstar = rstar;
  c->imethod = imethod; c->nrref = nrref;
  c->jb_armed = 1;
  if ((code = setjmp(c->jb)) != 0) { c->jb_armed = 0; return code; }
  setup_rays_circular(c);
  setup_ring_edges(c);
  c->jb_armed = 0;
  c->cam_set = 1;
  return 0;
}

int orc_get_camera_dims(orc_ctx *c, int *nrr, int *nphi, int *nray) {
  if (!c->cam_set) return 13;
  if (nrr) *nrr = c->rays_nrr;
  if (nphi) *nphi = c->rays_nrphi;
  if (nray) *nray = c->rays_amount;
  return 0;
}
int orc_get_rings(orc_ctx *c, double *rays_r, double *imcir_ri) {
  int i;
  if (!c->cam_set) return 13;
  for (i = 0; i <= c->rays_nrr; i++) rays_r[i] = c->rays_r[i];
  for (i = 0; i <= c->rays_nrr + 1; i++) imcir_ri[i] = c->imcir_ri[i];
  return 0;
}

/* ---- telescope.F:2787-3720 make_trajectory_c(x0,z0,theta0,send,istar=0) ----------------- */
static void make_trajectory_c(orc_ctx *c, double x0, double z0, double theta0, double send) {
  const double pi = PICONST, eps = TELESC_EPS, epsplus = 1.e1 * TELESC_EPS;
  const int nr = c->nr, nt = c->nt;
  double *th_radius = c->th_radius, *th_theta = c->th_theta, *th_s = c->th_s;
  double *r_radius = c->r_radius, *r_theta = c->r_theta, *r_s = c->r_s;
  double *sar1 = c->sar1, *sar2 = c->sar2, *rrgrid = c->rrgrid, *ttgrid = c->ttgrid;
  int *th_ir = c->th_ir, *th_itheta = c->th_itheta, *r_ir = c->r_ir, *r_itheta = c->r_itheta;
  int *iyar = c->iyar;
  double ex_radius[RAYEXPT + 1], ex_theta[RAYEXPT + 1], ex_s[RAYEXPT + 1];
  int ex_ir[RAYEXPT + 1], ex_itheta[RAYEXPT + 1];
  int ix, iy, iyeq, iyend, is, iss, ist, isr, iad, irng, isex, nrex, isnr, isdblnr, isrt = 0;
  int ir_min, ith_amount, ir_amount, iup;
  double pitheta0, r, theta, pitheta, tanth2, sdiscr, costh0, sinth0, costh02, sinth02, a, b, cc;
  double s0, s1, s2, s3, bimpact, dum1, ds, rmaxr, rmaxt, rminr, rmint, sbeg, sprev, rr;
  double snew, znew, bnew, sinphi, dummy;

  memset(ex_ir, 0, sizeof ex_ir);
  memset(ex_itheta, 0, sizeof ex_itheta);
  memset(ex_radius, 0, sizeof ex_radius);
  memset(ex_theta, 0, sizeof ex_theta);
  r_s[0] = 0.0; /* r_s(0) is read uninitialised for the outermost ring; convention 0 */

  for (ix = 1; ix <= nr; ix++) rrgrid[ix] = RC(ix); /* leaves ix = nr+1 */
  for (iy = 1; iy <= nt; iy++) ttgrid[iy] = TC(iy); /* leaves iy = nt+1 */
  pitheta0 = 0.5 * pi - theta0;
  costh0 = cos(theta0);
  sinth0 = sin(theta0);
  costh02 = costh0 * costh0;
  sinth02 = sinth0 * sinth0;
  iyeq = nt / 2;
  iyend = nt;
  /* theta crossings: telescope.F:2958-3000 */
  is = 1;
  for (iy = 1; iy <= iyeq; iy++) {
    double t;
    theta = TC(iy);
    t = tan(theta);
    tanth2 = t * t;
    a = tanth2 * costh02 - sinth02;
    b = 2.0 * tanth2 * costh0 * z0;
    cc = tanth2 * z0 * z0 - x0 * x0;
    sdiscr = b * b - 4.0 * a * cc;
    if (sdiscr > 0.0) {
      sdiscr = sqrt(sdiscr);
      iyar[is] = iy;
      sar1[is] = (-b - sdiscr) / (2.0 * a);
      sar2[is] = (-b + sdiscr) / (2.0 * a);
      if (sar1[is] > sar2[is]) {
        dum1 = sar1[is];
        sar1[is] = sar2[is];
        sar2[is] = dum1;
      }
      is = is + 1;
    }
  } /* iy = iyeq+1 */
  isnr = is - 1;
  isdblnr = 0;
  for (is = 1; is <= isnr; is++) {
    pitheta = 0.5 * pi - TC(iyar[is]);
    if (fabs(pitheta) > fabs(pitheta0)) isdblnr = isdblnr + 1;
  }
  iup = (pitheta0 > 0.0) ? 1 : 0;
#define TH_SET(iss_, sval_, itheta_)                                                        \
  do {                                                                                      \
    double sv_ = (sval_);                                                                   \
    th_radius[iss_] = sqrt(x0 * x0 + z0 * z0 + sv_ * sv_ + 2.0 * z0 * costh0 * sv_);        \
    th_s[iss_] = sv_;                                                                       \
    th_itheta[iss_] = (itheta_);                                                            \
    th_theta[iss_] = TC(th_itheta[iss_]);                                                   \
  } while (0)
  if (z0 * pitheta0 > 0.0) { /* telescope.F:3024-3112 */
    is = 1 + isdblnr;
    for (iss = 1; iss <= isnr - isdblnr; iss++) {
      TH_SET(iss, sar1[is], iup == 1 ? iyend + 1 - iyar[is] : iyar[is]);
      is = is + 1;
    }
    is = isnr;
    for (iss = isnr - isdblnr + 1; iss <= 2 * isnr - 2 * isdblnr; iss++) {
      TH_SET(iss, sar2[is], iup == 1 ? iyar[is] : iyend + 1 - iyar[is]);
      is = is - 1;
    }
    if (isdblnr > 0) {
      for (iss = 2 * isnr + 1 - 2 * isdblnr; iss <= 2 * isnr - isdblnr; iss++) {
        TH_SET(iss, sar1[is], iup == 1 ? iyar[is] : iyend + 1 - iyar[is]);
        is = is - 1;
      }
      is = is + 1;
      for (iss = 2 * isnr + 1 - isdblnr; iss <= 2 * isnr; iss++) {
        TH_SET(iss, sar2[is], iup == 1 ? iyar[is] : iyend + 1 - iyar[is]);
        is = is + 1;
      }
    }
  } else { /* telescope.F:3113-3194 */
    is = isdblnr;
    if (isdblnr > 0) {
      for (iss = 1; iss <= isdblnr; iss++) {
        TH_SET(iss, sar1[is], iup == 1 ? iyend + 1 - iyar[is] : iyar[is]);
        is = is - 1;
      }
      is = is + 1;
      for (iss = isdblnr + 1; iss <= 2 * isdblnr; iss++) {
        TH_SET(iss, sar2[is], iup == 1 ? iyend + 1 - iyar[is] : iyar[is]);
        is = is + 1;
      }
    }
    is = isdblnr + 1;
    for (iss = 2 * isdblnr + 1; iss <= isnr + isdblnr; iss++) {
      TH_SET(iss, sar1[is], iup == 1 ? iyend + 1 - iyar[is] : iyar[is]);
      is = is + 1;
    }
    is = isnr;
    for (iss = isnr + isdblnr + 1; iss <= 2 * isnr; iss++) {
      TH_SET(iss, sar2[is], iup == 1 ? iyar[is] : iyend + 1 - iyar[is]);
      is = is - 1;
    }
  }
#undef TH_SET
  ith_amount = 2 * isnr;
  for (is = 1; is <= ith_amount; is++) { /* telescope.F:3204-3214 */
    r = th_radius[is];
    if (r < RC(1))
      th_ir[is] = 0;
    else if (r > RC(nr))
      th_ir[is] = nr;
    else {
      hunt1(rrgrid, nr, r, &ix);
      th_ir[is] = ix;
    }
  }
  /* R crossings: telescope.F:3241-3344 */
  bimpact = sqrt(x0 * x0 + z0 * z0 * (1.0 - costh02));
  for (ix = 1; ix <= nr; ix++)
    if (RC(ix) > bimpact) goto l30;
  STOP(13, "Cannot find minimal approach radius!");
l30:
  ir_min = ix;
  if (bimpact <= c->rstar) ir_min = 0;
  c->tr_b = bimpact;
  ir_amount = 2 * (nr + 1 - ir_min);
  is = 1;
  for (ix = nr; ix >= ir_min; ix--) {
    if (ix == 0) r = c->rstar;
    else r = RC(ix);
    a = 1.0;
    b = 2.0 * costh0 * z0;
    cc = x0 * x0 + z0 * z0 - r * r;
    sdiscr = b * b - 4.0 * a * cc + eps * b * b;
    if (sdiscr < 0.0) STOP(13, "BUG IN CODE make_trajectory_i() 2 !");
    sdiscr = sqrt(sdiscr);
    s1 = (-b - sdiscr) / (2.0 * a);
    s2 = (-b + sdiscr) / (2.0 * a);
    r_theta[is] = atan(sqrt(x0 * x0 + sinth02 * s1 * s1) / (z0 + costh0 * s1));
    if (r_theta[is] < 0.0) r_theta[is] = r_theta[is] + pi;
    r_s[is] = s1;
    r_ir[is] = ix;
    r_radius[is] = r;
    r_theta[ir_amount + 1 - is] = atan(sqrt(x0 * x0 + sinth02 * s2 * s2) / (z0 + costh0 * s2));
    if (r_theta[ir_amount + 1 - is] < 0.0) r_theta[ir_amount + 1 - is] += pi;
    r_s[ir_amount + 1 - is] = s2;
    r_ir[ir_amount + 1 - is] = ix;
    r_radius[ir_amount + 1 - is] = r;
    is = is + 1;
  } /* ix = ir_min-1 */
  for (is = 1; is <= ir_amount; is++) {
    theta = r_theta[is];
    hunt1(ttgrid, nt, theta, &iy);
    r_itheta[is] = iy;
  }
  /* merge set-up: telescope.F:3352-3366 */
  c->tr_amount = ir_amount + ith_amount;
  ist = 1;
  isr = 1;
  th_s[ith_amount + 1] = 1.e30;
  r_s[ir_amount + 1] = 1.e30;
  rmaxt = RC(nr) * (1.0 - epsplus);
  rmaxr = RC(nr) * (1.0 + epsplus);
  rmint = RC(1) * (1.0 + epsplus);
  rminr = RC(1) * (1.0 - epsplus);
  is = 1;
  sbeg = -1.e30;
  sprev = -1.e30;
  /* (istar==0: star-surface start never taken, telescope.F:3370-3389) */
  /* extra points: telescope.F:3417-3600 */
  for (isex = 1; isex <= RAYEXPT; isex++) ex_s[isex] = 1.e30;
  isex = 1;
  if (bimpact > RC(1)) { /* radial extremum, telescope.F:3434-3491 */
    s2 = 0.0 - z0 * costh0;
    ex_s[isex] = s2;
    ex_radius[isex] = bimpact;
    ex_theta[isex] = atan(sqrt(x0 * x0 + sinth02 * s2 * s2) / (z0 + costh0 * s2));
    if (ex_theta[isex] < 0.0) ex_theta[isex] = ex_theta[isex] + pi;
    ex_s[isex] = s2;
    ex_ir[isex] = ir_min - 1;
    hunt1(ttgrid, nt, ex_theta[isex], &iy);
    ex_itheta[isex] = iy;
    isex = isex + 1;
    s3 = s2;
    hunt0(r_s, ir_amount, s2, &isrt);
    for (irng = 1; irng <= 4; irng++) {
      if (irng == 1) { s1 = s3; s2 = r_s[isrt + 1]; }
      else if (irng == 2) { s1 = r_s[isrt]; s2 = s3; }
      else if (irng == 3) { s1 = r_s[isrt + 1]; s2 = r_s[isrt + 2]; }
      else { s1 = r_s[isrt - 1]; s2 = r_s[isrt]; }
      ds = (s2 - s1) / (1.0 + 1.0 * RAYADPT);
      for (iad = 1; iad <= RAYADPT; iad++) {
        s0 = iad * ds + s1;
        ex_s[isex] = s0;
        rr = sqrt(x0 * x0 + z0 * z0 + s0 * s0 + 2.0 * z0 * costh0 * s0);
        ex_radius[isex] = rr;
        ex_theta[isex] = atan(sqrt(x0 * x0 + sinth02 * s0 * s0) / (z0 + costh0 * s0));
        if (ex_theta[isex] < 0.0) ex_theta[isex] = ex_theta[isex] + pi;
        hunt1(rrgrid, nr, ex_radius[isex], &ix);
        ex_ir[isex] = ix;
        hunt1(ttgrid, nt, ex_theta[isex], &iy);
        ex_itheta[isex] = iy;
        isex = isex + 1;
      }
    }
  }
  /* theta extremum, telescope.F:3498-3589 */
  s2 = x0 * x0 * costh0 / (z0 * sinth02);
  rr = sqrt(x0 * x0 + z0 * z0 + s2 * s2 + 2.0 * z0 * costh0 * s2);
  if (rr > RC(1) && rr < RC(nr)) {
    ex_s[isex] = s2;
    ex_radius[isex] = rr;
    ex_theta[isex] = atan(sqrt(x0 * x0 + sinth02 * s2 * s2) / (z0 + costh0 * s2));
    if (ex_theta[isex] < 0.0) ex_theta[isex] = ex_theta[isex] + pi;
    ex_s[isex] = s2;
    hunt1(rrgrid, nr, ex_radius[isex], &ix);
    if (ix == 0 || ix == nr) STOP(192, "telescope.F/make_traject_t(): hunt failed");
    ex_ir[isex] = ix;
    hunt1(ttgrid, nt, ex_theta[isex], &iy);
    ex_itheta[isex] = iy;
    isex = isex + 1;
    s3 = s2;
    hunt0(th_s, ith_amount, s3, &isrt);
    if ((isrt - RAYRNPT + 1 < 1) || (isrt + RAYRNPT > ith_amount)) goto l78;
    for (irng = 1; irng <= 4; irng++) {
      if (irng == 1) { s1 = s3; s2 = th_s[isrt + 1]; }
      else if (irng == 2) { s1 = th_s[isrt]; s2 = s3; }
      else if (irng == 3) { s1 = th_s[isrt + 1]; s2 = th_s[isrt + 2]; }
      else { s1 = th_s[isrt - 1]; s2 = th_s[isrt]; }
      if (s2 == s1) STOP(987, "s2=s1");
      if (s2 < s1) STOP(988, "s2<s1");
      ds = (s2 - s1) / (1.0 + 1.0 * RAYADPT);
      for (iad = 1; iad <= RAYADPT; iad++) {
        s0 = iad * ds + s1;
        ex_s[isex] = s0;
        rr = sqrt(x0 * x0 + z0 * z0 + s0 * s0 + 2.0 * z0 * costh0 * s0);
        ex_radius[isex] = rr;
        ex_theta[isex] = atan(sqrt(x0 * x0 + sinth02 * s0 * s0) / (z0 + costh0 * s0));
        if (ex_theta[isex] < 0.0) ex_theta[isex] = ex_theta[isex] + pi;
        hunt1(rrgrid, nr, ex_radius[isex], &ix);
        if (ix == 0 || ix == nr) continue; /* goto 79 */
        ex_ir[isex] = ix;
        hunt1(ttgrid, nt, ex_theta[isex], &iy);
        ex_itheta[isex] = iy;
        isex = isex + 1;
      }
    }
  }
l78:
  nrex = isex - 1;
  isex = 1;
  ray_sort(nrex, ex_s, ex_radius, ex_theta, ex_ir, ex_itheta);
  /* 3-way merge: telescope.F:3607-3677 */
  {
    int total = c->tr_amount;
    for (iss = 1; iss <= total; iss++) {
      for (;;) {
        double m = th_s[ist] < r_s[isr] ? th_s[ist] : r_s[isr];
        if (!(ex_s[isex] < m)) break;
        if (ex_radius[isex] <= rmaxt && ex_radius[isex] >= rmint &&
            (ex_s[isex] - sprev) > eps * ex_radius[isex] && ex_s[isex] >= sbeg &&
            ex_s[isex] <= send) {
          c->tr_icross[is] = 3;
          c->tr_radius[is] = ex_radius[isex];
          c->tr_theta[is] = ex_theta[isex];
          c->tr_iradius[is] = ex_ir[isex];
          c->tr_itheta[is] = ex_itheta[isex];
          c->tr_s[is] = ex_s[isex];
          sprev = c->tr_s[is];
          is = is + 1;
        }
        isex = isex + 1;
      }
      if (isex > RAYEXPT) STOP(83459, "Internal exception error");
      if (th_s[ist] < r_s[isr]) {
        if (th_radius[ist] <= rmaxt && th_radius[ist] >= rmint &&
            (th_s[ist] - sprev) > eps * th_radius[ist] && th_s[ist] >= sbeg && th_s[ist] <= send) {
          c->tr_icross[is] = 2;
          c->tr_radius[is] = th_radius[ist];
          c->tr_theta[is] = th_theta[ist];
          c->tr_iradius[is] = th_ir[ist];
          c->tr_itheta[is] = th_itheta[ist];
          c->tr_s[is] = th_s[ist];
          sprev = c->tr_s[is];
          is = is + 1;
        }
        ist = ist + 1;
      } else {
        if (r_radius[isr] <= rmaxr && r_radius[isr] >= rminr &&
            (r_s[isr] - sprev) > eps * r_radius[isr] && r_s[isr] >= sbeg && r_s[isr] <= send) {
          c->tr_icross[is] = 1;
          c->tr_radius[is] = r_radius[isr];
          c->tr_theta[is] = r_theta[isr];
          c->tr_iradius[is] = r_ir[isr];
          c->tr_itheta[is] = r_itheta[isr];
          c->tr_s[is] = r_s[isr];
          sprev = c->tr_s[is];
          is = is + 1;
        }
        isr = isr + 1;
      }
    }
  }
  c->tr_amount = is - 1;
  /* local direction: telescope.F:3687-3718 */
  znew = z0 * sinth02;
  bnew = sqrt(x0 * x0 + z0 * z0 * sinth02);
  for (is = 1; is <= c->tr_amount; is++) {
    snew = c->tr_s[is] + z0 * costh0;
    c->tr_mu[is] = snew / sqrt(bnew * bnew + snew * snew);
    dummy = bnew * sqrt(bnew * bnew + snew * snew - (znew + snew * costh0) * (znew + snew * costh0));
    if (dummy > 0.0)
      sinphi = (bnew * bnew * costh0 - znew * snew) / dummy;
    else
      sinphi = epsplus;
    if (x0 < 0.0)
      sinphi = asin(sinphi);
    else
      sinphi = pi - asin(sinphi);
    while (sinphi < 0.0) sinphi = sinphi + 2.0 * pi;
    while (sinphi >= 2.0 * pi) sinphi = sinphi - 2.0 * pi;
    c->tr_phi[is] = sinphi;
  }
}

/* ---- line.F:2650-2662 omega_dot_v ------------------------------------------------------ */
static double omega_dot_v(orc_ctx *c, double mu, double phi, const double *v) {
  if (mu > 1.0) STOP(393, "omega_dot_v: mu>1");
  return 3.335668e-11 * (mu * v[0] + sqrt(1.0 - mu * mu) * (v[1] * sin(phi) + v[2] * cos(phi)));
}

/* ---- line.F:3965-4217 get_line_dust_values --------------------------------------------- */
static void get_line_dust_values(orc_ctx *c, int icr, int ir, int it, double dr, double dt,
                                 double mu, double phi, int iline, double *src_dust,
                                 double *alp_dust, double *dvmu, double *linewidth, double *Nup,
                                 double *Ndown) {
  int indexr[2], indext[2], i;
  double molpg, velocity[3];
  const size_t ncell = (size_t)c->nr * c->nth;
  const double *lsrc = c->ld_src + (size_t)(iline - 1) * ncell;
  const double *lalp = c->ld_alp + (size_t)(iline - 1) * ncell;
  const int lup = c->lev_up[iline - 1], ldn = c->lev_down[iline - 1], nlev = c->nlevels;
  if (dr < 0.0 || dr > 1.0) STOP(6024, "ERROR: Erroneous dr found");
  if (dt < 0.0 || dt > 1.0) STOP(6023, "ERROR: Erroneous dt found");
  molpg = 1.0 / (c->umass_av * 1.6726e-24);
  if (dr > 0.0) {
    indexr[0] = ir;
    indexr[1] = ir + 1;
    if (indexr[1] > c->nr) indexr[1] = c->nr;
  } else {
    indexr[0] = ir;
    indexr[1] = ir - 1;
    if (indexr[1] < 1) indexr[1] = 1;
  }
  if (dt > 0.0) {
    indext[0] = it;
    indext[1] = it + 1;
  } else {
    indext[0] = it;
    indext[1] = it - 1;
  }
  indext[0] = RIDX(indext[0]);
  indext[1] = RIDX(indext[1]);
#define NCELL(lev, t, r_)                                                              \
  (c->popul[CELL(t, r_) * nlev + ((lev)-1)] * c->abund[CELL(t, r_)] * c->rho[CELL(t, r_)] * molpg)
  if (icr == 1) {
    size_t c0 = CELL(indext[0], indexr[0]), c1 = CELL(indext[1], indexr[0]);
    *src_dust = (1.0 - dt) * lsrc[c0] + dt * lsrc[c1];
    *alp_dust = (1.0 - dt) * lalp[c0] + dt * lalp[c1];
    *linewidth = (1.0 - dt) * c->lw[c0] + dt * c->lw[c1];
    *Nup = (1.0 - dt) * NCELL(lup, indext[0], indexr[0]) + dt * NCELL(lup, indext[1], indexr[0]);
    *Ndown = (1.0 - dt) * NCELL(ldn, indext[0], indexr[0]) + dt * NCELL(ldn, indext[1], indexr[0]);
    for (i = 0; i < 3; i++) velocity[i] = (1.0 - dt) * c->vel[3 * c0 + i] + dt * c->vel[3 * c1 + i];
  } else if (icr == 2) {
    size_t c0 = CELL(indext[0], indexr[0]), c1 = CELL(indext[0], indexr[1]);
    *src_dust = (1.0 - dr) * lsrc[c0] + dr * lsrc[c1];
    *alp_dust = (1.0 - dr) * lalp[c0] + dr * lalp[c1];
    *linewidth = (1.0 - dr) * c->lw[c0] + dr * c->lw[c1];
    *Nup = (1.0 - dr) * NCELL(lup, indext[0], indexr[0]) + dr * NCELL(lup, indext[0], indexr[1]);
    *Ndown = (1.0 - dr) * NCELL(ldn, indext[0], indexr[0]) + dr * NCELL(ldn, indext[0], indexr[1]);
    for (i = 0; i < 3; i++) velocity[i] = (1.0 - dr) * c->vel[3 * c0 + i] + dr * c->vel[3 * c1 + i];
  } else {
    size_t c0 = CELL(indext[0], indexr[0]), c1 = CELL(indext[1], indexr[0]);
    size_t c2 = CELL(indext[0], indexr[1]), c3 = CELL(indext[1], indexr[1]);
    double n0, n1, n2, n3;
    *src_dust = (1.0 - dr) * ((1.0 - dt) * lsrc[c0] + dt * lsrc[c1]) +
                dr * ((1.0 - dt) * lsrc[c2] + dt * lsrc[c3]);
    *alp_dust = (1.0 - dr) * ((1.0 - dt) * lalp[c0] + dt * lalp[c1]) +
                dr * ((1.0 - dt) * lalp[c2] + dt * lalp[c3]);
    *linewidth = (1.0 - dr) * ((1.0 - dt) * c->lw[c0] + dt * c->lw[c1]) +
                 dr * ((1.0 - dt) * c->lw[c2] + dt * c->lw[c3]);
    n0 = NCELL(lup, indext[0], indexr[0]); n1 = NCELL(lup, indext[1], indexr[0]);
    n2 = NCELL(lup, indext[0], indexr[1]); n3 = NCELL(lup, indext[1], indexr[1]);
    *Nup = (1.0 - dr) * ((1.0 - dt) * n0 + dt * n1) + dr * ((1.0 - dt) * n2 + dt * n3);
    n0 = NCELL(ldn, indext[0], indexr[0]); n1 = NCELL(ldn, indext[1], indexr[0]);
    n2 = NCELL(ldn, indext[0], indexr[1]); n3 = NCELL(ldn, indext[1], indexr[1]);
    *Ndown = (1.0 - dr) * ((1.0 - dt) * n0 + dt * n1) + dr * ((1.0 - dt) * n2 + dt * n3);
    for (i = 0; i < 3; i++)
      velocity[i] = (1.0 - dr) * ((1.0 - dt) * c->vel[3 * c0 + i] + dt * c->vel[3 * c1 + i]) +
                    dr * ((1.0 - dt) * c->vel[3 * c2 + i] + dt * c->vel[3 * c3 + i]);
  }
#undef NCELL
  *dvmu = omega_dot_v(c, mu, phi, velocity);
}

/* ---- line.F:2280-2314 voigt_profile (a Gaussian) ---------------------------------------- */
static double voigt_profile(double nu00, double aks, double dnu) {
  double nu0 = fabs(nu00), aa, norm, uvoigt;
  aa = 3.33567e-6 * nu0 * aks;
  norm = 0.56419583546 / aa;
  uvoigt = dnu / aa;
  return norm * exp(-(uvoigt * uvoigt));
}

typedef struct {
  double phiprof0, srcl0, alpl0;
  int init;
} carry_t;

/* ---- line.F:4515-4624 integrate_element_linedust --------------------------------------- */
static void integrate_element_linedust(orc_ctx *c, double *intensity, double ds, double srcd0,
                                       double srcd1, double alpd0, double alpd1, double lw0,
                                       double lw1, double dvmu0, double dvmu1, double nup0,
                                       double nup1, double ndown0, double ndown1, int inu,
                                       int iline, carry_t *k) {
  double lwav, dnu0, dnu1, phiprof1, srcl1, alpl1, src0, src1, alp0, alp1;
  const double lf = c->linefreq[iline - 1], A = c->aud[iline - 1];
  const double Bdu = c->bdu[iline - 1], Bud = c->bud[iline - 1];
  c->cnt_E += 1.0;
  lwav = 0.5 * (lw0 + lw1);
  dnu0 = c->line_dnu[inu] - c->line_nu0 * dvmu0;
  dnu1 = c->line_dnu[inu] - c->line_nu0 * dvmu1;
  if (k->init == 1) k->phiprof0 = voigt_profile(lf, lwav, dnu0);
  phiprof1 = voigt_profile(lf, lwav, dnu1);
  if (k->init == 1) k->srcl0 = 5.27296241956e-28 * lf * nup0 * A * k->phiprof0;
  srcl1 = 5.27296241956e-28 * lf * nup1 * A * phiprof1;
  if (k->init == 1) k->alpl0 = 5.27296241956e-28 * lf * k->phiprof0 * (ndown0 * Bdu - nup0 * Bud);
  alpl1 = 5.27296241956e-28 * lf * phiprof1 * (ndown1 * Bdu - nup1 * Bud);
  src0 = srcd0 + k->srcl0;
  src1 = srcd1 + srcl1;
  alp0 = alpd0 + k->alpl0;
  alp1 = alpd1 + alpl1;
  *intensity = orc_qdr_src_2(*intensity, src0, alp0, src1, alp1, ds);
  c->char_tau = c->char_tau + 0.5 * (alp0 + alp1) * ds;
  c->char_emis = c->char_emis + 0.5 * (src0 + src1) * ds;
  k->phiprof0 = phiprof1;
  k->srcl0 = srcl1;
  k->alpl0 = alpl1;
  k->init = 0;
}

/* ---- line.F:4636-4848 clever_integrate_element_linedust -------------------------------- */
static void clever_integrate_element_linedust(orc_ctx *c, double *intensity, double ds,
                                              double srcd0, double srcd1, double alpd0,
                                              double alpd1, double lw0, double lw1, double dvmu0,
                                              double dvmu1, double nup0, double nup1,
                                              double ndown0, double ndown1, int inu, int iline,
                                              carry_t *k) {
  const double crit_linecenter = 3.0;
  if (c->subgrid) {
    double lw = 0.5 * (lw0 + lw1);
    double ds_over_deltal_s = fabs((dvmu1 - dvmu0) / (lw / 2.99792458e5));
    if (2.0 * crit_linecenter * ds_over_deltal_s > 1.0) {
      double s_c = ds * ((c->line_dnu[inu] / c->line_nu0) - dvmu0) / (dvmu1 - dvmu0);
      double dls = ds / ds_over_deltal_s;
      double sright = s_c + crit_linecenter * dls;
      double sleft = s_c - crit_linecenter * dls;
      if (sright > 0.0 && sleft < ds) {
        double lg_s[LG_NRMAX + 3], lg_dvmu[LG_NRMAX + 3], lg_nup[LG_NRMAX + 3];
        double lg_ndown[LG_NRMAX + 3], lg_srcd[LG_NRMAX + 3], lg_alpd[LG_NRMAX + 3];
        double lg_ds = (sright - sleft) / (LG_NRMAX - 1.0);
        int lg_nr = 1, i;
        lg_s[1] = 0.0; lg_nup[1] = nup0; lg_ndown[1] = ndown0; lg_dvmu[1] = dvmu0;
        lg_srcd[1] = srcd0; lg_alpd[1] = alpd0;
        for (i = 1; i <= LG_NRMAX; i++) {
          double s = sleft + (i - 1) * lg_ds;
          if (s > 0.0 && s < ds) {
            double eps = s / ds, epsp = 1.0 - eps;
            lg_nr = lg_nr + 1;
            lg_s[lg_nr] = s;
            lg_nup[lg_nr] = epsp * nup0 + eps * nup1;
            lg_ndown[lg_nr] = epsp * ndown0 + eps * ndown1;
            lg_dvmu[lg_nr] = epsp * dvmu0 + eps * dvmu1;
            lg_srcd[lg_nr] = epsp * srcd0 + eps * srcd1;
            lg_alpd[lg_nr] = epsp * alpd0 + eps * alpd1;
          }
        }
        lg_nr = lg_nr + 1;
        lg_s[lg_nr] = ds; lg_nup[lg_nr] = nup1; lg_ndown[lg_nr] = ndown1; lg_dvmu[lg_nr] = dvmu1;
        lg_srcd[lg_nr] = srcd1; lg_alpd[lg_nr] = alpd1;
        for (i = 2; i <= lg_nr; i++) {
          lg_ds = lg_s[i] - lg_s[i - 1];
          integrate_element_linedust(c, intensity, lg_ds, lg_srcd[i - 1], lg_srcd[i],
                                     lg_alpd[i - 1], lg_alpd[i], lw, lw, lg_dvmu[i - 1],
                                     lg_dvmu[i], lg_nup[i - 1], lg_nup[i], lg_ndown[i - 1],
                                     lg_ndown[i], inu, iline, k);
        }
        return;
      }
    }
  }
  integrate_element_linedust(c, intensity, ds, srcd0, srcd1, alpd0, alpd1, lw0, lw1, dvmu0, dvmu1,
                             nup0, nup1, ndown0, ndown1, inu, iline, k);
}

/* ---- telescope.F:3889-4312 charintline ------------------------------------------------- */
static double charintline(orc_ctx *c, int iline, int inu, int iray, double rbeam0) {
  double charint = 0.0, ds, r, theta, dr, dt, mu, phi, s0, s1, starfract;
  double srcd0 = 0, srcd1, alpd0 = 0, alpd1, lw0 = 0, lw1, dvmu0 = 0, dvmu1, Nup0 = 0, Nup1;
  double Ndown0 = 0, Ndown1;
  int is, ir, it, icr, ir_old, icr_old, amount, istar_done = 0;
  carry_t k;
  k.init = 1;
  k.phiprof0 = k.srcl0 = k.alpl0 = 0.0;
  c->cnt_R += 1.0;
  c->char_tau = 0.0;
  icr = c->tr_icross[1];
  r = c->tr_radius[1];
  theta = c->tr_theta[1];
  mu = c->tr_mu[1];
  phi = c->tr_phi[1];
  ir = c->tr_iradius[1];
  it = c->tr_itheta[1];
  dr = (r - RC(ir)) / (RC(ir + 1) - RC(ir));
  dt = (theta - TC(it)) / (TC(it + 1) - TC(it));
  if (ir == 0) {
    charint = c->starspec_line[inu];
  } else {
    if (c->out_itype == 2) {
      double freq = c->linefreq[iline - 1], temp = TEMPCMB;
      charint = 1.47455253991e-47 * (freq * freq * freq) / (exp(4.7991598e-11 * freq / temp) - 1.0);
    } else if (c->out_itype == 0) {
      charint = 0.0;
    } else if (c->out_itype == 1) {
      STOP(13, "Outer BC type 1 not allowed for telescope");
    } else if (c->out_itype == 3) {
      charint = c->isrf_line[inu];
    } else {
      STOP(13, "Telecope: dont know this type of outer BC");
    }
    get_line_dust_values(c, icr, ir, it, dr, dt, mu, phi, iline, &srcd0, &alpd0, &dvmu0, &lw0,
                         &Nup0, &Ndown0);
  }
  amount = c->tr_amount;
  for (is = 2; is <= amount; is++) {
    c->cnt_S += 1.0;
    s1 = c->tr_s[is];
    s0 = c->tr_s[is - 1];
    ds = (s1 - s0);
    if (ds < 0.0) STOP(749, "charintline: ds<0");
    ir_old = ir;
    icr_old = icr;
    icr = c->tr_icross[is];
    r = c->tr_radius[is];
    theta = c->tr_theta[is];
    mu = c->tr_mu[is];
    phi = c->tr_phi[is];
    ir = c->tr_iradius[is];
    it = c->tr_itheta[is];
    if (ir == 0) STOP(7454, "The stellar surface is not done correctly");
    dr = (r - RC(ir)) / (RC(ir + 1) - RC(ir));
    dt = (theta - TC(it)) / (TC(it + 1) - TC(it));
    if (ir == 1 && ir_old == 1 && icr == 1 && icr_old == 1) { ds = 0.0; k.init = 1; }
    if (ir == 1 && ir_old == 0 && icr == 1 && icr_old == 1) { ds = 0.0; k.init = 1; }
    if (ir == 0 && ir_old == 1) STOP(137, "Huh?? Not possible... (charintline())");
    if (!(ir > 1)) {
      if (ir == -1) STOP(13, "SAFETY STOP: charintline(): ir.eq.-1");
      if (c->in_itype == 0) {
        STOP(13, "STOPPED: Inner BC type 0 is temporarily disabled");
      } else if (c->in_itype == 1) {
        if (ir_old == 1 && icr_old == 1) { charint = 0.0; k.init = 1; }
      } else if (c->in_itype == 2) {
        if ((ir == 1 || ir == -1) && istar_done == 0 && rbeam0 > 0.0) {
          if (rbeam0 < c->rstar) STOP(124, "central beam smaller than the stellar radius");
          starfract = (c->rstar / rbeam0) * (c->rstar / rbeam0);
          charint = (1.0 - starfract) * charint + starfract * c->starspec_line[inu];
          istar_done = 1;
          k.init = 1;
        }
        if (ir == 1 && ir_old == 1 && rbeam0 == 0.0 && c->tr_b <= c->rstar) {
          charint = c->starspec_line[inu];
          k.init = 1;
        }
      } else {
        STOP(13, "Dont know this type of inner bc");
      }
    }
    get_line_dust_values(c, icr, ir, it, dr, dt, mu, phi, iline, &srcd1, &alpd1, &dvmu1, &lw1,
                         &Nup1, &Ndown1);
    clever_integrate_element_linedust(c, &charint, ds, srcd0, srcd1, alpd0, alpd1, lw0, lw1, dvmu0,
                                      dvmu1, Nup0, Nup1, Ndown0, Ndown1, inu, iline, &k);
    if (dvmu0 < (double)c->minvel[iray] && Nup0 + Ndown0 > c->levthres)
      c->minvel[iray] = (float)dvmu0;
    if (dvmu0 > (double)c->maxvel[iray] && Nup0 + Ndown0 > c->levthres)
      c->maxvel[iray] = (float)dvmu0;
    srcd0 = srcd1;
    alpd0 = alpd1;
    lw0 = lw1;
    dvmu0 = dvmu1;
    Nup0 = Nup1;
    Ndown0 = Ndown1;
    if (k.alpl0 * ds < (double)(-0.01f)) c->maserflag = 1; /* telescope.F:4295 REAL literal */
  }
  return charint;
}

/* ---- line.F:427-545 line_setup_passband; 3797-3845 starbc; 3855-3903 outerbc ------------ */
static void setup_passband_and_bc(orc_ctx *c, int iline, double width, int nfr) {
  double nu0 = fabs(c->linefreq[iline - 1]), passb, nu1, dnu;
  int k, inudust = 0;
  if (nu0 == 0.0) STOP(13, "Problem in line_setup_passband(): nu0=0 !");
  if (nfr < 1) STOP(13, "Number of frequencies for this line is out of range");
  if (nfr == 1) STOP(13, "ERROR: Simple square line profile deactivated");
  if (c->nfr != nfr || !c->line_dnu) {
    free(c->line_dnu); free(c->freq_nu); free(c->starspec_line); free(c->isrf_line);
    c->line_dnu = (double *)xcalloc((size_t)nfr + 2, sizeof(double));
    c->freq_nu = (double *)xcalloc((size_t)nfr + 2, sizeof(double));
    c->starspec_line = (double *)xcalloc((size_t)nfr + 2, sizeof(double));
    c->isrf_line = (double *)xcalloc((size_t)nfr + 2, sizeof(double));
    c->nfr = nfr;
  }
  passb = 3.33567e-6 * nu0 * width;
  nu1 = 0.0 - passb;
  dnu = 2.0 * passb / (nfr - 1.0);
  for (k = 1; k <= nfr; k++) {
    c->line_dnu[k] = nu1 + (k - 1) * dnu;
    c->freq_nu[k] = nu0 + c->line_dnu[k];
  }
  c->line_nu0 = c->linefreq[iline - 1];
  /* star BC */
  if (c->ncf_b == 0) STOP(1, "Cannot use line stellar BC without having read the stellar spectrum.");
  {
    const double *cf = c->cfreq_b - 1;
    int ncf = c->ncf_b;
    for (k = 1; k <= nfr; k++) {
      double freq = c->freq_nu[k];
      hunt1(cf, ncf, freq, &inudust);
      if (inudust == 0 || inudust == ncf) {
        c->starspec_line[k] = 0.0;
      } else {
        double wgt = (freq - cf[inudust]) / (cf[inudust + 1] - cf[inudust]);
        c->starspec_line[k] =
            (1.0 - wgt) * c->starspec_cont[inudust - 1] + wgt * c->starspec_cont[inudust];
      }
    }
    if (c->out_itype == 3) {
      if (!c->isrf_cont) STOP(1, "Cannot use line outer BC without the interstellar spectrum.");
      inudust = 0;
      for (k = 1; k <= nfr; k++) {
        double freq = c->freq_nu[k];
        hunt1(cf, ncf, freq, &inudust);
        if (inudust == 0 || inudust == nfr) { /* line.F:3892: compares with freq_nr (sic) */
          c->isrf_line[k] = 0.0;
        } else {
          double wgt = (freq - cf[inudust]) / (cf[inudust + 1] - cf[inudust]);
          c->isrf_line[k] = (1.0 - wgt) * c->isrf_cont[inudust - 1] + wgt * c->isrf_cont[inudust];
        }
      }
    } else {
      for (k = 1; k <= nfr; k++) c->isrf_line[k] = 0.0;
    }
  }
}

/* line.F:2968-3033: aksmax = global max of locprof_linewidth */
static double compute_aksmax(orc_ctx *c) {
  double aksmax = 0.0;
  size_t i, n = (size_t)c->nr * c->nth;
  for (i = 0; i < n; i++)
    if (c->lw[i] > aksmax) aksmax = c->lw[i];
  return aksmax;
}

/* line.F:1600-1673 global_prepare_transitions: the reference recomputes the B's for every cell
 * on every line; kept here (cost only) so the CPU baseline carries the reference's redundancy. */
static void global_prepare_transitions(orc_ctx *c) {
  int ir, it;
  for (ir = 1; ir <= c->nr; ir++)
    for (it = 1; it <= c->nth; it++) prepare_lines(c);
}

/* ---- telescope.F:339-631 make_image_circular + 1320-1437 calc_freq_flux_observer -------- */
#define IMIDX(inu, iphi, ir) (((size_t)(ir) * (size_t)nphi + (size_t)((iphi)-1)) * (size_t)nfr + (size_t)((inu)-1))
static void render_line(orc_ctx *c, int iline, int nfr, double passband, double distance,
                        double *flux, double *imcir_out, int *cmask_out, double *tau_center,
                        int *maser_out, double *velo_out) {
  const int nphi = c->rays_nrphi, nrr = c->rays_nrr;
  double *imcir_int, *imcir_cont, *velo, aksmax, aksmax_c;
  int inu, iray, ir, iphi;
  size_t nim = (size_t)(nrr + 1) * nphi * nfr;
  c->maserflag = 0;
  setup_passband_and_bc(c, iline, passband, nfr);
  global_prepare_transitions(c);
  imcir_int = imcir_out ? imcir_out : (double *)xcalloc(nim, sizeof(double));
  imcir_cont = (double *)xcalloc((size_t)(nrr + 1) * nphi + 1, sizeof(double));
  velo = (double *)xcalloc((size_t)nfr + 2, sizeof(double));
  if (!c->cmask_persist || c->cmask_nfr != nfr) {
    free(c->cmask_persist);
    c->cmask_persist = (int *)xcalloc(nim, sizeof(int));
    c->cmask_nfr = nfr;
  }
  for (iray = 0; iray <= c->rays_amount; iray++) {
    c->minvel[iray] = 1;
    c->maxvel[iray] = -1;
  }
  for (inu = 1; inu <= nfr; inu++) velo[inu] = c->line_dnu[inu] / c->line_nu0;
  aksmax = (c->aksmax_opt >= 0.0) ? c->aksmax_opt : compute_aksmax(c);
  aksmax_c = aksmax / 2.99792458e5;
  /* centre ray */
  iray = 1;
  make_trajectory_c(c, c->rp_x0[iray], c->rp_z0[iray], c->rp_theta0[iray], c->rp_s0[iray]);
  for (inu = 1; inu <= nfr; inu++) {
    imcir_int[IMIDX(inu, 1, 0)] = charintline(c, iline, inu, iray, c->imcir_ri[1]);
    if (tau_center) *tau_center = c->char_tau;
    for (iphi = 2; iphi <= nphi; iphi++) imcir_int[IMIDX(inu, iphi, 0)] = imcir_int[IMIDX(inu, 1, 0)];
  }
  iray = 2;
  for (ir = 1; ir <= nrr; ir++)
    for (iphi = 1; iphi <= nphi; iphi++) {
      double *cont = &imcir_cont[(size_t)ir * nphi + (iphi - 1)];
      if (c->ring_hi > 0 && (ir < c->ring_lo || ir > c->ring_hi || (ir - c->ring_lo) % c->ring_stride != 0)) { /* benchmark sampling only */
        for (inu = 1; inu <= nfr; inu++) imcir_int[IMIDX(inu, iphi, ir)] = 0.0;
        iray = iray + 1;
        continue;
      }
      make_trajectory_c(c, c->rp_x0[iray], c->rp_z0[iray], c->rp_theta0[iray], c->rp_s0[iray]);
      inu = 1;
      imcir_int[IMIDX(inu, iphi, ir)] = charintline(c, iline, inu, iray, -1.0);
      c->cmask_persist[IMIDX(inu, iphi, ir)] = 1;
      if (velo[inu] > (double)c->maxvel[iray] + 2.f * aksmax_c ||
          velo[inu] < (double)c->minvel[iray] - 2.f * aksmax_c)
        *cont = imcir_int[IMIDX(inu, iphi, ir)];
      for (inu = 2; inu <= nfr; inu++) {
        if (c->nonredundant) {
          if (velo[inu] <= (double)c->maxvel[iray] + 2.f * aksmax_c &&
              velo[inu] >= (double)c->minvel[iray] - 2.f * aksmax_c) {
            imcir_int[IMIDX(inu, iphi, ir)] = charintline(c, iline, inu, iray, -1.0);
            c->cmask_persist[IMIDX(inu, iphi, ir)] = 1;
          } else {
            if (*cont != 0) {
              imcir_int[IMIDX(inu, iphi, ir)] = *cont;
            } else {
              imcir_int[IMIDX(inu, iphi, ir)] = charintline(c, iline, inu, iray, -1.0);
              *cont = imcir_int[IMIDX(inu, iphi, ir)];
            }
          }
        } else {
          imcir_int[IMIDX(inu, iphi, ir)] = charintline(c, iline, inu, iray, -1.0);
        }
      }
      iray = iray + 1;
    }
  /* flux: telescope.F:1388-1433 (rapert = 1d99) */
  for (inu = 1; inu <= nfr; inu++) {
    double slum = 0.0, surf, dslum;
    const double rapert = 1e99;
    surf = 3.14159265359 * (c->imcir_ri[1] * c->imcir_ri[1]);
    dslum = surf * imcir_int[IMIDX(inu, 1, 0)];
    slum = slum + dslum;
    for (ir = 1; ir <= nrr; ir++) {
      if (c->imcir_ri[ir] < rapert) {
        if (c->imcir_ri[ir + 1] < rapert)
          surf = 3.14159265359 * (c->imcir_ri[ir + 1] * c->imcir_ri[ir + 1] - c->imcir_ri[ir] * c->imcir_ri[ir]);
        else
          surf = 3.14159265359 * (rapert * rapert - c->imcir_ri[ir] * c->imcir_ri[ir]);
        dslum = 0.0;
        for (iphi = 1; iphi <= nphi; iphi++) dslum = dslum + imcir_int[IMIDX(inu, iphi, ir)];
        dslum = dslum / (1.0 * nphi);
        dslum = dslum * surf;
        slum = slum + dslum;
      }
    }
    flux[inu - 1] = slum / (distance * distance);
  }
  if (cmask_out) memcpy(cmask_out, c->cmask_persist, nim * sizeof(int));
  if (maser_out) *maser_out = c->maserflag;
  if (velo_out)
    for (inu = 1; inu <= nfr; inu++) velo_out[inu - 1] = velo[inu];
  if (!imcir_out) free(imcir_int);
  free(imcir_cont);
  free(velo);
}

/* ---- telescope.F:2229-2475 setup_rays_rectang ------------------------------------------------------------ */
static void setup_rays_rectang(orc_ctx *c, int nrx, int nry, double sizepix_x, double sizepix_y, double anginf,
                               double phioffset, double xoffset, double yoffset) {
  int ix, iy, ir, nrxhalf, nryhalf;
  double x_c, y_c, p_c, r_c, theta0, sinth0, xh0, zh0, zh02, dum;
  size_t maxrays = (size_t)nrx * nry + 3;
  if ((nrx + 1) / 2 != nrx / 2) STOP(13, "ERROR Telescope: nrx must be even");
  if ((nry + 1) / 2 != nry / 2) STOP(13, "ERROR Telescope: nry must be even");
  if (sizepix_x <= 0.0) STOP(13, "setup_rays_rectang(): sizepix_x.le.0");
  if (sizepix_y <= 0.0) STOP(13, "setup_rays_rectang(): sizepix_y.le.0");
  if (anginf < 1.e-1) anginf = (double)0.1f; /* telescope.F:2303 assigns the REAL literal 0.1 */
  theta0 = anginf + 1.e-4;
  sinth0 = sin(theta0);
  free(c->rp_x0); free(c->rp_z0); free(c->rp_theta0); free(c->rp_s0); free(c->rp_b);
  c->rp_x0 = (double *)xcalloc(maxrays, sizeof(double));
  c->rp_z0 = (double *)xcalloc(maxrays, sizeof(double));
  c->rp_theta0 = (double *)xcalloc(maxrays, sizeof(double));
  c->rp_s0 = (double *)xcalloc(maxrays, sizeof(double));
  c->rp_b = (double *)xcalloc(maxrays, sizeof(double));
  c->rays_nrx = nrx;
  c->rays_nry = nry;
  c->rays_sizepix_x = sizepix_x;
  c->rays_sizepix_y = sizepix_y;
  ir = 1;
  nrxhalf = nrx / 2;
  nryhalf = nry / 2;
  for (ix = 1; ix <= nrx; ix++)
    for (iy = 1; iy <= nry; iy++) {
      x_c = (ix - nrxhalf - 0.5) * sizepix_x - xoffset;
      y_c = (iy - nryhalf - 0.5) * sizepix_y - yoffset;
      r_c = sqrt(x_c * x_c + y_c * y_c);
      if (x_c == 0.0) x_c = x_c + 0.001 * sizepix_x;
      p_c = atan(y_c / x_c) - phioffset;
      if (x_c < 0.0) p_c = p_c + 3.14159265359;
      while (p_c < 0.0) p_c = p_c + 6.28318530718;
      while (p_c >= 6.28318530718) p_c = p_c - 6.28318530718;
      zh0 = -sin(p_c) / sinth0;
      zh02 = zh0 * zh0;
      dum = 1.0 - zh02 * sinth0 * sinth0;
      dum = dum + 1e-4;
      if (dum < 0.0) STOP(13, "ERROR in setup_rays_rectang");
      xh0 = (cos(p_c) > 0.0) ? sqrt(dum) : -sqrt(dum);
      c->rp_x0[ir] = r_c * xh0;
      c->rp_z0[ir] = r_c * zh0;
      c->rp_theta0[ir] = theta0;
      c->rp_s0[ir] = 1.e30;
      c->rp_b[ir] = r_c;
      ir = ir + 1;
    }
  c->imrec_starunres = 0;
  if (c->imrec_addstar > 0) {
    if (xoffset != 0.0 || yoffset != 0.0) STOP(13, "unresolved central star only for star-centred images");
    c->rp_x0[ir] = 0.0;
    c->rp_z0[ir] = 0.0;
    c->rp_theta0[ir] = theta0;
    c->rp_s0[ir] = 1.e30;
    ir = ir + 1;
    if (sizepix_x * sizepix_x + sizepix_y * sizepix_y > c->rstar * c->rstar) c->imrec_starunres = 1;
  }
  c->rays_amount = ir - 1;
  free(c->minvel); free(c->maxvel);
  c->minvel = (float *)xcalloc((size_t)c->rays_amount + 1, sizeof(float));
  c->maxvel = (float *)xcalloc((size_t)c->rays_amount + 1, sizeof(float));
}

int orc_set_camera_rect(orc_ctx *c, double anginf, int nx, int ny, double sizepix_x, double sizepix_y,
                        double phioffset, double xoffset, double yoffset, double rstar, int addstar) {
  int code;
  if (!c->rc) { snprintf(c->err, sizeof c->err, "set_camera_rect: call set_grid first"); return 13; }
  c->rstar = rstar;
  c->imrec_addstar = addstar;
  c->jb_armed = 1;
  if ((code = setjmp(c->jb)) != 0) { c->jb_armed = 0; return code; }
  setup_rays_rectang(c, nx, ny, sizepix_x, sizepix_y, anginf, phioffset, xoffset, yoffset);
  c->jb_armed = 0;
  c->rect_set = 1;
  c->cam_set = 0; /* the circular camera's ray arrays were replaced */
  return 0;
}

/* ---- telescope.F:2061-2227 make_freq_image_rectang (per line; called from calc_write_line_posvel :1828) ---- */
#define RECIDX(inu, ix, iy) ((((size_t)((ix)-1)) * (size_t)nry + (size_t)((iy)-1)) * (size_t)nfr + (size_t)((inu)-1))
int orc_render_rect(orc_ctx *c, int iline0, int nl, int nfr, double vmax_kms, double *image, double *tau,
                    int *maserflag) {
  int code, l;
  if (!c->rc || !c->rho || !c->nlines || !c->have_line_dust || !c->rect_set || !c->bc_set) {
    snprintf(c->err, sizeof c->err, "stop 13: make_freq_image_rectang(): Ray paramters not yet set");
    return 13;
  }
  if (iline0 < 1 || iline0 + nl - 1 > c->nlines) {
    snprintf(c->err, sizeof c->err, "render_rect: line range out of bounds");
    return 13;
  }
  for (l = 1; l <= c->nr - 1; l++)
    if (RC(l + 1) / RC(l) - 1.0 < 1.e4 * TELESC_EPS) {
      snprintf(c->err, sizeof c->err, "stop 13: radial grid too fine for TELESC_EPS");
      return 13;
    }
  c->jb_armed = 1;
  if ((code = setjmp(c->jb)) != 0) { c->jb_armed = 0; return code; }
  for (l = 0; l < nl; l++) {
    const int iline = iline0 + l, nrx = c->rays_nrx, nry = c->rays_nry;
    double *im = image + (size_t)l * nrx * nry * nfr, *ta = tau ? tau + (size_t)l * nrx * nry * nfr : 0;
    const double rmax = RC(c->nr);
    int ix, iy, inu, iray = 1;
    c->maserflag = 0;
    setup_passband_and_bc(c, iline, vmax_kms, nfr);
    global_prepare_transitions(c);
    for (ix = 1; ix <= nrx; ix++)
      for (iy = 1; iy <= nry; iy++) {
        if (c->rp_b[iray] < 0.999 * rmax) {
          make_trajectory_c(c, c->rp_x0[iray], c->rp_z0[iray], c->rp_theta0[iray], c->rp_s0[iray]);
          for (inu = 1; inu <= nfr; inu++) {
            im[RECIDX(inu, ix, iy)] = charintline(c, iline, inu, iray, 0.0);
            if (ta) ta[RECIDX(inu, ix, iy)] = c->char_tau;
          }
        } else {
          for (inu = 1; inu <= nfr; inu++) {
            im[RECIDX(inu, ix, iy)] = (c->out_itype == 3) ? c->isrf_line[inu] : 0.0;
            if (ta) ta[RECIDX(inu, ix, iy)] = 0.0;
          }
        }
        iray = iray + 1;
      }
    if (c->imrec_addstar > 0 && c->imrec_starunres > 0) { /* telescope.F:2153-2200 */
      const double srat = (3.14159265 * c->rstar * c->rstar) / (4.0 * c->rays_sizepix_x * c->rays_sizepix_y);
      const double srat1 = 1.0 - srat;
      make_trajectory_c(c, c->rp_x0[iray], c->rp_z0[iray], c->rp_theta0[iray], c->rp_s0[iray]);
      for (inu = 1; inu <= nfr; inu++) {
        double dummy = charintline(c, iline, inu, iray, 0.0);
        dummy = dummy * srat;
        ix = nrx / 2;
        iy = nry / 2;
        im[RECIDX(inu, ix, iy)] = dummy + srat1 * im[RECIDX(inu, ix, iy)];
        im[RECIDX(inu, ix + 1, iy)] = dummy + srat1 * im[RECIDX(inu, ix + 1, iy)];
        im[RECIDX(inu, ix, iy + 1)] = dummy + srat1 * im[RECIDX(inu, ix, iy + 1)];
        im[RECIDX(inu, ix + 1, iy + 1)] = dummy + srat1 * im[RECIDX(inu, ix + 1, iy + 1)];
      }
    }
    if (maserflag) maserflag[l] = c->maserflag;
  }
  c->jb_armed = 0;
  return 0;
}

static int check_ready(orc_ctx *c) {
  if (!c->rc || !c->rho || !c->nlines || !c->have_line_dust || !c->cam_set || !c->bc_set) {
    snprintf(c->err, sizeof c->err,
             "stop 13: render before set_grid/medium/lines/dust/camera/bc (rays_ready.ne.321)");
    return 13;
  }
  return 0;
}

int orc_render(orc_ctx *c, int iline0, int nl, int nfr, double vmax_kms, double dist_cm,
               double *flux, double *imcir, int *cmask, double *tau_center, int *maserflag,
               double *velo) {
  int code, l;
  if ((code = check_ready(c)) != 0) return code;
  if (iline0 < 1 || iline0 + nl - 1 > c->nlines) {
    snprintf(c->err, sizeof c->err, "render: line range out of bounds");
    return 13;
  }
  /* telescope.F:4323-4340 telescope_check_safety_numbers */
  for (l = 1; l <= c->nr - 1; l++)
    if (RC(l + 1) / RC(l) - 1.0 < 1.e4 * TELESC_EPS) {
      snprintf(c->err, sizeof c->err, "stop 13: radial grid too fine for TELESC_EPS");
      return 13;
    }
  c->jb_armed = 1;
  if ((code = setjmp(c->jb)) != 0) { c->jb_armed = 0; return code; }
  for (l = 0; l < nl; l++) {
    size_t nim = (size_t)(c->rays_nrr + 1) * c->rays_nrphi * nfr;
    render_line(c, iline0 + l, nfr, vmax_kms, dist_cm, flux + (size_t)l * nfr,
                imcir ? imcir + (size_t)l * nim : 0, cmask ? cmask + (size_t)l * nim : 0,
                tau_center ? tau_center + l : 0, maserflag ? maserflag + l : 0,
                velo ? velo + (size_t)l * nfr : 0);
  }
  c->jb_armed = 0;
  return 0;
}

void orc_get_counters(const orc_ctx *c, double *R, double *E, double *S) {
  if (R) *R = c->cnt_R;
  if (E) *E = c->cnt_E;
  if (S) *S = c->cnt_S;
}
void orc_reset_counters(orc_ctx *c) { c->cnt_R = c->cnt_E = c->cnt_S = 0.0; }
void orc_set_ring_sample(orc_ctx *c, int lo, int hi, int stride) {
  c->ring_lo = lo;
  c->ring_hi = hi;
  c->ring_stride = stride > 0 ? stride : 1;
}

int orc_max_nodes(const orc_ctx *c) { return c->raysize; }

int orc_trajectory(orc_ctx *c, int iray, double *s, double *radius, double *theta, double *mu,
                   double *phi, int *icross, int *iradius, int *itheta) {
  int code, i;
  if (!c->cam_set || iray < 1 || iray > c->rays_amount) return -13;
  c->jb_armed = 1;
  if ((code = setjmp(c->jb)) != 0) { c->jb_armed = 0; return -code; }
  make_trajectory_c(c, c->rp_x0[iray], c->rp_z0[iray], c->rp_theta0[iray], c->rp_s0[iray]);
  c->jb_armed = 0;
  for (i = 1; i <= c->tr_amount; i++) {
    if (s) s[i - 1] = c->tr_s[i];
    if (radius) radius[i - 1] = c->tr_radius[i];
    if (theta) theta[i - 1] = c->tr_theta[i];
    if (mu) mu[i - 1] = c->tr_mu[i];
    if (phi) phi[i - 1] = c->tr_phi[i];
    if (icross) icross[i - 1] = c->tr_icross[i];
    if (iradius) iradius[i - 1] = c->tr_iradius[i];
    if (itheta) itheta[i - 1] = c->tr_itheta[i];
  }
  return c->tr_amount;
}

int orc_node_values(orc_ctx *c, int iray, int iline, double *srcd, double *alpd, double *dvmu,
                    double *lw, double *nup, double *ndown) {
  int code, i;
  if (!c->cam_set || iray < 1 || iray > c->rays_amount) return -13;
  c->jb_armed = 1;
  if ((code = setjmp(c->jb)) != 0) { c->jb_armed = 0; return -code; }
  make_trajectory_c(c, c->rp_x0[iray], c->rp_z0[iray], c->rp_theta0[iray], c->rp_s0[iray]);
  for (i = 1; i <= c->tr_amount; i++) {
    int ir = c->tr_iradius[i], it = c->tr_itheta[i];
    double dr = (c->tr_radius[i] - RC(ir)) / (RC(ir + 1) - RC(ir));
    double dt = (c->tr_theta[i] - TC(it)) / (TC(it + 1) - TC(it));
    double a, b, d, e, f, g;
    get_line_dust_values(c, c->tr_icross[i], ir, it, dr, dt, c->tr_mu[i], c->tr_phi[i], iline, &a,
                         &b, &d, &e, &f, &g);
    if (srcd) srcd[i - 1] = a;
    if (alpd) alpd[i - 1] = b;
    if (dvmu) dvmu[i - 1] = d;
    if (lw) lw[i - 1] = e;
    if (nup) nup[i - 1] = f;
    if (ndown) ndown[i - 1] = g;
  }
  c->jb_armed = 0;
  return c->tr_amount;
}

/*
 * radlite_oracle.h -- CPU restatement of RADLite's line ray-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (pontoppi/radlite, fixed-form Fortran 77) cannot
 * be compiled in this image (no Fortran front end) and ships no tests, golden
 * vectors or expected outputs for this path.  This file restates the reference
 * routines line by line (each function cites the file:line it follows); it is
 * checked against analytic known-answer tests, not against reference output.
 *
 * The API deliberately has the same shape as include/radlite_b200.h so the
 * parity tests can drive the oracle and the CUDA library with identical calls.
 * All arrays are host memory, C order, 0-based; "cell" arrays are [ir][it] with
 * it fastest over the STORED (upper) hemisphere, i.e. the Fortran (it,ir) order.
 */
#ifndef RADLITE_ORACLE_H
#define RADLITE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

int  orc_create(orc_ctx **out);
void orc_destroy(orc_ctx *c);
const char *orc_last_error(const orc_ctx *c);

/* grid.F:711-819 (radius.inp branch), 1098-1210 (theta.inp branch, mirror flag 1):
 * r[nr] in cm, theta[nth] = upper hemisphere in rad; ghost cells are built here. */
int orc_set_grid(orc_ctx *c, int nr, int nth, const double *r, const double *theta);
/* same, from the Fortran host's rsi_x_c(-1:nr+2,1) and rsi_x_c(-1:nt+2,2) as they are */
int orc_set_grid_ghosted(orc_ctx *c, int nr, int nt, const double *rc_m1, const double *tc_m1);

/* setup.F:1453 (density), 864 (abundance), 754 (velocity, cm/s), 803 (turbulence a-width km/s),
 * line.F:142 (umass_av).  rho/abund/linewidth: [nr][nth]; vel: [nr][nth][3] */
int orc_set_medium(orc_ctx *c, const double *rho, const double *abund, const double *vel,
                   const double *linewidth, double umass_av);

/* line.F:1826-1985 (moldata), 811-950 (levelpop).  lev_up/lev_down are 1-based level numbers,
 * popul: [nr][nth][nlevels] fractional populations. */
int orc_set_lines(orc_ctx *c, int nlines, int nlevels, const int *lev_up, const int *lev_down,
                  const double *linefreq, const double *aud, const double *gdeg,
                  const double *popul);

/* line.F:3502-3743 (dust source term at line centre), setup.F:937 (bplanck).
 * nsize[nspec]; kappa_abs/kappa_scat: [nspec][maxsize][ncf]; dust_rho: [nr][nth][nspec];
 * dust_temp: [nr][nth][nspec][maxsize]; scati_src: [nr][nth][ncf] or NULL (=0). */
int orc_set_dust(orc_ctx *c, int nspec, const int *nsize, int ncf, const double *cont_freq_nu,
                 const double *kappa_abs, const double *kappa_scat, const double *dust_rho,
                 const double *dust_temp, const double *scati_src);
/* alternative: the host already holds line_dust_src/alp(1,iline,it,ir): [nlines][nr][nth] */
int orc_set_line_dust(orc_ctx *c, const double *src, const double *alp);

/* telescope.F:715-1191 setup_rays_circular(1,nr,1,anginf,nphi,nrext,dbdr,rstar,imethod,nrref) */
int orc_set_camera(orc_ctx *c, double anginf, int nphi, int nrext, int dbdr, double rstar,
                   int imethod, int nrref);

/* common_boundary.h: iradbnd_in_itype / iradbnd_out_itype; star.F:449-528 (starspec_cont =
 * surface intensity on cont_freq_nu), star.F:675 (isrf_cont, may be NULL) */
int orc_set_bc(orc_ctx *c, int in_itype, int out_itype, int ncf, const double *cont_freq_nu,
               const double *starspec_cont, const double *isrf_cont);

/* configure.h:8 SUBGRID, :53 NONREDUNDANT, :52 LEVTHRES; aksmax<0 => line.F:2968-3033 */
int orc_set_options(orc_ctx *c, int subgrid, int nonredundant, double levthres, double aksmax);

int orc_get_camera_dims(orc_ctx *c, int *nrr, int *nphi, int *nray);

/* main.F:1043-1049 loop body for lines iline0..iline0+nl-1 (1-based):
 * calc_line_spectrum (telescope.F:1465) -> calc_freq_flux_observer (:1320).
 * flux: [nl][nfr]; imcir (opt): [nl][nrr+1][nphi][nfr]; cmask (opt) same shape, int;
 * tau_center (opt): [nl]; maserflag (opt): [nl]; velo (opt): [nl][nfr] = line_dnu/nu0 */
int orc_render(orc_ctx *c, int iline0, int nl, int nfr, double vmax_kms, double dist_cm,
               double *flux, double *imcir, int *cmask, double *tau_center, int *maserflag,
               double *velo);

/* Rectangular imager / position-velocity cube (linespectrum.inp command 2): telescope.F:2229-2475
 * setup_rays_rectang(nrx,nry,sizepix_x,sizepix_y,anginf,phioffset,xoffset,yoffset) with imrec_addstar, and
 * telescope.F:2061-2227 make_freq_image_rectang for each line (driven by calc_write_line_posvel :1828).
 * image, tau: [nl][nx][ny][nfr] = imrec_int(inu,ix,iy), imrec_tau(inu,ix,iy) (tau may be NULL). */
int orc_set_camera_rect(orc_ctx *c, double anginf, int nx, int ny, double sizepix_x, double sizepix_y,
                        double phioffset, double xoffset, double yoffset, double rstar, int addstar);
int orc_render_rect(orc_ctx *c, int iline0, int nl, int nfr, double vmax_kms, double *image, double *tau,
                    int *maserflag);

/* work counters accumulated over render calls: R = calls of charintline,
 * E = calls of integrate_element_linedust, S = segments visited (sum over charintline calls) */
void orc_get_counters(const orc_ctx *c, double *R, double *E, double *S);
void orc_reset_counters(orc_ctx *c);
/* benchmarking only: trace rings lo, lo+stride, .. <= hi (1-based) and leave the others zero; hi=0 = all.
 * Gives a bounded sample of a workload for bench.py's cpu_baseline; never used by parity tests. */
void orc_set_ring_sample(orc_ctx *c, int lo, int hi, int stride);

/* camera tables after set_camera+set_grid: rays_r[0..nrr], imcir_ri[0..nrr+1] (telescope.F:443-488) */
int orc_get_rings(orc_ctx *c, double *rays_r, double *imcir_ri);

/* diagnostics for node-level diffing (mirrors the MONITOR_CHARINT idea, telescope.F:408-418):
 * trajectory of ray iray (1-based; 1 = centre).  Arrays sized >= orc_max_nodes(). Returns count. */
int orc_max_nodes(const orc_ctx *c);
int orc_trajectory(orc_ctx *c, int iray, double *s, double *radius, double *theta, double *mu,
                   double *phi, int *icross, int *iradius, int *itheta);
/* per-node interpolated values for line iline along ray iray (get_line_dust_values, line.F:3965) */
int orc_node_values(orc_ctx *c, int iray, int iline, double *srcd, double *alpd, double *dvmu,
                    double *lw, double *nup, double *ndown);

/* small pieces exposed for unit tests */
double orc_qdr_src_2(double inten, double js1, double alp1, double js2, double alp2, double ds);
double orc_bplanck(double temp, double nu);
void   orc_hunt(const double *xx, int n, double x, int *jlo); /* xx[0..n-1] <-> xx(1..n) */

#ifdef __cplusplus
}
#endif
#endif

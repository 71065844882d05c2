/*
 * radlite_b200.h -- C ABI of the B200-native RADLite line ray-tracer (libradlite_b200.so).
 *
 * The reference (pontoppi/radlite, Fortran 77) has no FFI for this path: everything is
 * `call` + COMMON blocks.  The seam this ABI replaces is the body of the line loop
 *   main.F:1043-1049   do iln ... call calc_line_spectrum(iln,...)
 *   telescope.F:1465   calc_line_spectrum  -> :1320 calc_freq_flux_observer
 *   telescope.F:339    make_image_circular -> :2787 make_trajectory_c, :3889 charintline
 * i.e. "given the COMMON state, fill imcir_int / imcir_cmask / spec_flux_observer for line
 * iline", widened to a batch of lines.  Each entry point below names the reference routine /
 * COMMON arrays it takes over.  INTEGRATION.md shows the iso_c_binding interface module and the
 * patch to telescope.F/main.F that binds them.
 *
 * Conventions
 *  - every pointer is HOST memory owned by the caller; the library copies during the call.
 *  - arrays are dense, C order, 0-based.  "cell" arrays are [ir][it] with it fastest over the
 *    stored (upper) hemisphere, i.e. exactly the memory order of the Fortran (it,ir) arrays once
 *    the unused padding of the COMPILE-TIME dimensions is squeezed out (the Fortran shim packs).
 *  - return value 0 = ok; otherwise the reference's `stop` code where one exists (13, 749, 393,
 *    6023/6024, 124, 137, 7454, 83991/91991, 192, 987/988 ...) or 1 for a bare `stop`;
 *    negative values are CUDA/runtime failures.  rl_last_error() gives the message; the Fortran
 *    shim prints it and STOPs with the same code, so drivers see the same behaviour
 *    (missing radlite.success).
 *  - one ctx per process and per GPU, calls serialised by the caller (the reference is
 *    single-threaded and not re-entrant).
 *  - there is NO CPU fallback: rl_create fails if no sm_100 device is usable.
 */
#ifndef RADLITE_B200_H
#define RADLITE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rl_ctx rl_ctx;

/* device = CUDA ordinal (one process per GPU: pass LOCAL_RANK) */
int  rl_create(rl_ctx **out, int device);
void rl_destroy(rl_ctx *ctx);
const char *rl_last_error(const rl_ctx *ctx);

/* grid.F:711-819 + 1098-1210 (create_spacegrid, radius.inp/theta.inp branches, mirror flag 1):
 * r[nr] cm, theta[nth] rad (upper hemisphere); ghost cells rsi_x_c(-1:0), (n+1:n+2) and the
 * theta mirror are built inside with the reference's literals.  Also does interpol.F:87-100
 * (make_index -> ridx_it). */
int rl_set_grid(rl_ctx *ctx, int nr, int nth, const double *r, const double *theta);
/* the Fortran host passes rsi_x_c(-1,1) and rsi_x_c(-1,2) as they are (common_grid.h:17);
 * nt = irsi_frsizey (both hemispheres). */
int rl_set_grid_ghosted(rl_ctx *ctx, int nr, int nt, const double *rc_m1, const double *tc_m1);

/* COMMON /mediumarr/ medium_arr_rho, medium_arr_velocity (common_setup.h:25-29),
 * /localabundance/ locabun_abund_mol (common_lines.h:133), locprof_linewidth(1,it,ir)
 * (setup.F:803-850: line independent), umass_av (line.F:142).
 * rho, abund, linewidth: [nr][nth]; vel: [nr][nth][3] (v_r, v_theta, v_phi) cm/s. */
int rl_set_medium(rl_ctx *ctx, const double *rho, const double *abund, const double *vel,
                  const double *linewidth, double umass_av);

/* COMMON lev_up, lev_down, linefreq, Aud_const, gdeg (line.F:1826-1985 read_species_lambda)
 * and line_level_popul(lev,it,ir) (line.F:811-950 read_levelpopul): popul [nr][nth][nlevels].
 * Takes over line.F:1600-1673 global_prepare_transitions / 1708-1788 prepare_lines (Bud,Bdu). */
int rl_set_lines(rl_ctx *ctx, int nlines, int nlevels, const int *lev_up, const int *lev_down,
                 const double *linefreq, const double *aud, const double *gdeg,
                 const double *popul);

/* Takes over line.F:3687-3743 global_prepare_line_dust / 3502-3608 line_dust_compute_src_alp
 * and setup.F:937 bplanck (the source.F dust source term).  COMMON dust_rho(spec,it,ir),
 * dust_temp(size,spec,it,ir), dust_kappawgt_abs/scat(freq,1,size,spec), cont_freq_nu,
 * scati_src(freq,it,ir) (source.F:494 read_scatsource; NULL = all zero).
 * nsize[nspec]; kappa_*: [nspec][maxsize][ncf]; dust_rho: [nr][nth][nspec];
 * dust_temp: [nr][nth][nspec][maxsize]; scati_src: [nr][nth][ncf]. Call after rl_set_lines. */
int rl_set_dust(rl_ctx *ctx, int nspec, const int *nsize, int ncf, const double *cont_freq_nu,
                const double *kappa_abs, const double *kappa_scat, const double *dust_rho,
                const double *dust_temp, const double *scati_src);
/* alternative for a host that already ran global_prepare_line_dust:
 * line_dust_src(1,iline,it,ir), line_dust_alp(1,iline,it,ir) as [nlines][nr][nth] */
int rl_set_line_dust(rl_ctx *ctx, const double *src, const double *alp);

/* telescope.F:715-1191 setup_rays_circular(1,nr,1,anginf,nrphiinf,nrext,dbdr,rstar,imethod,nrref)
 * (call site main.F:792; imethod=1, nrref=10 are hard-wired at main.F:37-38) and the ring
 * edges of telescope.F:443-488. */
int rl_set_camera(rl_ctx *ctx, double anginf, int nphi, int nrext, int dbdr, double rstar,
                  int imethod, int nrref);

/* iradbnd_in_itype / iradbnd_out_itype (common_boundary.h), radbnd_cont_starspec on
 * cont_freq_nu (star.F:449-528; surface intensity), radbnd_cont_interstellfield (star.F:675;
 * NULL unless out_itype==3).  Takes over line.F:3797-3845 / 3855-3903 per line. */
int rl_set_bc(rl_ctx *ctx, int in_itype, int out_itype, int ncf, const double *cont_freq_nu,
              const double *starspec_cont, const double *isrf_cont);

/* configure.h:8 SUBGRID, :53 NONREDUNDANT, :52 LEVTHRES as run-time switches; aksmax from
 * line.F:2968-3033 line_velo_prep_profiles_abun (pass <0 to have it computed). */
int rl_set_options(rl_ctx *ctx, int subgrid, int nonredundant, double levthres, double aksmax);

/* rays_nrr, rays_nrphi, rays_amount (common_telescope.h:20-23) */
int rl_get_camera_dims(rl_ctx *ctx, int *nrr, int *nphi, int *nray);
/* rays_r(0:nrr), imcir_ri(0:nrr+1) */
int rl_get_rings(rl_ctx *ctx, double *rays_r, double *imcir_ri);

/* main.F:1043-1049 for lines iline0 .. iline0+nl-1 (1-based):
 *   nfr, vmax_kms = main_passbnfr, main_passbwidth (main.F:209-211); dist_cm (main.F:213)
 *   flux       [nl][nfr]               spec_flux_observer(inu)           (required)
 *   imcir      [nl][nrr+1][nphi][nfr]  imcir_int(inu,iphi,ir)            (NULL = not wanted)
 *   cmask      [nl][nrr+1][nphi][nfr]  imcir_cmask (accumulates over lines like the reference)
 *   tau_center [nl]                    char_tau_center
 *   maserflag  [nl]                    maserflag (telescope.F:4295)
 *   velo       [nl][nfr]               line_dnu(inu,iline)/line_nu0(iline)
 * Channel order is the reference's inu = 1..nfr (ascending frequency). */
int rl_render(rl_ctx *ctx, int iline0, int nl, int nfr, double vmax_kms, double dist_cm,
              double *flux, double *imcir, int *cmask, double *tau_center, int *maserflag,
              double *velo);

/* ---- multi-GPU by camera-ring block (single-line configs) -----------------
 * The ring loop `do ir=1,nrr` of make_image_circular (telescope.F:532) and the ring sum of
 * calc_freq_flux_observer (telescope.F:1388-1433) split over ranks: rl_render_rings traces only
 * the camera rings ring_lo..ring_hi (0 = the central beam telescope.F:498-527, 1..nrr) and returns
 *   ringsum [nl][nrr+1][nfr]   row 0 = pi ri(1)^2 I_centre (telescope.F:1393-1396), row ir = the
 *                              `dslum` of ring ir (telescope.F:1418-1423); rows outside the block = 0
 *   imcir   (optional) same shape as in rl_render; only the rows of the block are written.
 * The caller adds the ringsum arrays of all ranks (disjoint blocks: x + 0 is exact) and hands the
 * total to rl_flux_from_rings, which does the reference's index-ordered sum `slum` and the division
 * by distance^2 -- the result is bit-identical to rl_render on one GPU. */
int rl_render_rings(rl_ctx *ctx, int iline0, int nl, int nfr, double vmax_kms, double dist_cm,
                    int ring_lo, int ring_hi, double *ringsum, double *imcir);
int rl_flux_from_rings(rl_ctx *ctx, int nl, int nfr, double dist_cm, const double *ringsum,
                       double *flux);
/* the same with everything a cube run writes (SAVE_IMCIR, telescope.F:1346-1371): the rows of the block in imcir
 * and cmask (same shapes as in rl_render; rows outside the block are left untouched), tau_center and maserflag
 * (tau_center is only set by the block that holds ring 0; maserflag is the block's own: OR them) */
int rl_render_rings_cube(rl_ctx *ctx, int iline0, int nl, int nfr, double vmax_kms, double dist_cm, int ring_lo,
                         int ring_hi, double *ringsum, double *imcir, int *cmask, double *tau_center,
                         int *maserflag);

/* Device-resident variants for sharded renders (one process per GPU): rl_render_rings_device leaves nothing
 * on the host and copies the ring sums [nl][nrr+1][nfr] into d_ringsum, a DEVICE pointer of the caller (the
 * buffer the ranks hand to their collective: the blocks are disjoint and the other rows exactly 0, so a sum
 * reduction is a concatenation); when it returns the buffer is complete (the library's stream is synchronised).
 * kernel_ms as in rl_render_device (may be NULL).  rl_flux_from_rings_device does the index-ordered ring sum
 * of telescope.F:1388-1433 on the reduced buffer and returns the spectra to the host.
 * A ring-block render builds the ray geometry of its own rings only. */
int rl_render_rings_device(rl_ctx *ctx, int iline0, int nl, int nfr, double vmax_kms, double dist_cm,
                           int ring_lo, int ring_hi, double *d_ringsum, float *kernel_ms);
int rl_flux_from_rings_device(rl_ctx *ctx, int nl, int nfr, double dist_cm, const double *d_ringsum,
                              double *flux);
/* Work estimate per camera ring [nrr+1] of rendering lines iline0..iline0+nl-1 (geometry, channel selection and
 * the integrate kernels' node steps; nothing is integrated): the weights for cutting the rings into blocks of
 * equal work, radlite_b200.shard.split_rings.  The reference's drivers balance by line count only
 * (radlite.py:1163-1169). */
int rl_plan_costs(rl_ctx *ctx, int iline0, int nl, int nfr, double vmax_kms, double *ring_cost);

/* ---- rectangular imager / position-velocity cube (linespectrum.inp command 2) ---------
 * rl_set_camera_rect = telescope.F:2229-2475 setup_rays_rectang(nrx,nry,sizepix_x,sizepix_y,anginf,phioffset,
 * xoffset,yoffset) (call site main.F:800) with imrec_addstar of linespectrum.inp (telescope.F:216): nx, ny even,
 * pixel sizes and offsets in cm, angles in rad.  rl_render_rect = the rendering part of calc_write_line_posvel
 * (telescope.F:1828) -> make_freq_image_rectang (:2061): every pixel at every channel (no NONREDUNDANT), rays
 * that hit the star take its intensity (:4194-4208), pixels off the model take the outer boundary value, and an
 * unresolved star is smeared over the four central pixels (:2153-2200).
 *   image, tau [nl][nx][ny][nfr]   imrec_int(inu,ix,iy), imrec_tau(inu,ix,iy) = char_tau (tau may be NULL)
 * The circular camera (rl_set_camera) and its renders are unaffected. */
int rl_set_camera_rect(rl_ctx *ctx, double anginf, int nx, int ny, double sizepix_x, double sizepix_y,
                       double phioffset, double xoffset, double yoffset, double rstar, int addstar);
int rl_render_rect(rl_ctx *ctx, int iline0, int nl, int nfr, double vmax_kms, double *image, double *tau,
                   int *maserflag);

/* ---- driver-side ends of the path: no ASCII intermediates for many-line spectra ----------
 * rl_set_lines_lte replaces rl_set_lines' `popul` by its recipe: LTE populations g exp(-E h c / k T) / Q(T)
 * (pyradlite radlite.py:1111-1119; PRO/make_levelpop.pro), Q interpolated linearly (extrapolating) in the
 * tabulated partition sum psum(psum_temp) (radlite.py:1147-1153), values below 1e-99 flushed to 0 -- computed
 * on the device straight into the table the ray tracer reads (the levelpop_<mol>.dat of BASELINE configs[4] is
 * 4 GB).  energy_cm[nlevels] in cm^-1, tgas [nr][nth] in K. */
int rl_set_lines_lte(rl_ctx *ctx, int nlines, int nlevels, const int *lev_up, const int *lev_down,
                     const double *linefreq, const double *aud, const double *gdeg, const double *energy_cm,
                     const double *tgas, int npsum, const double *psum_temp, const double *psum);
/* rl_synthesize_spectrum: pyradlite's RadliteSpectrum._process_spectrum (radlite.py:3001-3184; PRO/genspec.pro),
 * interpolation 'linear': vel, flux [nl][nfr] as linespectrum_<mol>.dat lists them (velocity ascending, km/s;
 * F_nu at 1 pc), freq [nl] line centres in Hz, dist_pc, obsres and vsampling in km/s.  Outputs on the
 * wavelength grid [um] of rl_synthesis_size's nout points: line + continuum, line emission only, continuum
 * [Jy at dist_pc]. */
int rl_synthesis_size(int nl, int nfr, const double *vel, const double *freq, double obsres, double vsampling,
                      int *nout, int *nfull);
int rl_synthesize_spectrum(rl_ctx *ctx, int nl, int nfr, const double *vel, const double *flux, const double *freq,
                           double dist_pc, double obsres, double vsampling, double *wavelength, double *spectrum,
                           double *emission, double *continuum);

/* work counters since the last reset: R = ray-channel integrations (charintline calls the
 * reference would make), E = element integrations (integrate_element_linedust calls incl.
 * sub-grid steps), S = ray segments visited. */
void rl_get_counters(const rl_ctx *ctx, double *R, double *E, double *S);
void rl_reset_counters(rl_ctx *ctx);
/* Opaque-wall start (no counterpart in the reference, which integrates every segment of a ray,
 * telescope.F:4079-4300): the segments of a ray that lie, seen from the observer, behind so much dust optical
 * depth that their contribution is provably below exp(-tau) of what the dust in front of them emits -- for
 * every line of the batch; criterion in DESIGN.md 4.3: upper bound of the dropped part against a lower bound
 * of the kept part -- are not integrated.  Default 64 (a relative 1e-28: twelve orders of magnitude below the
 * rounding of the result); a batch with an inverted level pair or a negative dust opacity, and a ray whose
 * front crosses a cell without emission, are never shortened; 0 integrates every segment.  R, E, S above
 * keep counting the reference's work; rl_get_executed returns the element integrations this library
 * actually performed. */
int rl_set_wall_tau(rl_ctx *ctx, double tau);
/* Integrate kernel: 0 = chosen by regime (ztile_kernel + zcont_kernel from 8 lines per batch, chan_kernel
 * below), 1 = ztile_kernel, 2 = tile_kernel, 3 = chan_kernel.  ztile_kernel and chan_kernel give the same bits,
 * tile_kernel agrees with them to ~1e-13; sharded renders that must be bit-identical to an unsharded one pin
 * the kernel. */
int rl_set_kernel(rl_ctx *ctx, int mode);
double rl_get_executed(const rl_ctx *ctx);

/* ---- device-resident variant used by bench.py's kernel-only timing ----------------------
 * Same work as rl_render but inputs stay resident and nothing is copied back: the result stays
 * in device memory until rl_fetch_flux.  kernel_ms[5] (optional) receives CUDA-event times [ms]
 * on the library's stream: {geometry, per-line preparation + channel selection, the ray-integration kernel,
 * continuum fill + flux reduction, whole call}. */
int rl_render_device(rl_ctx *ctx, int iline0, int nl, int nfr, double vmax_kms, double dist_cm,
                     float *kernel_ms);
int rl_fetch_flux(rl_ctx *ctx, int nl, int nfr, double *flux);
/* forget the cached ray geometry so the next render rebuilds it (the reference rebuilds it for
 * every ray of every line, telescope.F:500,538) */
void rl_invalidate_geometry(rl_ctx *ctx);
/* measured FP64 FMA throughput of this GPU [TFLOP/s] (DFMA chains, best of 4): the roofline
 * denominator bench.py reports against (MEASURED_PEAKS.json carries no FP64 figure) */
int rl_fp64_peak(rl_ctx *ctx, double *tflops);
/* number of kernels this library launched since rl_create (bench.py "gpu_launches") */
long long rl_launch_count(const rl_ctx *ctx);
/* diagnostics: nodes of ray iray (1-based) as built on the device; arrays sized rl_max_nodes() */
int rl_max_nodes(const rl_ctx *ctx);
int rl_get_ray_nodes(rl_ctx *ctx, int iray, double *ds, double *dvmu, double *lw, double *wr,
                     double *wt, int *cells4, int *flags);
long long rl_total_nodes(const rl_ctx *ctx);
/* diagnostics: copy a named device buffer of the last render batch to the host ("ztiles", "nstart",
 * "node_off", "rng", "nitems", "zlines"); returns the bytes the buffer holds (-1: unknown name) and copies
 * at most nbytes of them */
long long rl_debug_fetch(rl_ctx *ctx, const char *what, void *out, long long nbytes);

#ifdef __cplusplus
}
#endif
#endif

// rl_types.h -- shared host/device types of libradlite_b200 (internal, not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rl {

// reference compile-time constants (configure.h:46, main.h:452-453, common_telescope.h:4-6)
constexpr double kTelescEps = 1.0e-10;
constexpr double kPi = 3.1415926535897932385;
constexpr double kTempCmb = 2.728;
constexpr int kRayAdpt = 4;
constexpr int kRayRnpt = 4;
constexpr int kMaxExtra = 40;  // 2 + 4*RAYADPT*2 = 34 used
constexpr int kLgNrMax = 31;   // line.F:4657-4661
constexpr int kTileIpt = 1;        // ray-channel items per tile_kernel thread
constexpr int kTileChunk = 31;     // nodes staged per chunk (upper bound: chunk + 1 previous node = one per producer lane)
#ifndef RL_GEOM_WARP_MAX
#define RL_GEOM_WARP_MAX (1 << 30)
#endif
constexpr int kGeomWarpMax = RL_GEOM_WARP_MAX;  // rays of a build below which geom_kernel runs one warp per ray
constexpr int kSpanThreads = 128;  // span_kernel block = lines per batch upper bound (one mask bit per line)

// exp table of the integrate kernel: 2^(j / kTabN), j = 0..kTabN-1.  16 entries of 8 bytes span the 32
// shared-memory banks exactly once, so the per-lane lookups never conflict (a 256-entry table cost ~5.6
// data-pipe wavefronts per lookup on B200; this one costs 2); the polynomial is two or three terms longer.
#ifndef RL_TAB_BITS
#define RL_TAB_BITS 4
#endif
constexpr int kTabBits = RL_TAB_BITS;
constexpr int kTabN = 1 << kTabBits;
#if RL_TAB_BITS == 4
constexpr double kTabSqrtScale = 4.804489635145799;  // sqrt(kTabN / ln 2)
constexpr unsigned kHiUmax = 0x40564f52u;             // hi word of sqrt(345 kTabN / ln 2): exp(-345) ~ 1e-150
constexpr int kDegGauss = 5, kDegTau = 7;             // |r| <= ln2/32: r^6/720 < 1.5e-13, r^8/40320 < 2e-18
#elif RL_TAB_BITS == 5
constexpr double kTabSqrtScale = 6.7945744023041525;
constexpr unsigned kHiUmax = 0x405f8d08u;
constexpr int kDegGauss = 5, kDegTau = 6;
#elif RL_TAB_BITS == 8
constexpr double kTabSqrtScale = 19.217958540583197;
constexpr unsigned kHiUmax = 0x40764f52u;
constexpr int kDegGauss = 3, kDegTau = 4;             // |r| <= ln2/512: r^4/24 < 1.5e-13, r^5/120 < 4e-17
#else
#error "RL_TAB_BITS must be 4, 5 or 8"
#endif

// node flags (low 2 bits = tr_icross: 1 = R crossing, 2 = theta crossing, 3 = extra point)
constexpr uint32_t kFlagIcrMask = 3u;
constexpr uint32_t kFlagInit = 4u;   // carried profile state is reset before this segment
constexpr uint32_t kFlagStar = 8u;   // centre beam: mix in the star before this segment
constexpr uint32_t kFlagZero = 16u;  // inner BC type 1: intensity := 0 before this segment
constexpr uint32_t kFlagSub = 32u;   // 6 q > 1: the segment ending here may be velocity sub-gridded (line.F:4715)

// ghosted grids: rc[i+1] = rsi_x_c(i,1), i=-1..nr+2 ; tc[i+1] = rsi_x_c(i,2), i=-1..nt+2 ;
// ridx[i+4] = ridx_it(i), i=-4..nt+4
struct GridDev {
  int nr, nt, nth;
  const double *rc;
  const double *tc;
  const int *ridx;
};

// Line-independent per-cell fields, one 32-byte record per cell: {linewidth, v_r, v_theta, v_phi}
// Per-line per-cell fields, one 32-byte record: {src_dust, alp_dust, N_up, N_down}
// Cell index = (ir-1)*nth + (it-1): the Fortran (it,ir) order.

// Node lists of all rays: one 64-byte record per node, ray r owns [node_off[r], node_off[r+1]).
// The record is read as four 16-byte loads; all lanes working on the same ray read the same
// address (broadcast), so an array of structures is the right layout here.
constexpr int kCellFlagShift = 26;                 // cells.x = cell(t0,r0) | flags << 26
constexpr int kCellMask = (1 << kCellFlagShift) - 1;
struct __align__(16) NodeRec {
  double ds;    // segment length to the previous node (vacuum rule applied); 0 for a ray's first node
  double dvmu;  // Omega.v/c at the node (line independent)
  double lw;    // interpolated line width [km/s]
  double inv_lwav;  // 1 / (0.5 (lw + lw_prev)): reciprocal mean width of the segment ending here (1/lw for node 0)
  double wr;    // dr
  double wt;    // dt
  int4 cells;   // cells (t0,r0)|flags<<26, (t1,r0), (t0,r1), (t1,r1)
};
struct NodesDev {
  NodeRec *rec;
};

struct GeomParams {
  GridDev g;
  int nray;
  int ray_lo, ray_hi;  // rays built by this call (a camera-ring block); the others get no nodes
  int rect;            // rectangular camera (telescope.F:2229): every ray is traced with rbeam0 = 0 (a ray whose
                       // impact parameter is inside the star takes the stellar intensity, telescope.F:4194-4208)
  const double *rb;    // [nray] rectangular camera: pixel radius rp_b; pixels at or beyond bskip are not traced
  double bskip;        // 0.999 R_nr (telescope.F:2127)
  const double *tan2;  // [nt/2 + 1] tan^2 of the cone angles
  double4 *thr;        // [rays of the block][nt/2] cone roots and their radii (roots_kernel)
  double2 *rrt;        // [rays of the block][nr+1] sphere roots (roots_kernel)
  const double *x0;  // [nray]
  const double *z0;
  double theta0;
  double costh0, sinth0;  // cos / sin of theta0, evaluated once on the host (the reference's libm)
  double rstar;
  int in_itype;
  double rbeam0_center;  // imcir_ri(1): star mixing of the centre ray
  const double4 *cellS;  // {lw, v1, v2, v3}
  const long long *node_off;  // [nray+1] (fill pass)
  int *node_cnt;              // [nray]   (count pass)
  NodesDev nodes;
  void *light;      // [total nodes] 32-byte LightNode scratch of the fill pass (rl_geom.cu)
  long long ntot;   // total nodes (fill pass)
  int *status;  // first error code (reference stop code), 0 = ok
};

// per-line constants (host-prepared, line.F:427-545, 1708-1788, 3797-3903)
struct LineDev {
  double nu0;       // linefreq = line_nu0
  double aud;       // Aud
  double bud, bdu;  // Einstein B's
  double dnu0;      // line_dnu(1)
  double ddnu;      // channel spacing dnu
  double i_outer;   // outer-BC start intensity for out_itype 0/2 (type 3: per channel array)
  // derived constants (host): line.F:2301 aa = k_aa * width ; line.F:4571-4588 j_l = c_src N_up phi,
  // alpha_l = c_alp (N_down B_du - N_up B_ud) phi
  double k_aa, c_src, c_alp, inv_nu0;
  double kia;  // sqrt(kTabN/ln 2) / k_aa: scaled reciprocal Doppler width per unit 1/width (tile_kernel)
};

// one tile_kernel block: items [g0, g1) of ray `ray`'s item list, which belong to the nlc lines
// l0 .. l0+nlc-1 of the batch
struct __align__(16) TileDesc {
  int ray;
  unsigned g0, g1;
  unsigned short l0, nlc;
};

// one ztile_kernel warp: ray `ray` x the nlt lines zlines[loff .. loff+nlt) (one line per lane, the lane
// pattern repeated 32 >> lwshift times) x the channels of list positions j0 .. j0+nchk-1, where the
// ray's channel list is {0, cmin, cmin+1, ...}
struct __align__(16) ZTile {
  int ray;
  unsigned loff;
  unsigned short cmin, j0, nchk;
  unsigned char nlt, lwshift;
};
#ifndef RL_ZCW
#define RL_ZCW 9
#endif
constexpr int kZCw = RL_ZCW;         // channels per ztile_kernel thread (three groups of three)
constexpr int kZTab = 1024;     // profile-table entries per warp
#ifndef RL_ZWARPS
#define RL_ZWARPS 1
#endif
constexpr int kZWarps = RL_ZWARPS;  // warps (= tiles) per ztile_kernel block

// per cell, one bit per line of the batch: N_up + N_down surely above / surely below LEVTHRES
struct __align__(16) CellMask {
  uint4 on, off;
};

struct RenderParams {
  GridDev g;
  int nray, nphi, nrr;
  int nl;    // lines in this batch
  int nfr;
  int subgrid, nonredundant;
  int ring_lo, ring_hi;  // camera rings traced by this call (0 = the centre ray, 1..nrr); others are left out
  double levthres;
  double aksmax_c;  // aksmax/2.99792458d5
  double starfract;  // (rstar/rbeam0)^2 for the centre ray
  int out_itype;
  const long long *node_off;
  NodesDev nodes;
  const double4 *cellL;  // [ncell][nl] (cell-major: the lines of a tile at one cell are contiguous)
  const double2 *cellD;  // [ncell][nl] the dust pair {src_dust, alp_dust} of cellL once more, compact: zcont_kernel's
                         // dust-only gathers touch half the sectors
  long long ncell;
  const LineDev *lines;      // [nl]
  const double *line_dnu;    // [nl][nfr]
  const double *velo;        // [nl][nfr]
  const double *velz;        // [nfr] channel velocities / c of the passband, the same for every line (ztile_kernel)
  const double *star_line;   // [nl][nfr]
  const double *isrf_line;   // [nl][nfr]
  // per task (line_local*nray + ray)
  int4 *rng;                 // {lo, hi, c0, ch0_in_range}
  // tasks are ray-major: task = ray * nl + line_slot
  unsigned int *nitems;      // [ntask+1] channels the reference integrates for the task
  const unsigned int *item_off;  // [ntask+1] exclusive scan of nitems (ray-major item list)
  unsigned int *ncta;        // [nray+1] thread blocks of tile_kernel per ray
  const unsigned int *cta_off;   // [nray+1] exclusive scan of ncta
  int smem_budget;           // dynamic shared memory per tile_kernel block [bytes]
  int tile_threads;          // threads per tile_kernel block (64 or 128) = most items of a tile
  int tile_max_lines;        // most lines a tile may span (shared-memory stage)
  TileDesc *tiles;           // [total tiles] written by plan_kernel<true>
  int use_z;                 // 1: ztile_kernel (lines across lanes), 0: tile_kernel
  int zlw;                   // lines per ztile_kernel tile (power of two <= 32)
  ZTile *ztiles;             // [nztile] written by zplan_kernel<true>
  // cost order (ztile_kernel): zplan_kernel<true> writes the general tiles to ztiles_in with a key that falls with
  // the tile's work; a radix sort then leaves them in ztiles longest first, so that the last blocks of the launch
  // are its shortest (null: the tiles stay in ray order)
  ZTile *ztiles_in;
  unsigned *zkeys;
  unsigned short *zlines;    // [nray][nl] per ray: the lines with a channel window, then the others
  unsigned nztile;
  CellMask *masks;           // [ncell]
  int sparse;                // 1: skipped channels are not materialised in img (spectrum only)
  unsigned char *dense;      // [ntask] sparse mode: row was completed by fill_kernel
  double *img;               // [nl][nrr+1][nphi][nfr]
  unsigned char *integ;      // [nl][nrow][nfr] 1 = channel was integrated with cmask=1
  double *tau_center;        // [nl]
  int *maser;                // [nl]
  unsigned long long *counters;  // {R, E, S, executed element integrations}
  // opaque-wall start (ztile_kernel): nstart[ray] = first segment that is integrated; the segments before it
  // lie behind more than wall_tau of dust optical depth for every line of the batch
  const int *nstart;         // [nray] or null
  double wall_tau;
  int *status;
};

// per-line preparation (prep_cells_kernel)
struct PrepParams {
  long long ncell;
  int nl;
  int nlevels;
  const double *popul;   // [ncell][nlevels]
  const double *abund;   // [ncell]
  const double *rho;     // [ncell]
  double molpg;
  const int *lev_up;     // [nl] (1-based)
  const int *lev_down;
  // dust
  int use_dust;          // 1: compute from dust tables, 0: ld_src/ld_alp given
  int nspec, maxsize, ncf;
  const int *nsize;
  const double *kabs;    // [nspec][maxsize][ncf]
  const double *kscat;
  const double *drho;    // [ncell][nspec]
  const double *dtemp;   // [ncell][nspec][maxsize]
  const double *scat;    // [ncell][ncf] or null
  const int *inudust;    // [nl] hunt result on cont_freq_nu (1-based), 0 or ncf = outside
  const double *wgt;     // [nl]
  const double *freq;    // [nl]
  const double *ld_src;  // [nl][ncell]
  const double *ld_alp;
  double4 *cellL;        // [ncell][nl]
  double2 *cellD;        // [ncell][nl] {src_dust, alp_dust}
};

}  // namespace rl

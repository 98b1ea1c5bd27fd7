/* taxila_oracle.c -- CPU restatement of Taxila-LBM's flow time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under taxila-lbm_b200/ links, loads or calls
 * this file; it is the checker for the CUDA path (tests/, __graft_entry__.smoke(),
 * and the cpu_baseline / --impl reference legs of bench.py).
 *
 * It follows the reference's algorithm in the reference's own structure:
 * array-of-structures fields with PETSc-DMDA style ghost layers, one full sweep
 * per phase, a full-size streaming temporary, push streaming followed by a
 * bounce-back sweep over the ghosted box.  Every function cites the reference
 * file:line it restates (paths relative to the reference tree).  No reference
 * source is copied: the Fortran is restated in C, and the fluid-fluid stencil
 * rows are numbers extracted by oracle/gen_stencil_tables.py.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this oracle against the
 * reference's shipped golden vector tests/bubble_2D/reference_solution/fi001.dat
 * (committed copy: tests/golden/bubble_2D_fi001.dat) -- D2Q9/SRT/iso-4 directly,
 * the MRT tables through MRT(all rates 1) == SRT, and the D3Q19 tables through a
 * z-invariant extrusion projected onto the same golden.  D3Q19 MRT with general
 * rates, iso-8/10, walls, minerals and body force have no golden in the
 * reference ("parity unpinned" for those features).  For them
 * tests/test_oracle_textbook.py holds this file against tests/textbook_lbm.py, an
 * independently written whole-array model of the same equations (which itself
 * reproduces the golden): agreement to round-off on MRT with general rates, the
 * order-8/10 stencils, bounce-back, mineral and body forces, the EOS kinds, planar
 * free-slip walls, density / flux / velocity faces and pressure outlets.
 *
 * Floating point: compile with -ffp-contract=off and without -ffast-math so the
 * evaluation order below is the one executed.  Default-real Fortran literals
 * that are NOT exactly representable are reproduced as float constants
 * (1./6., 1./12., 1./3. in the fluid-solid force, 1.e-12 for eps).
 *
 * Layout (first index fastest, as in lbm_flow.F90:963-970):
 *   fi    [S][Q][gx][gy][gz]      ghost width 1
 *   rho   [S][rgx][rgy][rgz]      ghost width R = stencil_size_rho
 *   flux, forces [S][D][gx][gy][gz]
 *   walls [rgx][rgy][rgz]         doubles holding small integers
 * For ndims == 2 the z extent is 1 with no z ghost.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/taxila_gpu.h"
#include "ff_stencil_tables.h"

#define MAXQ 19
#define MAXS TXG_NMAX_COMPONENTS

typedef struct txo_state {
  txg_config cfg;
  int D, Q, S, R;
  int NX, NY, NZ;
  int gz;  /* ghost width in z for "g" arrays: 1 (3-D) or 0 (2-D) */
  int rgz; /* ghost width in z for "rg" arrays: R or 0 */
  int gnx, gny, gnz;    /* ghosted extents, width 1 */
  int rgnx, rgny, rgnz; /* ghosted extents, width R */
  /* lattice (lbm_discretization_d3q19.F90:64-232, lbm_discretization_d2q9.F90:53-144) */
  int ci[MAXQ][3];
  double weights[MAXQ];
  int opposites[MAXQ], reflect[3][MAXQ];
  double mt_mrt[MAXQ][MAXQ]; /* mt_mrt[n][i] = M(n, i): row n of M */
  double mmt_mrt[MAXQ];
  double ffw[41];
  double c_0;
  /* per component */
  double tau_mrt[MAXS][MAXQ];
  double d_k[MAXS], c_s2[MAXS], s_c[MAXS];
  /* fields */
  double *fi, *fi_eq, *rho, *psi, *flux, *forces, *walls, *stream_tmp;
  double *fi_old; /* DistributionCalcDeltaNorm */
  int have_old;
  int threads; /* 1 = literal serial order everywhere */
  /* external boundary conditions (bc_type, lbm_bc.F90:32-48): flags in cfg.bc_flags */
  double *bc_vals[6]; /* face arrays xm,xp,ym,yp,zm,zp: [t2][t1][nbcs], nbcs = D*S */
  int bc_pressure_outlet[6];   /* flow%bc_flags(b) .eq. BC_PRESSURE_OUTLET (lbm_flow.F90:1170-1189) */
  double bc_outlet_pressure[6]; /* flow%bc_data(1,b) of such a face */
  int eos_bad;        /* EOS_PR: values whose inner square root went negative so far */
  int prestream;      /* 1 (default): BCPreStream runs as in the reference; 0: skipped (test of its effect) */
} txo_state;

/* ---------------------------------------------------------------- indexing */
static inline size_t gnode(const txo_state *s, int i, int j, int k) {
  return ((size_t)(k + s->gz) * s->gny + (size_t)(j + 1)) * s->gnx + (size_t)(i + 1);
}
static inline size_t rgnode(const txo_state *s, int i, int j, int k) {
  return ((size_t)(k + s->rgz) * s->rgny + (size_t)(j + s->R)) * s->rgnx + (size_t)(i + s->R);
}
#define FI(s, m, n, i, j, k) ((s)->fi[gnode(s, i, j, k) * (size_t)((s)->S * (s)->Q) + (size_t)(n) * (s)->S + (m)])
#define FEQ(s, m, n, i, j, k) ((s)->fi_eq[gnode(s, i, j, k) * (size_t)((s)->S * (s)->Q) + (size_t)(n) * (s)->S + (m)])
#define RHO(s, m, i, j, k) ((s)->rho[rgnode(s, i, j, k) * (size_t)(s)->S + (m)])
#define PSI(s, m, i, j, k) ((s)->psi[rgnode(s, i, j, k) * (size_t)(s)->S + (m)])
#define FLUX(s, m, d, i, j, k) ((s)->flux[gnode(s, i, j, k) * (size_t)((s)->S * (s)->D) + (size_t)(d) * (s)->S + (m)])
#define FRC(s, m, d, i, j, k) ((s)->forces[gnode(s, i, j, k) * (size_t)((s)->S * (s)->D) + (size_t)(d) * (s)->S + (m)])
#define WALLS(s, i, j, k) ((s)->walls[rgnode(s, i, j, k)])

/* ---------------------------------------------------------------- lattices */
/* DiscretizationSetup_D3Q19, lbm_discretization_d3q19.F90:64-232 */
static void setup_d3q19(txo_state *s) {
  static const int cx[19] = {0, 1, 0, -1, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1, 0, 0, 0, 0};
  static const int cy[19] = {0, 0, 1, 0, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, -1, 1};
  static const int cz[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
  static const int M[19][19] = {
      {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
      {-30, -11, -11, -11, -11, -11, -11, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8},
      {12, -4, -4, -4, -4, -4, -4, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
      {0, 1, 0, -1, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1, 0, 0, 0, 0},
      {0, -4, 0, 4, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1, 0, 0, 0, 0},
      {0, 0, 1, 0, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, -1, 1},
      {0, 0, -4, 0, 4, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, -1, 1},
      {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1},
      {0, 0, 0, 0, 0, -4, 4, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1},
      {0, 2, -1, 2, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2},
      {0, -4, 2, -4, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2},
      {0, 0, 1, 0, 1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0},
      {0, 0, -2, 0, -2, 2, 2, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0},
      {0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 0, 0, 0, 0},
      {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1},
      {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0},
      {0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, -1, 1, 1, -1, 0, 0, 0, 0},
      {0, 0, 0, 0, 0, 0, 0, -1, -1, 1, 1, 0, 0, 0, 0, 1, -1, -1, 1},
      {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, -1, -1, 1, 1}};
  static const int mmt[19] = {19, 2394, 252, 10, 40, 10, 40, 10, 40, 36, 72, 12, 24, 4, 4, 4, 8, 8, 8};
  s->Q = 19;
  s->D = 3;
  for (int n = 0; n < 19; ++n) {
    s->ci[n][0] = cx[n];
    s->ci[n][1] = cy[n];
    s->ci[n][2] = cz[n];
    s->weights[n] = n == 0 ? 1.0 / 3.0 : (n <= 6 ? 1.0 / 18.0 : 1.0 / 36.0);
    s->mmt_mrt[n] = mmt[n];
    for (int i = 0; i < 19; ++i) s->mt_mrt[n][i] = M[n][i];
  }
  s->c_0 = 6.0;
  memset(s->ffw, 0, sizeof s->ffw);
  if (s->cfg.isotropy_order == 4) {
    s->ffw[1] = 1.0 / 6.;
    s->ffw[2] = 1.0 / 12.;
  } else if (s->cfg.isotropy_order == 8) {
    s->ffw[1] = 4.0 / 45.;
    s->ffw[2] = 1.0 / 21.;
    s->ffw[3] = 2.0 / 105.;
    s->ffw[4] = 5.0 / 504.;
    s->ffw[5] = 1.0 / 315.;
    s->ffw[6] = 1.0 / 630.;
    s->ffw[8] = 1.0 / 5040.;
  }
}

/* DiscretizationSetUp_D2Q9, lbm_discretization_d2q9.F90:53-144 */
static void setup_d2q9(txo_state *s) {
  static const int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
  static const int cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
  static const int M[9][9] = {{1, 1, 1, 1, 1, 1, 1, 1, 1},     {-4, -1, -1, -1, -1, 2, 2, 2, 2},
                              {4, -2, -2, -2, -2, 1, 1, 1, 1}, {0, 1, 0, -1, 0, 1, -1, -1, 1},
                              {0, -2, 0, 2, 0, 1, -1, -1, 1},  {0, 0, 1, 0, -1, 1, 1, -1, -1},
                              {0, 0, -2, 0, 2, 1, 1, -1, -1},  {0, 1, -1, 1, -1, 0, 0, 0, 0},
                              {0, 0, 0, 0, 0, 1, -1, 1, -1}};
  static const int mmt[9] = {9, 36, 36, 6, 12, 6, 12, 4, 4};
  s->Q = 9;
  s->D = 2;
  for (int n = 0; n < 9; ++n) {
    s->ci[n][0] = cx[n];
    s->ci[n][1] = cy[n];
    s->ci[n][2] = 0;
    s->weights[n] = n == 0 ? 4.0 / 9.0 : (n <= 4 ? 1.0 / 9.0 : 1.0 / 36.0);
    s->mmt_mrt[n] = mmt[n];
    for (int i = 0; i < 9; ++i) s->mt_mrt[n][i] = M[n][i];
  }
  s->c_0 = 6.0;
  memset(s->ffw, 0, sizeof s->ffw);
  if (s->cfg.isotropy_order == 4) {
    s->ffw[1] = 1.0 / 3.;
    s->ffw[2] = 1.0 / 12.;
  } else if (s->cfg.isotropy_order == 8) {
    s->ffw[1] = 4.0 / 21.;
    s->ffw[2] = 4.0 / 45.;
    s->ffw[4] = 1.0 / 60.;
    s->ffw[5] = 2.0 / 315.;
    s->ffw[8] = 1.0 / 5040.;
  } else if (s->cfg.isotropy_order == 10) {
    s->ffw[1] = 262.0 / 1785.;
    s->ffw[2] = 93.0 / 1190.;
    s->ffw[4] = 7.0 / 340.;
    s->ffw[5] = 6.0 / 595.;
    s->ffw[8] = 9.0 / 9520.;
    s->ffw[9] = 2.0 / 5355.;
    s->ffw[10] = 1.0 / 7140.;
  }
}

/* opposites / reflect_{x,y,z}: the reference spells them out direction by direction
 * (lbm_discretization_d3q19.F90:83-161, lbm_discretization_d2q9.F90:68-97); they are
 * exactly "negate all / one component of c_n", which is what is computed here. */
static int find_dir(const txo_state *s, int x, int y, int z) {
  for (int n = 0; n < s->Q; ++n)
    if (s->ci[n][0] == x && s->ci[n][1] == y && s->ci[n][2] == z) return n;
  return -1;
}
static void setup_permutations(txo_state *s) {
  for (int n = 0; n < s->Q; ++n) {
    const int *c = s->ci[n];
    s->opposites[n] = find_dir(s, -c[0], -c[1], -c[2]);
    s->reflect[0][n] = find_dir(s, -c[0], c[1], c[2]);
    s->reflect[1][n] = find_dir(s, c[0], -c[1], c[2]);
    s->reflect[2][n] = find_dir(s, c[0], c[1], -c[2]);
  }
}

/* DiscretizationSetupRelax_D3Q19 (lbm_discretization_d3q19.F90:234-263),
 * DiscretizationSetUpRelax_D2Q9 (lbm_discretization_d2q9.F90:147-166),
 * RelaxationSetFromOptions (lbm_relaxation.F90:117-151),
 * ComponentSetFromOptions d_k (lbm_component.F90:158) */
static void setup_relaxation(txo_state *s) {
  const txg_config *c = &s->cfg;
  for (int m = 0; m < s->S; ++m) {
    s->c_s2[m] = 1.0 / 3.;
    s->d_k[m] = 1. - 2. / (3. * c->mm[m]);
    if (c->relaxation_mode == TXG_RELAXATION_MODE_SRT)
      s->s_c[m] = 1.0 / c->tau[m];
    else
      s->s_c[m] = c->s_c[m];
    double *t = s->tau_mrt[m];
    if (s->D == 3) {
      const double r[19] = {c->s_c[m],  c->s_e[m],  c->s_e2[m], c->s_c[m],  c->s_q[m],  c->s_c[m], c->s_q[m],
                            c->s_c[m],  c->s_q[m],  c->s_nu[m], c->s_pi[m], c->s_nu[m], c->s_pi[m], c->s_nu[m],
                            c->s_nu[m], c->s_nu[m], c->s_m[m],  c->s_m[m],  c->s_m[m]};
      memcpy(t, r, sizeof r);
    } else {
      const double r[9] = {c->s_c[m], c->s_e[m], c->s_e2[m], c->s_c[m], c->s_q[m],
                           c->s_c[m], c->s_q[m], c->s_nu[m], c->s_nu[m]};
      memcpy(t, r, sizeof r);
    }
  }
}

/* ---------------------------------------------------------------- lifecycle */
txo_state *txo_create(const txg_config *cfg) {
  if (!cfg || cfg->struct_bytes != (int32_t)sizeof(txg_config)) return NULL;
  txo_state *s = (txo_state *)calloc(1, sizeof *s);
  s->cfg = *cfg;
  s->S = cfg->ncomponents;
  s->R = cfg->stencil_size_rho;
  s->NX = cfg->NX;
  s->NY = cfg->NY;
  s->NZ = cfg->ndims == 3 ? cfg->NZ : 1;
  if (cfg->discretization == TXG_D3Q19_DISCRETIZATION)
    setup_d3q19(s);
  else
    setup_d2q9(s);
  setup_permutations(s);
  setup_relaxation(s);
  s->gz = s->D == 3 ? 1 : 0;
  s->rgz = s->D == 3 ? s->R : 0;
  s->gnx = s->NX + 2;
  s->gny = s->NY + 2;
  s->gnz = s->NZ + 2 * s->gz;
  s->rgnx = s->NX + 2 * s->R;
  s->rgny = s->NY + 2 * s->R;
  s->rgnz = s->NZ + 2 * s->rgz;
  size_t ng = (size_t)s->gnx * s->gny * s->gnz, nrg = (size_t)s->rgnx * s->rgny * s->rgnz;
  /* PETSc Vecs start zeroed (VecSet in WallsSetUp lbm_walls.F90:146-147; local
   * vectors from DMCreateLocalVector are zero-filled) */
  s->fi = (double *)calloc(ng * s->S * s->Q, sizeof(double));
  s->fi_eq = (double *)calloc(ng * s->S * s->Q, sizeof(double));
  s->stream_tmp = (double *)calloc(ng * s->S * s->Q, sizeof(double));
  s->rho = (double *)calloc(nrg * s->S, sizeof(double));
  s->psi = (double *)calloc(nrg * s->S, sizeof(double));
  s->flux = (double *)calloc(ng * s->S * s->D, sizeof(double));
  s->forces = (double *)calloc(ng * s->S * s->D, sizeof(double));
  s->walls = (double *)calloc(nrg, sizeof(double));
  s->fi_old = NULL;
  s->threads = 1;
  s->prestream = 1;
  return s;
}

void txo_destroy(txo_state *s) {
  if (!s) return;
  free(s->fi);
  free(s->fi_eq);
  free(s->stream_tmp);
  free(s->rho);
  free(s->psi);
  free(s->flux);
  free(s->forces);
  free(s->walls);
  free(s->fi_old);
  for (int b = 0; b < 6; ++b) free(s->bc_vals[b]);
  free(s);
}

void txo_set_threads(txo_state *s, int n) {
  s->threads = n < 1 ? 1 : n;
#ifdef _OPENMP
  omp_set_num_threads(s->threads);
#endif
}

/* ---------------------------------------------------------------- ghost exchange
 * DMLocalToLocalBegin/End on a DMDA with a box stencil of width w, single rank:
 * a ghost entry is overwritten by its periodic image iff every out-of-range
 * coordinate lies in a DM_BOUNDARY_PERIODIC direction; ghosts beyond a
 * DM_BOUNDARY_GHOSTED face are left alone (lbm_grid.F90:150-157). */
static void local_to_local(const txo_state *s, double *a, int dof, int w, int wz) {
  const int NX = s->NX, NY = s->NY, NZ = s->NZ;
  const int nx = NX + 2 * w, ny = NY + 2 * w;
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = -wz; k < NZ + wz; ++k)
    for (int j = -w; j < NY + w; ++j)
      for (int i = -w; i < NX + w; ++i) {
        int oi = i < 0 || i >= NX, oj = j < 0 || j >= NY, ok = k < 0 || k >= NZ;
        if (!(oi || oj || ok)) continue;
        if ((oi && !s->cfg.periodic[0]) || (oj && !s->cfg.periodic[1]) || (ok && !s->cfg.periodic[2])) continue;
        int si = ((i % NX) + NX) % NX, sj = ((j % NY) + NY) % NY, sk = ((k % NZ) + NZ) % NZ;
        size_t dst = ((size_t)(k + wz) * ny + (size_t)(j + w)) * nx + (size_t)(i + w);
        size_t src = ((size_t)(sk + wz) * ny + (size_t)(sj + w)) * nx + (size_t)(si + w);
        memcpy(a + dst * dof, a + src * dof, sizeof(double) * (size_t)dof);
      }
}
/* DistributionCommunicateFi, lbm_distribution_function.F90:309-334 */
static void communicate_fi(txo_state *s) { local_to_local(s, s->fi, s->S * s->Q, 1, s->gz); }
/* DistributionCommunicateDensityBegin/End, lbm_distribution_function.F90:342-357 */
static void communicate_density(txo_state *s) { local_to_local(s, s->rho, s->S, s->R, s->rgz); }

/* WallsSetGhostNodesD2/D3 (lbm_walls.F90:190-231) then WallsCommunicate (:233-244).
 * `walls_natural` is the global Vec in natural ordering (x fastest). */
void txo_set_walls(txo_state *s, const double *walls_natural) {
  const int NX = s->NX, NY = s->NY, NZ = s->NZ, R = s->R, RZ = s->rgz;
  for (int k = 0; k < NZ; ++k)
    for (int j = 0; j < NY; ++j)
      for (int i = 0; i < NX; ++i) WALLS(s, i, j, k) = walls_natural[((size_t)k * NY + j) * NX + i];
  for (int k = -RZ; k < NZ + RZ; ++k)
    for (int j = -R; j < NY + R; ++j)
      for (int i = -R; i < NX + R; ++i) {
        int ghost = 0;
        if ((i < 0 || i >= NX) && !s->cfg.periodic[0]) ghost = 1;
        if ((j < 0 || j >= NY) && !s->cfg.periodic[1]) ghost = 1;
        if (s->D == 3 && (k < 0 || k >= NZ) && !s->cfg.periodic[2]) ghost = 1;
        if (ghost) WALLS(s, i, j, k) = TXG_WALL_GHOST;
      }
  local_to_local(s, s->walls, 1, R, RZ);
}

/* the ghosted walls array as the Fortran side would hold it */
void txo_get_walls_rg(const txo_state *s, double *out) {
  memcpy(out, s->walls, sizeof(double) * (size_t)s->rgnx * s->rgny * s->rgnz);
}

/* LBMInitializeState (lbm.F90:444-453): rho from natural order [node][m], u = 0 */
void txo_set_rho(txo_state *s, const double *rho_natural) {
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i)
        for (int m = 0; m < s->S; ++m)
          RHO(s, m, i, j, k) = rho_natural[(((size_t)k * s->NY + j) * s->NX + i) * s->S + m];
  memset(s->flux, 0, sizeof(double) * (size_t)s->gnx * s->gny * s->gnz * s->S * s->D);
}

void txo_set_fi(txo_state *s, const double *fi_natural) {
  const int SQ = s->S * s->Q;
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i)
        memcpy(&FI(s, 0, 0, i, j, k), fi_natural + (((size_t)k * s->NY + j) * s->NX + i) * SQ, sizeof(double) * SQ);
}

/* ---------------------------------------------------------------- equilibrium
 * DiscretizationEquilf_D3Q19 (lbm_discretization_d3q19.F90:265-305),
 * DiscretizationEquilf_D2Q9 (lbm_discretization_d2q9.F90:168-205),
 * driver FlowUpdateFeq (lbm_flow.F90:822-834). */
static void update_feq(txo_state *s) {
  const int D = s->D, Q = s->Q;
  for (int m = 0; m < s->S; ++m) {
    const double d_k = s->d_k[m], c_s2 = s->c_s2[m];
#pragma omp parallel for schedule(static) if (s->threads > 1)
    for (int k = 0; k < s->NZ; ++k)
      for (int j = 0; j < s->NY; ++j)
        for (int i = 0; i < s->NX; ++i) {
          if (WALLS(s, i, j, k) != 0.) continue;
          double usqr = 0.;
          for (int d = 0; d < D; ++d) usqr += FLUX(s, m, d, i, j, k) * FLUX(s, m, d, i, j, k);
          const double rho = RHO(s, m, i, j, k);
          if (D == 3)
            FEQ(s, m, 0, i, j, k) = rho * (d_k - usqr / 2.);
          else
            FEQ(s, m, 0, i, j, k) = rho * ((1. + d_k * 5.) / 6. - 2. * usqr / 3.);
          for (int n = 1; n < Q; ++n) {
            double udote = 0.;
            for (int d = 0; d < D; ++d) udote += s->ci[n][d] * FLUX(s, m, d, i, j, k);
            FEQ(s, m, n, i, j, k) =
                s->weights[n] * rho *
                (1.5 * (1. - d_k) + udote / c_s2 + udote * udote / (2. * c_s2 * c_s2) - usqr / (2. * c_s2));
          }
        }
  }
}

/* FlowFiBarEqPrefactor, lbm_flow.F90:836-851 */
static void prefactor_node(const txo_state *s, int i, int j, int k, double pref[MAXS][MAXQ]) {
  for (int n = 0; n < s->Q; ++n)
    for (int m = 0; m < s->S; ++m) {
      double acc = 0.;
      for (int d = 0; d < s->D; ++d) acc += FRC(s, m, d, i, j, k) * (s->ci[n][d] - FLUX(s, m, d, i, j, k));
      pref[m][n] = acc / (RHO(s, m, i, j, k) * s->c_s2[m]);
    }
}

/* RelaxationCollideSRT (lbm_relaxation.F90:171-180), RelaxationCollideMRT (:182-200) */
static void relaxation_collide(const txo_state *s, int m, double f[MAXS][MAXQ], double feqbar[MAXS][MAXQ]) {
  const int Q = s->Q;
  if (s->cfg.relaxation_mode == TXG_RELAXATION_MODE_SRT) {
    const double tau = s->cfg.tau[m];
    for (int n = 0; n < Q; ++n) f[m][n] = f[m][n] - (f[m][n] - feqbar[m][n]) / tau;
  } else {
    double d_fi[MAXQ];
    for (int n = 0; n < Q; ++n) d_fi[n] = f[m][n] - feqbar[m][n];
    for (int n = 0; n < Q; ++n) {
      double mdiff = 0.;
      for (int i = 0; i < Q; ++i) mdiff += s->mt_mrt[n][i] * d_fi[i];
      for (int i = 0; i < Q; ++i) f[m][i] = f[m][i] - s->tau_mrt[m][n] * mdiff / s->mmt_mrt[n] * s->mt_mrt[n][i];
    }
  }
}

/* FlowCollision -> FlowCollisionD3/D2, lbm_flow.F90:936-1029 */
static void flow_collision(txo_state *s) {
  update_feq(s);
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i) {
        if (WALLS(s, i, j, k) != 0.) continue;
        double pref[MAXS][MAXQ], f[MAXS][MAXQ], feqbar[MAXS][MAXQ];
        prefactor_node(s, i, j, k, pref);
        for (int m = 0; m < s->S; ++m)
          for (int n = 0; n < s->Q; ++n) {
            f[m][n] = FI(s, m, n, i, j, k);
            feqbar[m][n] = (1. - .5 * pref[m][n]) * FEQ(s, m, n, i, j, k);
          }
        for (int m = 0; m < s->S; ++m) relaxation_collide(s, m, f, feqbar);
        for (int m = 0; m < s->S; ++m)
          for (int n = 0; n < s->Q; ++n) FI(s, m, n, i, j, k) = f[m][n] + pref[m][n] * FEQ(s, m, n, i, j, k);
      }
}

/* ---------------------------------------------------------------- streaming
 * DistributionStreamD3/D2, lbm_distribution_function.F90:560-652.  The reference
 * streams into an uninitialised automatic array and copies it back; entries it
 * never writes are garbage.  They are NaN here so that any dependence on them
 * would poison the result and fail the golden comparison. */
static void stream(txo_state *s) {
  const int S = s->S, Q = s->Q, SQ = S * Q;
  const size_t ng = (size_t)s->gnx * s->gny * s->gnz;
  double *tmp = s->stream_tmp;
  const double qnan = NAN;
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (size_t a = 0; a < ng * SQ; ++a) tmp[a] = qnan;
  for (int n = 0; n < Q; ++n) {
    const int cx = s->ci[n][0], cy = s->ci[n][1], cz = s->ci[n][2];
    /* destination box = owned box extended by one cell along +c_n */
    int xs = 0, xe = s->NX - 1, ys = 0, ye = s->NY - 1, zs = 0, ze = s->NZ - 1;
    if (cx < 0) xs += cx; else if (cx > 0) xe += cx;
    if (cy < 0) ys += cy; else if (cy > 0) ye += cy;
    if (cz < 0) zs += cz; else if (cz > 0) ze += cz;
#pragma omp parallel for schedule(static) if (s->threads > 1)
    for (int k = zs; k <= ze; ++k)
      for (int j = ys; j <= ye; ++j)
        for (int i = xs; i <= xe; ++i) {
          const double *src = &FI(s, 0, n, i - cx, j - cy, k - cz);
          double *dst = tmp + gnode(s, i, j, k) * SQ + (size_t)n * S;
          for (int m = 0; m < S; ++m) dst[m] = src[m];
        }
  }
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (size_t a = 0; a < ng * SQ; ++a) s->fi[a] = tmp[a];
}

/* DistributionBouncebackD3/D2, lbm_distribution_function.F90:669-784: literal
 * serial sweep over the ghosted (width 1) box in k, j, i order. */
static inline int owned(const txo_state *s, int i, int j, int k) {
  return i >= 0 && i < s->NX && j >= 0 && j < s->NY && k >= 0 && k < s->NZ;
}
static void bounceback_node(txo_state *s, int i, int j, int k, int zero) {
  const double w = WALLS(s, i, j, k);
  const int S = s->S, Q = s->Q;
  if (!(w > 0)) return;
  int axis = -1;
  if (w == TXG_WALL_NORMAL_X) axis = 0;
  else if (w == TXG_WALL_NORMAL_Y) axis = 1;
  else if (w == TXG_WALL_NORMAL_Z && s->D == 3) axis = 2;
  if (zero != 1) {
    for (int n = 0; n < Q; ++n) {
      int ni = i, nj = j, nk = k, nn;
      if (axis < 0) {
        ni = i - s->ci[n][0];
        nj = j - s->ci[n][1];
        nk = k - s->ci[n][2];
        nn = s->opposites[n];
      } else {
        if (axis == 0) ni = i - s->ci[n][0];
        if (axis == 1) nj = j - s->ci[n][1];
        if (axis == 2) nk = k - s->ci[n][2];
        nn = s->reflect[axis][n];
      }
      if (owned(s, ni, nj, nk))
        for (int m = 0; m < S; ++m) FI(s, m, nn, ni, nj, nk) = FI(s, m, n, i, j, k);
    }
  }
  if (zero != 0)
    for (int a = 0; a < S * Q; ++a) (&FI(s, 0, 0, i, j, k))[a] = 0.;
}
static void bounceback(txo_state *s) {
  const int gz = s->gz;
  if (s->threads <= 1) {
    for (int k = -gz; k < s->NZ + gz; ++k)
      for (int j = -1; j <= s->NY; ++j)
        for (int i = -1; i <= s->NX; ++i) bounceback_node(s, i, j, k, 2);
  } else {
    /* timing variant: all pushes, then all zeroing.  Identical to the serial sweep
     * whenever solid nodes hold f = 0 on entry (true for every state the flow
     * update itself produces) and no 900-902 codes are present; checked in
     * tests/test_oracle_selfchecks.py. */
#pragma omp parallel for schedule(static)
    for (int k = -gz; k < s->NZ + gz; ++k)
      for (int j = -1; j <= s->NY; ++j)
        for (int i = -1; i <= s->NX; ++i) bounceback_node(s, i, j, k, 0);
#pragma omp parallel for schedule(static)
    for (int k = -gz; k < s->NZ + gz; ++k)
      for (int j = -1; j <= s->NY; ++j)
        for (int i = -1; i <= s->NX; ++i) bounceback_node(s, i, j, k, 1);
  }
}

/* ---------------------------------------------------------------- moments */
/* DistributionCalcDensityD3/D2, lbm_distribution_function.F90:379-428 */
static void calc_density(txo_state *s) {
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i)
        for (int m = 0; m < s->S; ++m) {
          double acc = 0.;
          if (WALLS(s, i, j, k) == 0.)
            for (int n = 0; n < s->Q; ++n) acc += FI(s, m, n, i, j, k);
          RHO(s, m, i, j, k) = acc;
        }
}

/* DistributionCalcFluxD3/D2, lbm_distribution_function.F90:451-508 */
static void calc_flux_into(const txo_state *s, double *flux) {
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i)
        for (int m = 0; m < s->S; ++m)
          for (int d = 0; d < s->D; ++d) {
            double acc = 0.;
            if (WALLS(s, i, j, k) == 0.)
              for (int n = 0; n < s->Q; ++n) acc += FI(s, m, n, i, j, k) * (double)s->ci[n][d];
            flux[gnode(s, i, j, k) * (size_t)(s->S * s->D) + (size_t)d * s->S + m] = acc;
          }
}

/* FlowUpdateUED3/D2, lbm_flow.F90:494-574 */
static void update_ue(txo_state *s) {
  double mmot[MAXS];
  for (int m = 0; m < s->S; ++m) mmot[m] = s->cfg.mm[m] * s->s_c[m];
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i) {
        if (WALLS(s, i, j, k) != 0.) continue;
        double up[3];
        for (int d = 0; d < s->D; ++d)
          for (int m = 0; m < s->S; ++m) FLUX(s, m, d, i, j, k) = FLUX(s, m, d, i, j, k) + .5 * FRC(s, m, d, i, j, k);
        for (int d = 0; d < s->D; ++d) {
          double num = 0., den = 0.;
          for (int m = 0; m < s->S; ++m) num += FLUX(s, m, d, i, j, k) * mmot[m];
          for (int m = 0; m < s->S; ++m) den += RHO(s, m, i, j, k) * mmot[m];
          up[d] = num / den;
        }
        for (int m = 0; m < s->S; ++m)
          for (int d = 0; d < s->D; ++d) FLUX(s, m, d, i, j, k) = up[d];
      }
}

/* ---------------------------------------------------------------- forces */
/* EOSApply -> EOSApply_Rho / _SC / _Thermo / _PR, lbm_eos.F90:149-349 (whole ghosted array, component by
 * component; g_mm = gf(m,m), lbm_flow.F90:799).  0.37464, 1.54226, 0.26992 are default-real literals in
 * EOSApply_PR (:331): their single-precision values enter the double arithmetic.  Returns the number of
 * values whose inner square root went negative (the reference stops with LBMError at the first, :337-341). */
static int eos_apply(txo_state *s) {
  const size_t nrg = (size_t)s->rgnx * s->rgny * s->rgnz;
  int bad = 0;
  for (int m = 0; m < s->S; ++m) {
    const int type = s->cfg.eos_type[m];
    const double rho0 = s->cfg.eos_rho0[m], psi0 = s->cfg.eos_psi0[m];
    const double a = s->cfg.eos_pr_a[m], b = s->cfg.eos_pr_b[m], R = s->cfg.eos_pr_R[m], T = s->cfg.eos_pr_T[m];
    const double omega = s->cfg.eos_pr_omega[m], g_mm = s->cfg.gf[m][m];
    double alpha = 1. + ((double)0.37464f + (double)1.54226f * omega - (double)0.26992f * (omega * omega)) *
                            (1. - sqrt(T / s->cfg.eos_pr_Tc[m]));
    alpha = alpha * alpha;
    for (size_t i = 0; i < nrg; ++i) {
      const double r = s->rho[i * s->S + m];
      double psi = r;
      if (type == TXG_EOS_SC) {
        psi = rho0 * (1. - exp(-r / rho0));
      } else if (type == TXG_EOS_THERMO) {
        psi = psi0 * exp(-rho0 / r);
      } else if (type == TXG_EOS_PR) {
        const double tmp =
            2. * (r * R * T / (1. - b * r) - (a * alpha * (r * r)) / (1. + 2. * b * r - (b * r) * (b * r)) - r / 3.) /
            (s->c_0 * g_mm);
        if (tmp < 0) ++bad;
        psi = sqrt(tmp);
      }
      s->psi[i * s->S + m] = psi;
    }
  }
  s->eos_bad += bad;
  return bad;
}

/* LBMAddFluidSolidForcesD3/D2, lbm_forcing.F90:1326-1421 (+ the neighbour gather
 * DistributionGatherValueToDirectionD*, lbm_distribution_function.F90:525-542,786-807).
 * 1./6., 1./12., 1./3. are default-real (single precision) literals there.
 * Codes >= nminerals+1 below 998 (800, 900-902) index minerals(:) out of bounds in
 * the reference; they contribute nothing here. */
static void add_fluid_solid(txo_state *s) {
  const float w_axis_f = s->D == 3 ? 1.f / 6.f : 1.f / 3.f;
  const float w_diag_f = 1.f / 12.f;
  const int naxis = 2 * s->D;
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i) {
        if (WALLS(s, i, j, k) != 0.) continue;
        for (int n = 1; n < s->Q; ++n) {
          const double t = WALLS(s, i + s->ci[n][0], j + s->ci[n][1], k + s->ci[n][2]);
          if (!(t > 0. && t < 998.)) continue;
          const int mineral = (int)t;
          if (mineral > s->cfg.nminerals) continue;
          const double w = n <= naxis ? (double)w_axis_f : (double)w_diag_f;
          for (int d = 0; d < s->D; ++d)
            for (int m = 0; m < s->S; ++m)
              FRC(s, m, d, i, j, k) =
                  FRC(s, m, d, i, j, k) - w * RHO(s, m, i, j, k) * s->cfg.gw[mineral - 1][m] * s->ci[n][d];
        }
      }
}

/* LBMAddBodyForcesD3/D2, lbm_forcing.F90:1440-1496 */
static void add_body(txo_state *s) {
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i) {
        if (WALLS(s, i, j, k) != 0.) continue;
        for (int m = 0; m < s->S; ++m)
          for (int d = 0; d < s->D; ++d)
            FRC(s, m, d, i, j, k) = FRC(s, m, d, i, j, k) + s->cfg.gvt[d] * s->cfg.mm[m] * RHO(s, m, i, j, k);
      }
}

/* LBMAddFluidFluidForcesD3 (lbm_forcing.F90:51-958) / D2 (:960-1299).  One BLOCK
 * row per `if (walls...) then` block of the reference, in source order. */
static void add_fluid_fluid(txo_state *s, const double *field /* rho or psi, rg layout */) {
  const int S = s->S, D = s->D, order = s->cfg.isotropy_order;
  const double eps = (double)1.e-12f;
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i) {
        if (WALLS(s, i, j, k) != 0.) continue;
        double gradrho[3][MAXS];
        double weightsum[3] = {0., 0., 0.};
        for (int d = 0; d < 3; ++d)
          for (int m = 0; m < S; ++m) gradrho[d][m] = 0.;
        const double *here = field + rgnode(s, i, j, k) * (size_t)S;
#define F(a, b, c) (WALLS(s, i + (a), j + (b), k + (c)) == 0.)
#define BLOCK(minorder, L, dx, dy, dz, cx, cy, cz, wx, wy, wz, los)                         \
  if (order >= (minorder) && (los)) {                                                       \
    const double *there = field + rgnode(s, i + (dx), j + (dy), k + (dz)) * (size_t)S;      \
    const int cc[3] = {cx, cy, cz};                                                          \
    const int ww[3] = {wx, wy, wz};                                                          \
    for (int d = 0; d < D; ++d)                                                              \
      if (cc[d] != 0) {                                                                      \
        const double coef = (cc[d] > 0 ? cc[d] : -cc[d]) * s->ffw[L];                        \
        for (int m = 0; m < S; ++m) {                                                        \
          if (cc[d] > 0)                                                                     \
            gradrho[d][m] = gradrho[d][m] + coef * (there[m] - here[m]);                     \
          else                                                                               \
            gradrho[d][m] = gradrho[d][m] - coef * (there[m] - here[m]);                     \
        }                                                                                    \
        weightsum[d] = weightsum[d] + s->ffw[L] * ww[d];                                     \
      }                                                                                      \
  }
        if (D == 3) {
          TXO_FF_BLOCKS_D3(BLOCK)
        } else {
          TXO_FF_BLOCKS_D2(BLOCK)
        }
#undef BLOCK
#undef F
        for (int m = 0; m < S; ++m)
          for (int d = 0; d < D; ++d)
            if (weightsum[d] > eps) {
              double acc = 0.;
              for (int mp = 0; mp < S; ++mp) acc += s->cfg.gf[m][mp] * (gradrho[d][mp] / weightsum[d]);
              FRC(s, m, d, i, j, k) = FRC(s, m, d, i, j, k) - s->c_0 * here[m] * acc;
            }
      }
}

/* FlowCalcForces, lbm_flow.F90:760-808 */
static void calc_forces(txo_state *s) {
  memset(s->forces, 0, sizeof(double) * (size_t)s->gnx * s->gny * s->gnz * s->S * s->D);
  if (s->cfg.fluidsolid_forces) add_fluid_solid(s);
  if (s->cfg.body_forces) add_body(s);
  if (s->cfg.fluidfluid_forces) {
    communicate_density(s);
    if (s->cfg.use_nonideal_eos) {
      eos_apply(s);
      add_fluid_fluid(s, s->psi);
    } else {
      add_fluid_fluid(s, s->rho);
    }
  }
}

/* ---------------------------------------------------------------- init */
/* FlowFiInit (lbm_flow.F90:923-934) with FlowFiBarInit -> FlowFeqBarD3/D2 (:853-921) */
void txo_fi_init(txo_state *s) {
  calc_forces(s);
  update_feq(s);
#pragma omp parallel for schedule(static) if (s->threads > 1)
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i) {
        if (WALLS(s, i, j, k) != 0.) continue;
        double pref[MAXS][MAXQ];
        prefactor_node(s, i, j, k, pref);
        for (int m = 0; m < s->S; ++m)
          for (int n = 0; n < s->Q; ++n) FI(s, m, n, i, j, k) = (1. - 0.5 * pref[m][n]) * FEQ(s, m, n, i, j, k);
      }
}

/* FlowUpdateMoments, lbm_flow.F90:466-478 */
void txo_update_moments(txo_state *s) {
  calc_density(s);
  calc_flux_into(s, s->flux);
  calc_forces(s);
  update_ue(s);
}

/* ---------------------------------------------------------------- external boundary conditions
 * lbm_bc.F90.  A face is addressed by its boundary number b = 0..5 (XM, XP, YM, YP, ZM, ZP;
 * lbm_definitions.h:45-50 minus one).  The node routines are written for one boundary and rotated
 * onto the others through DiscSetLocalDirections (lbm_discretization_d3q19.F90:566-714,
 * lbm_discretization_d2q9.F90:372-434); the only thing they take from the rotation is
 * ci(directions(local_normal), cardinals(CARDINAL_NORMAL)), which is the INWARD normal sign on
 * every boundary of both lattices (local_normal = UP / SOUTH, :81 / :67), and the set of
 * tangential axes, each of which is treated independently. */
typedef struct {
  int axis, sign, coord; /* normal axis, inward sign, index of the face plane */
  int t1, t2, n1, n2;    /* tangential axes (t1 fastest in the face array) and their extents */
} txo_face;

static txo_face face_of(const txo_state *s, int b) {
  const int N[3] = {s->NX, s->NY, s->NZ};
  txo_face f;
  f.axis = b / 2;
  f.sign = (b % 2 == 0) ? 1 : -1;
  f.coord = (b % 2 == 0) ? 0 : N[f.axis] - 1;
  f.t1 = f.axis == 0 ? 1 : 0;
  f.t2 = f.axis == 2 ? 1 : 2;
  f.n1 = N[f.t1];
  f.n2 = s->D == 3 ? N[f.t2] : 1;
  return f;
}
static inline void face_node(const txo_face *f, int a, int b, int ijk[3]) {
  ijk[f->axis] = f->coord;
  ijk[f->t1] = a;
  ijk[f->t2] = b;
}
static inline int bc_active(const txo_state *s, int b) {
  const int fl = s->cfg.bc_flags[b];
  return b < 2 * s->D && (fl == TXG_BC_DIRICHLET || fl == TXG_BC_NEUMANN || fl == TXG_BC_VELOCITY);
}

/* BCSetValues (lbm_bc.F90:215-228): the face array of one boundary in the reference's layout */
void txo_set_bc_values(txo_state *s, int b, const double *vals) {
  const txo_face f = face_of(s, b);
  const size_t n = (size_t)f.n1 * f.n2 * s->D * s->S;
  free(s->bc_vals[b]);
  s->bc_vals[b] = (double *)malloc(n * sizeof(double));
  memcpy(s->bc_vals[b], vals, n * sizeof(double));
}
void txo_set_prestream(txo_state *s, int on) { s->prestream = on; }
void txo_set_bc_pressure_outlet(txo_state *s, int b, double pressure) {
  s->bc_pressure_outlet[b] = 1;
  s->bc_outlet_pressure[b] = pressure;
}
int txo_eos_bad(const txo_state *s) { return s->eos_bad; }

/* BCPreStream -> BCPreStream_D2/D3, lbm_bc.F90:613-779: on a Dirichlet / Neumann / velocity face every
 * population with a component along the inward normal is copied from the face node X into X - c_n (a
 * ghost node), so that the stream brings it back.  No wall test, whole owned face. */
static void bc_prestream(txo_state *s) {
  for (int b = 0; b < 2 * s->D; ++b) {
    if (!bc_active(s, b)) continue;
    const txo_face f = face_of(s, b);
    for (int n = 1; n < s->Q; ++n) {
      if (s->ci[n][f.axis] * f.sign <= 0) continue;
      for (int bb = 0; bb < f.n2; ++bb)
        for (int a = 0; a < f.n1; ++a) {
          int x[3] = {0, 0, 0};
          face_node(&f, a, bb, x);
          for (int m = 0; m < s->S; ++m)
            FI(s, m, n, x[0] - s->ci[n][0], x[1] - s->ci[n][1], x[2] - s->ci[n][2]) = FI(s, m, n, x[0], x[1], x[2]);
        }
    }
  }
}

/* BCApplyDirichletToRho -> _D2/_D3, lbm_bc.F90:252-434 */
static void bc_dirichlet_to_rho(txo_state *s) {
  const int nbcs = s->D * s->S;
  for (int b = 0; b < 2 * s->D; ++b) {
    if (s->cfg.bc_flags[b] != TXG_BC_DIRICHLET) continue;
    const txo_face f = face_of(s, b);
    for (int bb = 0; bb < f.n2; ++bb)
      for (int a = 0; a < f.n1; ++a) {
        int x[3] = {0, 0, 0};
        face_node(&f, a, bb, x);
        if (WALLS(s, x[0], x[1], x[2]) != 0.) continue;
        const double *v = s->bc_vals[b] + ((size_t)bb * f.n1 + a) * nbcs;
        for (int m = 0; m < s->S; ++m) RHO(s, m, x[0], x[1], x[2]) = v[m];
      }
  }
}

/* BCUpdateRho -> _D2/_D3, lbm_bc.F90:436-611: rho = sum(fi(m,:)) on the fluid nodes of every
 * Dirichlet / Neumann / velocity face */
static void bc_update_rho(txo_state *s) {
  for (int b = 0; b < 2 * s->D; ++b) {
    if (!bc_active(s, b)) continue;
    const txo_face f = face_of(s, b);
    for (int bb = 0; bb < f.n2; ++bb)
      for (int a = 0; a < f.n1; ++a) {
        int x[3] = {0, 0, 0};
        face_node(&f, a, bb, x);
        if (WALLS(s, x[0], x[1], x[2]) != 0.) continue;
        for (int m = 0; m < s->S; ++m) {
          double acc = 0.;
          for (int n = 0; n < s->Q; ++n) acc += FI(s, m, n, x[0], x[1], x[2]);
          RHO(s, m, x[0], x[1], x[2]) = acc;
        }
      }
  }
}

/* FlowUpdateDensityFromPressure, lbm_flow.F90:2235-2263 (components(2)%gf(1) = gf[1][0]) */
static void density_from_pressure(const txo_state *s, double pressure, double rho1frac, double *rho /* stride 1, S entries */) {
  const double eps = 1.e-10;
  if (s->S == 1) {
    rho[0] = pressure * 3.;
  } else {
    if (rho1frac < eps) {
      rho[0] = 0.;
      rho[1] = pressure * 3.;
    } else if (rho1frac > 1 - eps) {
      rho[0] = pressure * 3.;
      rho[1] = 0.;
    } else {
      const double alpha = 1. / (1. / rho1frac - 1.);
      const double g21 = s->cfg.gf[1][0];
      rho[0] = (-(1. + alpha) / 3. + sqrt((1. + alpha) / 3. * (1. + alpha) / 3. + 4. * s->c_0 * g21 * alpha * pressure)) /
               (2 * s->c_0 * g21);
      rho[1] = rho[0] / alpha;
    }
  }
}

/* FlowUpdateBCPressureOutlet -> D2/D3, lbm_flow.F90:1993-2233, called at the top of FlowApplyBCs (:1965-1972,
 * only when ncomponents /= 1): on every fluid node of a pressure-outlet face the Dirichlet densities are
 * re-derived from the face's pressure and the phase fraction rho_1 / sum(rho) of the node one step inside
 * (of the node itself when that one is solid), read from dist%rho as the previous step left it. */
static void bc_pressure_outlet_update(txo_state *s) {
  if (s->S == 1) return;
  const int nbcs = s->D * s->S;
  for (int b = 0; b < 2 * s->D; ++b) {
    if (!s->bc_pressure_outlet[b]) continue;
    const txo_face f = face_of(s, b);
    for (int bb = 0; bb < f.n2; ++bb)
      for (int a = 0; a < f.n1; ++a) {
        int x[3] = {0, 0, 0};
        face_node(&f, a, bb, x);
        if (WALLS(s, x[0], x[1], x[2]) != 0.) continue;
        int y[3] = {x[0], x[1], x[2]};
        y[f.axis] += f.sign;
        if (WALLS(s, y[0], y[1], y[2]) != 0.) y[f.axis] = x[f.axis];
        double sum = 0.;
        for (int m = 0; m < s->S; ++m) sum += RHO(s, m, y[0], y[1], y[2]);
        const double rho1frac = RHO(s, 0, y[0], y[1], y[2]) / sum;
        density_from_pressure(s, s->bc_outlet_pressure[b], rho1frac, s->bc_vals[b] + ((size_t)bb * f.n1 + a) * nbcs);
      }
  }
}

/* "incoming": c_n has a component along the inward normal
 * (ci(local_normal, normal) * ci(n, normal) .eq. 1, lbm_bc.F90:1300) */
static inline int incoming(const txo_state *s, const txo_face *f, int n) { return f->sign * s->ci[n][f->axis] == 1; }

/* the shared tail of the three node routines: fi(m,n) += w_n sum_d c_n,d Q_d on the incoming directions */
static void bc_distribute(const txo_state *s, const txo_face *f, double *fi_m /* stride S */, const double Q[3]) {
  for (int n = 1; n < s->Q; ++n)
    if (incoming(s, f, n)) {
      double acc = 0.;
      for (int d = 0; d < s->D; ++d) acc += s->ci[n][d] * Q[d];
      fi_m[(size_t)n * s->S] = fi_m[(size_t)n * s->S] + s->weights[n] * acc;
    }
}
static double bc_sum_fi(const txo_state *s, const double *fi_m) {
  double acc = 0.;
  for (int n = 0; n < s->Q; ++n) acc += fi_m[(size_t)n * s->S];
  return acc;
}

/* BCApplyDirichletNode, lbm_bc.F90:1273-1333.  pvals(m,1) = v[m]. */
static void bc_dirichlet_node(const txo_state *s, const txo_face *f, double *fi, const double *v) {
  const int D = s->D, S = s->S, N = f->axis;
  for (int m = 0; m < S; ++m) {
    double *fm = fi + m;
    double Q[3] = {0., 0., 0.}, weightsum[3] = {0., 0., 0.}, momentum[3] = {0., 0., 0.};
    for (int n = 1; n < s->Q; ++n)
      if (incoming(s, f, n)) weightsum[N] = weightsum[N] + s->weights[n];
    Q[N] = f->sign * (v[m] - bc_sum_fi(s, fm)) / weightsum[N];
    for (int p = 0; p < D; ++p) {
      if (p == N) continue;
      for (int n = 1; n < s->Q; ++n) {
        momentum[p] = momentum[p] + fm[(size_t)n * S] * s->ci[n][p];
        if (incoming(s, f, n) && s->ci[n][p] != 0) weightsum[p] = weightsum[p] + s->weights[n];
      }
      Q[p] = -momentum[p] / weightsum[p];
    }
    bc_distribute(s, f, fm, Q);
  }
}

/* BCApplyNeumannNode, lbm_bc.F90:1533-1593.  mvals(m,p) = v[p*S + m]. */
static void bc_neumann_node(const txo_state *s, const txo_face *f, double *fi, const double *frc, const double *v) {
  const int D = s->D, S = s->S;
  for (int m = 0; m < S; ++m) {
    double *fm = fi + m;
    double Q[3] = {0., 0., 0.}, weightsum[3] = {0., 0., 0.}, momentum[3] = {0., 0., 0.};
    for (int p = 0; p < D; ++p) {
      for (int n = 1; n < s->Q; ++n) {
        momentum[p] = momentum[p] + fm[(size_t)n * S] * s->ci[n][p];
        if (incoming(s, f, n) && s->ci[n][p] != 0) weightsum[p] = weightsum[p] + s->weights[n];
      }
      Q[p] = (v[p * S + m] - frc[p * S + m] / 2. - momentum[p]) / weightsum[p];
    }
    bc_distribute(s, f, fm, Q);
  }
}

/* BCApplyVelocityNode, lbm_bc.F90:1793-1865.  uvals(1,p) = v[p*S]: every component takes the
 * velocity row of the first one. */
static void bc_velocity_node(const txo_state *s, const txo_face *f, double *fi, const double *frc, const double *v) {
  const int D = s->D, S = s->S, N = f->axis;
  for (int m = 0; m < S; ++m) {
    double *fm = fi + m;
    double Q[3] = {0., 0., 0.}, weightsum[3] = {0., 0., 0.}, momentum[3] = {0., 0., 0.};
    for (int n = 1; n < s->Q; ++n) {
      momentum[N] = momentum[N] + fm[(size_t)n * S] * s->ci[n][N];
      if (incoming(s, f, n)) weightsum[N] = weightsum[N] + s->weights[n];
    }
    const double uN = v[N * S];
    Q[N] = (bc_sum_fi(s, fm) * uN - momentum[N] - frc[N * S + m] / 2.) / (1. - f->sign * uN) / weightsum[N];
    const double rho = bc_sum_fi(s, fm) + weightsum[N] * Q[N] * f->sign;
    for (int p = 0; p < D; ++p) {
      if (p == N) continue;
      for (int n = 1; n < s->Q; ++n) {
        momentum[p] = momentum[p] + fm[(size_t)n * S] * s->ci[n][p];
        if (incoming(s, f, n) && s->ci[n][p] != 0) weightsum[p] = weightsum[p] + s->weights[n];
      }
      Q[p] = (rho * v[p * S] - frc[p * S + m] / 2. - momentum[p]) / weightsum[p];
    }
    bc_distribute(s, f, fm, Q);
  }
}

/* BCApplyReflectingD3/D2, lbm_bc.F90:825-1073: fi(:,n) = fi(:,p) for every incoming n and every p
 * (ascending, assignments in sequence) that passes the face's own test.  The tests are restated
 * face by face because two of them are not the mirror rule: XM in 3-D compares ci(n,X) with
 * -ci(p,Z) (:849), and XM in 2-D reads ci(p,Z_DIRECTION) of a two-column array (:1001, out of
 * bounds in the reference -> rejected by txo_bc_supported). */
static int reflecting_match(const txo_state *s, int b, int n, int p) {
  const int(*c)[3] = s->ci;
  if (s->D == 3) switch (b) {
      case 0: return c[n][1] == c[p][1] && c[n][2] == c[p][2] && c[n][0] == -c[p][2];
      case 1: return c[n][1] == c[p][1] && c[n][2] == c[p][2] && c[n][0] == -c[p][0];
      case 2:
      case 3: return c[n][0] == c[p][0] && c[n][2] == c[p][2] && c[n][1] == -c[p][1];
      default: return c[n][0] == c[p][0] && c[n][1] == c[p][1] && c[n][2] == -c[p][2];
    }
  switch (b) {
    case 1: return c[n][1] == c[p][1] && c[n][0] == -c[p][0];
    case 2:
    case 3: return c[n][0] == c[p][0] && c[n][1] == -c[p][1];
    default: return 0;
  }
}
int txo_bc_supported(const txo_state *s) {
  return !(s->D == 2 && s->cfg.bc_flags[TXG_BOUNDARY_XM] == TXG_BC_REFLECTING);
}
/* the (n <- p) assignment list of one reflecting face, in execution order; returns the count */
int txo_reflecting_pairs(const txo_state *s, int b, int *n_out, int *p_out) {
  const txo_face f = face_of(s, b);
  int k = 0;
  for (int n = 1; n < s->Q; ++n) {
    if (f.sign * s->ci[n][f.axis] <= 0) continue;
    for (int p = 1; p < s->Q; ++p)
      if (reflecting_match(s, b, n, p)) {
        n_out[k] = n;
        p_out[k] = p;
        ++k;
      }
  }
  return k;
}

/* BCApply, lbm_bc.F90:781-807: every BC type once, in the order of its first face; inside a type
 * the faces in the order XM, XP, YM, YP, ZM, ZP (BCApplyDirichletD3 etc.). */
static void bc_apply(txo_state *s) {
  int done[16] = {0};
  done[TXG_BC_PERIODIC] = 1;
  done[TXG_BC_NULL] = 1; /* select case has no branch for BC_NULL */
  const int nbcs = s->D * s->S, SD = s->S * s->D, SQ = s->S * s->Q;
  for (int side = 0; side < 6; ++side) {
    const int type = side < 2 * s->D ? s->cfg.bc_flags[side] : TXG_BC_NULL;
    if (type < 0 || type > 15 || done[type]) continue;
    done[type] = 1;
    for (int b = 0; b < 2 * s->D; ++b) {
      if (s->cfg.bc_flags[b] != type) continue;
      const txo_face f = face_of(s, b);
      int rn[64], rp[64], nr = 0;
      if (type == TXG_BC_REFLECTING) nr = txo_reflecting_pairs(s, b, rn, rp);
      for (int bb = 0; bb < f.n2; ++bb)
        for (int a = 0; a < f.n1; ++a) {
          int x[3] = {0, 0, 0};
          face_node(&f, a, bb, x);
          if (WALLS(s, x[0], x[1], x[2]) != 0.) continue;
          double *fi = &FI(s, 0, 0, x[0], x[1], x[2]);
          const double *frc = s->forces + gnode(s, x[0], x[1], x[2]) * (size_t)SD;
          const double *v = s->bc_vals[b] ? s->bc_vals[b] + ((size_t)bb * f.n1 + a) * nbcs : NULL;
          (void)SQ;
          if (type == TXG_BC_REFLECTING) {
            for (int e = 0; e < nr; ++e)
              for (int m = 0; m < s->S; ++m) fi[(size_t)rn[e] * s->S + m] = fi[(size_t)rp[e] * s->S + m];
          } else if (type == TXG_BC_DIRICHLET) {
            bc_dirichlet_node(s, &f, fi, v);
          } else if (type == TXG_BC_NEUMANN) {
            bc_neumann_node(s, &f, fi, frc, v);
          } else if (type == TXG_BC_VELOCITY) {
            bc_velocity_node(s, &f, fi, frc, v);
          }
        }
    }
  }
}

/* ---------------------------------------------------------------- time step
 * LBMRun2 body, lbm.F90:286-361:
 * FlowCollision; DistributionCommunicateFi; FlowStream (= BCPreStream + DistributionStream,
 * lbm_flow.F90:810-814); FlowBounceback; FlowApplyBCs (lbm_flow.F90:1958-1991 = FlowCalcRhoForces
 * [density, BCApplyDirichletToRho, forces; :445-456], BCApply, BCUpdateRho; the outlet value updates
 * at its top are host-side set-up of the face arrays and not part of this restatement);
 * FlowUpdateFlux (:458-464). */
static void apply_bcs(txo_state *s) {
  bc_pressure_outlet_update(s);
  calc_density(s);
  bc_dirichlet_to_rho(s);
  calc_forces(s);
  bc_apply(s);
  bc_update_rho(s);
}
void txo_step(txo_state *s, int nsteps) {
  for (int it = 0; it < nsteps; ++it) {
    flow_collision(s);
    communicate_fi(s);
    if (s->prestream) bc_prestream(s);
    stream(s);
    bounceback(s);
    apply_bcs(s);
    calc_flux_into(s, s->flux);
    update_ue(s);
  }
}

/* the phases individually, for white-box tests */
void txo_phase_collision(txo_state *s) { flow_collision(s); }
void txo_phase_communicate_fi(txo_state *s) { communicate_fi(s); }
void txo_phase_stream(txo_state *s) {
  if (s->prestream) bc_prestream(s);
  stream(s);
}
void txo_phase_bounceback(txo_state *s) { bounceback(s); }
void txo_phase_apply_bcs(txo_state *s) { apply_bcs(s); }
void txo_phase_update_flux(txo_state *s) {
  calc_flux_into(s, s->flux);
  update_ue(s);
}

/* ---------------------------------------------------------------- export (natural order, owned only) */
void txo_get_fi(const txo_state *s, double *out) {
  const int SQ = s->S * s->Q;
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i)
        memcpy(out + (((size_t)k * s->NY + j) * s->NX + i) * SQ, &FI(s, 0, 0, i, j, k), sizeof(double) * SQ);
}
void txo_get_rho(const txo_state *s, double *out) {
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i)
        for (int m = 0; m < s->S; ++m) out[(((size_t)k * s->NY + j) * s->NX + i) * s->S + m] = RHO(s, m, i, j, k);
}
static void get_sd(const txo_state *s, const double *a, double *out) {
  const int SD = s->S * s->D;
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i)
        memcpy(out + (((size_t)k * s->NY + j) * s->NX + i) * SD, a + gnode(s, i, j, k) * SD, sizeof(double) * SD);
}
void txo_get_u(const txo_state *s, double *out) { get_sd(s, s->flux, out); }
void txo_get_forces(const txo_state *s, double *out) { get_sd(s, s->forces, out); }

/* FlowUpdateDiagnostics -> FlowUpdateDiagnosticsD3/D2, lbm_flow.F90:603-758.
 * rhot, prs: [node]; velt: [node][d].  Solid nodes: prs = null_pressure, rhot and
 * velt keep their initial zero. */
void txo_diagnostics(txo_state *s, double *rhot, double *prs, double *velt) {
  const int S = s->S, D = s->D;
  const size_t ng = (size_t)s->gnx * s->gny * s->gnz;
  double *u = (double *)calloc(ng * S * D, sizeof(double));
  calc_flux_into(s, u);
  if (s->cfg.use_nonideal_eos) eos_apply(s);
  const double *psi = s->cfg.use_nonideal_eos ? s->psi : s->rho;
  for (int k = 0; k < s->NZ; ++k)
    for (int j = 0; j < s->NY; ++j)
      for (int i = 0; i < s->NX; ++i) {
        const size_t o = ((size_t)k * s->NY + j) * s->NX + i;
        if (WALLS(s, i, j, k) != 0.) {
          prs[o] = s->cfg.null_pressure;
          rhot[o] = 0.;
          for (int d = 0; d < D; ++d) velt[o * D + d] = 0.;
          continue;
        }
        double rt = 0.;
        for (int m = 0; m < S; ++m) rt += RHO(s, m, i, j, k) * s->cfg.mm[m];
        rhot[o] = rt;
        double p = rt / 3.;
        if (s->cfg.use_nonideal_eos || S > 1) {
          const double *ps = psi + rgnode(s, i, j, k) * (size_t)S;
          for (int m = 0; m < S; ++m) {
            double acc = 0.;
            for (int mp = 0; mp < S; ++mp) acc += s->cfg.gf[m][mp] * ps[mp];
            p = p + s->c_0 / 2. * ps[m] * acc;
          }
        }
        prs[o] = p;
        double *un = u + gnode(s, i, j, k) * (size_t)(S * D);
        for (int m = 0; m < S; ++m)
          for (int d = 0; d < D; ++d) un[d * S + m] = un[d * S + m] + .5 * FRC(s, m, d, i, j, k);
        for (int d = 0; d < D; ++d) {
          double acc = 0.;
          for (int m = 0; m < S; ++m) acc += un[d * S + m] * s->cfg.mm[m];
          velt[o * D + d] = acc / rt;
        }
      }
  free(u);
}

/* DistributionCalcDeltaNorm (fi variant), lbm_distribution_function.F90:809-833:
 * || (fi_old - fi) / fi ||_inf over the global (owned) vector. */
double txo_delta_norm(txo_state *s) {
  const size_t n = (size_t)s->NX * s->NY * s->NZ * s->S * s->Q;
  double *cur = (double *)malloc(n * sizeof(double));
  txo_get_fi(s, cur);
  double norm = 1.e99;
  if (s->fi_old) {
    norm = 0.;
    for (size_t a = 0; a < n; ++a) {
      double v = fabs((s->fi_old[a] - cur[a]) / cur[a]);
      if (v > norm || v != v) norm = v;
    }
    free(s->fi_old);
  }
  s->fi_old = cur;
  return norm;
}

/* lattice tables, for table-identity tests */
void txo_get_lattice(const txo_state *s, int *ci /*Q*3*/, double *weights, int *opp, double *mt /*Q*Q rows of M*/,
                     double *mmt, double *ffw /*41*/) {
  for (int n = 0; n < s->Q; ++n) {
    for (int d = 0; d < 3; ++d) ci[n * 3 + d] = s->ci[n][d];
    weights[n] = s->weights[n];
    opp[n] = s->opposites[n];
    mmt[n] = s->mmt_mrt[n];
    for (int i = 0; i < s->Q; ++i) mt[n * s->Q + i] = s->mt_mrt[n][i];
  }
  memcpy(ffw, s->ffw, sizeof s->ffw);
}

/* replay_bubble_2d.c -- the reference's shipped regression test (tests/bubble_2D, `make test`) driven from
 * compiled code through the C ABI of include/taxila_gpu.h, exactly the call sequence an ISO_C_BINDING shim
 * makes from lbm.F90: LBMSetUp/LBMInit2 (FlowSetUp, walls, LBMInitializeState, FlowFiInit,
 * FlowUpdateMoments), then LBMRun2's loop body as its six procedure calls per step (lbm.F90:286-361), then
 * the output of fi.  The result is compared with the reference's own golden file the way
 * src/testing/check_solution.py does (max |a - b| < 1e-5), and at round-off.
 *
 *   gcc -O2 -Iinclude shim/replay_bubble_2d.c -Ltaxila-lbm_b200 -ltaxila_gpu -Wl,-rpath,$PWD/taxila-lbm_b200 -lm -o replay
 *   ./replay tests/golden/bubble_2D_fi001.dat
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "taxila_gpu.h"

#define NXG 128
#define NYG 128
#define Q 9
#define S 2
#define NSTEPS 100 /* -npasses 100, tests/bubble_2D/input_data */

static void die(txg_handle h, const char *what, int rc) {
  fprintf(stderr, "%s failed with code %d: %s\n", what, rc, txg_last_error(h));
  exit(2);
}
#define CALL(h, f, ...)                     \
  do {                                      \
    int rc_ = f(__VA_ARGS__);               \
    if (rc_) die(h, #f, rc_);               \
  } while (0)

static double be_double(const unsigned char *p) {
  unsigned char b[8];
  for (int i = 0; i < 8; ++i) b[i] = p[7 - i];
  double v;
  memcpy(&v, b, 8);
  return v;
}

int main(int argc, char **argv) {
  const char *golden = argc > 1 ? argv[1] : "tests/golden/bubble_2D_fi001.dat";

  /* options of tests/bubble_2D/input_data */
  txg_config cfg;
  txg_config_defaults(&cfg);
  cfg.ndims = 2;
  cfg.discretization = TXG_D2Q9_DISCRETIZATION;
  cfg.ncomponents = S;
  cfg.NX = NXG, cfg.NY = NYG, cfg.NZ = 1;
  cfg.zs = 0, cfg.zl = 1;
  cfg.periodic[0] = cfg.periodic[1] = 1;
  cfg.stencil_size_rho = 1;
  cfg.relaxation_mode = TXG_RELAXATION_MODE_SRT;
  cfg.tau[0] = cfg.tau[1] = 1.0;
  cfg.mm[0] = cfg.mm[1] = 1.0;
  cfg.gf[0][1] = cfg.gf[1][0] = 0.1; /* -g_12 -g_21 */
  cfg.fluidfluid_forces = 1;
  cfg.isotropy_order = 4;

  txg_handle h = NULL;
  CALL(NULL, txg_create, &h, &cfg, 0);

  /* walls(rgxs:rgxe, rgys:rgye): no walls; rho(S, rg..): tests/bubble_2D/initialize_state.F90:187-203 */
  const int gx = NXG + 2, gy = NYG + 2;
  double *walls = calloc((size_t)gx * gy, sizeof(double));
  double *rho = calloc((size_t)gx * gy * S, sizeof(double));
  const int lx = (NXG + 1) / 2 - 26, rx = (NXG + 1) / 2 + 26, ly = (NYG + 1) / 2 - 26, ry = (NYG + 1) / 2 + 26;
  for (int j = 1; j <= NYG; ++j)
    for (int i = 1; i <= NXG; ++i) {
      const int in = i >= lx && i <= rx && j >= ly && j <= ry;
      double *r = rho + ((size_t)j * gx + i) * S; /* ghost width 1: owned (i,j) 1-based sits at [j][i] */
      r[0] = in ? 0.03 : 0.97;                    /* -rho_inner 0.03,0.97  -rho_outer 0.97,0.03 */
      r[1] = in ? 0.97 : 0.03;
    }
  CALL(h, txg_set_walls, h, walls);
  CALL(h, txg_set_rho_u, h, rho, NULL);
  CALL(h, txg_fi_init, h);
  CALL(h, txg_update_moments, h);
  for (int step = 0; step < NSTEPS; ++step) {
    CALL(h, txg_collision, h);
    CALL(h, txg_communicate_fi, h);
    CALL(h, txg_stream, h);
    CALL(h, txg_bounceback, h);
    CALL(h, txg_apply_bcs, h);
    CALL(h, txg_update_flux, h);
  }
  double *fi = calloc((size_t)gx * gy * Q * S, sizeof(double)); /* fi(S, 0:b, gxs:gxe, gys:gye) */
  CALL(h, txg_get_fi, h, fi);

  /* golden: PETSc binary Vec, big-endian, natural ordering (y, x, n, m) */
  FILE *fp = fopen(golden, "rb");
  if (!fp) {
    fprintf(stderr, "cannot open %s\n", golden);
    return 2;
  }
  unsigned char hdr[8];
  if (fread(hdr, 1, 8, fp) != 8) return 2;
  const long n = ((long)hdr[4] << 24) | (hdr[5] << 16) | (hdr[6] << 8) | hdr[7];
  if (n != (long)NXG * NYG * Q * S) {
    fprintf(stderr, "golden holds %ld values, expected %d\n", n, NXG * NYG * Q * S);
    return 2;
  }
  unsigned char *raw = malloc((size_t)n * 8);
  if (fread(raw, 8, (size_t)n, fp) != (size_t)n) return 2;
  fclose(fp);
  double maxdiff = 0., mass[S] = {0., 0.};
  for (int j = 0; j < NYG; ++j)
    for (int i = 0; i < NXG; ++i)
      for (int q = 0; q < Q; ++q)
        for (int m = 0; m < S; ++m) {
          const double a = fi[((((size_t)(j + 1) * gx + (i + 1)) * Q + q) * S) + m];
          const double b = be_double(raw + 8 * ((((size_t)j * NXG + i) * Q + q) * S + m));
          if (fabs(a - b) > maxdiff) maxdiff = fabs(a - b);
          mass[m] += a;
        }
  printf("bubble_2D, %d steps through the C ABI: max |fi - golden| = %.3e  mass = %.6f %.6f\n", NSTEPS, maxdiff, mass[0],
         mass[1]);
  CALL(h, txg_destroy, h);
  free(walls), free(rho), free(fi), free(raw);
  if (!(maxdiff < 1e-5)) {
    printf("FAIL (check_solution.py eps = 1e-5)\n");
    return 1;
  }
  printf("%s\n", maxdiff <= 1e-12 ? "PASS (round-off)" : "PASS (eps 1e-5 only)");
  return maxdiff <= 1e-12 ? 0 : 3;
}

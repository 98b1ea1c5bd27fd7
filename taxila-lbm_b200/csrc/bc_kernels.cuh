// bc_kernels.cuh -- surface kernels of the flow step: external face boundary conditions
// (lbm_bc.F90) and the free-slip (WALL_NORMAL_X/Y/Z = 900-902) part of DistributionBouncebackD*.
//
// These run over O(N^(2/3)) face nodes once per step, next to hot kernels that move ~600 B for every
// node of the box: they are written for clarity (run-time lattice tables, one thread per face node,
// all components in a loop) and follow the reference's evaluation order statement by statement.
#pragma once
#include "kernels.cuh"
#include "specular_table.h"

namespace txg {

// (FaceDesc and face_node: kernels.cuh -- the face variants of the hot kernels use them too)

// BCApplyDirichletToRho -> _D2/_D3 (lbm_bc.F90:252-434): rho(m, face node) = vals(m) on the fluid
// nodes of a Dirichlet face, before the density halo and the forces.
__global__ void k_bc_dirichlet_rho(Grid g, Phys p, int S, FaceDesc fd, const double *__restrict__ vals, int nbcs,
                                   double *__restrict__ rho, double *__restrict__ rho_true,
                                   const uint32_t *__restrict__ nbmask) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long pos;
  if (!face_node(g, fd, nbmask, idx, pos)) return;
  for (int m = 0; m < S; ++m) {
    const double v = vals[idx * nbcs + m];
    if (p.eos) {
      rho_true[(long long)m * g.fs + pos] = v;
      rho[(long long)m * g.fs + pos] = eos_psi(p, m, v);
    } else {
      rho[(long long)m * g.fs + pos] = v;
    }
  }
}

// FlowUpdateBCPressureOutletD2/D3 + FlowUpdateDensityFromPressure (lbm_flow.F90:1993-2263) on one face, two
// components: the Dirichlet densities of every fluid face node are re-derived from the face's pressure and
// the phase fraction rho_1 / sum(rho) of the node one step inside (of the node itself when that one is
// solid).  The reference reads dist%rho as the previous step left it, which is the sum of the populations
// of that state everywhere (DistributionCalcDensity, BCUpdateRho): f is the buffer BEFORE this step's collide.
__global__ void k_bc_pressure_outlet(Grid g, int Q, FaceDesc fd, double *__restrict__ vals, int nbcs,
                                     const double *__restrict__ f, const uint32_t *__restrict__ nbmask,
                                     double pressure, double g21) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long pos;
  if (!face_node(g, fd, nbmask, idx, pos)) return;
  int x[3] = {0, 0, 0};
  x[fd.axis] = fd.coord + fd.sign;
  x[fd.t1] = (int)(idx % fd.n1);
  x[fd.t2] = (int)(idx / fd.n1);
  const long long o = (long long)x[2] * g.plane + (long long)x[1] * g.NX + x[0];
  if (!(nbmask[o] >> 31)) pos = pos_of(g, o + (long long)g.Rz * g.plane);
  double rho[2];
  for (int m = 0; m < 2; ++m) {
    double a = 0.;
    for (int n = 0; n < Q; ++n) a += f[(long long)(m * Q + n) * g.fs + pos];
    rho[m] = a;
  }
  const double rho1frac = rho[0] / (rho[0] + rho[1]);
  const double eps = 1.e-10;
  double r0, r1;
  if (rho1frac < eps) {
    r0 = 0.;
    r1 = pressure * 3.;
  } else if (rho1frac > 1 - eps) {
    r0 = pressure * 3.;
    r1 = 0.;
  } else {
    const double alpha = 1. / (1. / rho1frac - 1.);
    r0 = (-(1. + alpha) / 3. + sqrt((1. + alpha) / 3. * (1. + alpha) / 3. + 4. * 6.0 * g21 * alpha * pressure)) / (2 * 6.0 * g21);
    r1 = r0 / alpha;
  }
  vals[idx * nbcs + 0] = r0;
  vals[idx * nbcs + 1] = r1;
}

// BCApplyReflectingD3 / D2 (lbm_bc.F90:809-1073) on one face: fi(:, n) = fi(:, p) for the face's (n <- p) list, in
// sequence (a later assignment may read what an earlier one wrote).  The list is formed on the host from the
// reference's own tests, face by face (flow.cu reflecting_pairs).
struct ReflectPairs {
  int count;
  unsigned char n[64], p[64];
};
__global__ void k_bc_reflect(Grid g, int Q, int S, FaceDesc fd, ReflectPairs rp, double *__restrict__ f,
                             const uint32_t *__restrict__ nbmask) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long pos;
  if (!face_node(g, fd, nbmask, idx, pos)) return;
  for (int m = 0; m < S; ++m) {
    double *fm = f + (long long)m * Q * g.fs + pos;
    for (int e = 0; e < rp.count; ++e) fm[(long long)rp.n[e] * g.fs] = fm[(long long)rp.p[e] * g.fs];
  }
}

// MASK_STALE on the fluid nodes of one face: set for reflecting faces, then cleared for the faces BCUpdateRho visits
__global__ void k_bc_mark_stale(Grid g, FaceDesc fd, uint32_t *__restrict__ nbmask, uint32_t *__restrict__ lmask, int set) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long pos;
  if (!face_node(g, fd, nbmask, idx, pos)) return;
  int x[3] = {0, 0, 0};
  x[fd.axis] = fd.coord;
  x[fd.t1] = (int)(idx % fd.n1);
  x[fd.t2] = (int)(idx / fd.n1);
  const long long o = (long long)x[2] * g.plane + (long long)x[1] * g.NX + x[0];
  if (set) {
    atomicOr(nbmask + o, MASK_STALE);
    atomicOr(lmask + pos, MASK_STALE);
  } else {
    atomicAnd(nbmask + o, ~MASK_STALE);
    atomicAnd(lmask + pos, ~MASK_STALE);
  }
}

// BCApply -> BCApply{Dirichlet,Neumann,Velocity}D* -> ...Node (lbm_bc.F90:781-807,1075-1865) on one
// face.  f: the populations after streaming and bounce-back (the incoming directions of a face node
// hold what the bounce-back off the 999 ghost layer put there); F: forces of FlowCalcRhoForces.
// vals(nbcs) of a node is viewed as (S, ndims): pvals(m,1), mvals(m,d), uvals(1,d).
__global__ void k_bc_apply(Grid g, LatticeTab lt, int S, FaceDesc fd, const double *__restrict__ vals,
                           double *__restrict__ f, const double *__restrict__ Fbuf,
                           const uint32_t *__restrict__ nbmask) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long pos;
  if (!face_node(g, fd, nbmask, idx, pos)) return;
  const int Q = lt.Q, D = lt.D, N = fd.axis;
  const int nbcs = S * D;
  const double *v = vals + idx * nbcs;
  for (int m = 0; m < S; ++m) {
    double *fm = f + (long long)m * Q * g.fs + pos;
    double fi[19];
    for (int n = 0; n < Q; ++n) fi[n] = fm[(long long)n * g.fs];
    double Qc[3] = {0., 0., 0.}, weightsum[3] = {0., 0., 0.}, momentum[3] = {0., 0., 0.};
    double sumf = 0.;
    for (int n = 0; n < Q; ++n) sumf += fi[n];
    // weightsum(d): incoming directions with a component along d; momentum(d) = sum_n fi(n) c_n,d
    for (int d = 0; d < D; ++d)
      for (int n = 1; n < Q; ++n) {
        momentum[d] = momentum[d] + fi[n] * (double)lt.c[n][d];
        if (fd.sign * lt.c[n][N] == 1 && lt.c[n][d] != 0) weightsum[d] = weightsum[d] + lt.w[n];
      }
    double Fm[3] = {0., 0., 0.};
    for (int d = 0; d < D; ++d) Fm[d] = Fbuf[(long long)(m * D + d) * g.fs + pos];
    if (fd.type == 3) {  // BC_DIRICHLET: BCApplyDirichletNode (lbm_bc.F90:1273-1333)
      Qc[N] = (double)fd.sign * (v[m] - sumf) / weightsum[N];
      for (int d = 0; d < D; ++d)
        if (d != N) Qc[d] = -momentum[d] / weightsum[d];
    } else if (fd.type == 4) {  // BC_NEUMANN: BCApplyNeumannNode (:1533-1593)
      for (int d = 0; d < D; ++d) Qc[d] = (v[d * S + m] - Fm[d] / 2. - momentum[d]) / weightsum[d];
    } else {  // BC_VELOCITY: BCApplyVelocityNode (:1793-1865); every component takes uvals(1,:)
      const double uN = v[N * S];
      Qc[N] = (sumf * uN - momentum[N] - Fm[N] / 2.) / (1. - (double)fd.sign * uN) / weightsum[N];
      const double rho = sumf + weightsum[N] * Qc[N] * (double)fd.sign;
      for (int d = 0; d < D; ++d)
        if (d != N) Qc[d] = (rho * v[d * S] - Fm[d] / 2. - momentum[d]) / weightsum[d];
    }
    for (int n = 1; n < Q; ++n)
      if (fd.sign * lt.c[n][N] == 1) {
        double acc = 0.;
        for (int d = 0; d < D; ++d) acc += (double)lt.c[n][d] * Qc[d];
        fm[(long long)n * g.fs] = fi[n] + lt.w[n] * acc;
      }
  }
}

// Free-slip walls.  After the push (which treats every solid neighbour as a plain bounce-back wall) and
// the z halo, the slots listed in the table are rewritten: dst <- value currently in src (or 0).  The
// table is built on the host from the geometry alone (specular_table.h); gather and scatter
// are separate launches because a slot can be both a source and a destination.
__global__ void k_specular_gather(const double *__restrict__ f, const uint32_t *__restrict__ src,
                                  double *__restrict__ tmp, long long n, int S, long long comp_stride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * S) return;
  const long long e = idx % n;
  const int m = (int)(idx / n);
  const uint32_t s = src[e];
  tmp[idx] = s == SPEC_ZERO ? 0. : f[(long long)m * comp_stride + s];
}
__global__ void k_specular_scatter(double *__restrict__ f, const uint32_t *__restrict__ dst,
                                   const double *__restrict__ tmp, long long n, int S, long long comp_stride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * S) return;
  const long long e = idx % n;
  const int m = (int)(idx / n);
  f[(long long)m * comp_stride + dst[e]] = tmp[idx];
}

}  // namespace txg

// kernels.cuh -- device code of the flow hot path (sm_100a).
//
// Storage (all fp64, one slab per GPU).  Only FLUID nodes are stored: every per-node array is indexed by
// the node's position in the ascending list of fluid nodes of the extended slab (owned planes plus R
// ghost planes each side in 3-D), so a porous medium costs neither memory nor partially used DRAM
// sectors for its solid voxels, and every population read of a warp is one contiguous run.
//   P     [nE+1]      u32   extended node index oe = (z+Rz)*plane + y*NX + x  ->  number of fluid nodes
//                           before oe, i.e. the position of oe if it is fluid      (nullptr: identity)
//   list  [nstore]    u32   position -> oe                                         (nullptr: identity)
//   f     [S][Q][fs]        populations in the reference's own sense: slot (m,n,pos) is fi(m,n,X), the value
//                           node X holds AFTER streaming and bounce-back.  The collide kernel PUSHES: the
//                           post-collision value leaving X along c_n goes into slot (n, pos(X+c_n)) of the
//                           other buffer, or -- when X+c_n is solid -- into slot (opp(n), pos(X)) (half-way
//                           bounce-back completed in the same step).  Reads of f are
//                           position-aligned and need no mask.  Pushes that leave the slab land in the
//                           ghost-plane positions; the halo exchange moves them into the neighbour's
//                           boundary plane.  fs = nstore rounded up (+ padding): the stride of every array.
//   rho   [S][fs]           per-component density (psi when a non-ideal EOS is on); ghost planes by halo
//   lmask [fs]        u32   per OWNED position: bit n = neighbour X+c_n is solid, bit 30 = X has a wall
//                           record (some lattice neighbour solid or some gradient-stencil entry inactive),
//                           bits 28/29 = X sits on the periodic x = 0 / x = NX-1 face
//   nbr   [NCEN][fs]  u32   per OWNED position: position of X + c_n for the centre directions (c_x = 0)
//   wallrec [S*D+D][fs]     per owned position with bit 30: A[m][d] = sum_n w_n gw(mineral(X+c_n),m) c_n,d
//                           (fluid-solid force = -rho_m A) and 1/W[d] of the gradient normalisation
//   cls   [NZl+2Rz][NY+2R][NX+2R]  u8 node class incl. ghosts, straight from the host walls(rg..) array
//   nbmask [NZl][NY][NX]    u32 dense copy of the masks (bit 31 = solid) for the set-up / export kernels
//   ffmask [NW][NZl][NY][NX] u32 words: bit e = fluid-fluid stencil entry e active (isotropy order > 4)
//
// Reference loops each kernel replaces are cited at the kernel.
#pragma once
#include <cstdint>

#include "lattice.cuh"

namespace txg {

struct Grid {
  int NX, NY, NZl;  // owned slab
  int R, Rz;        // ghost width of rho/cls in x,y (R) and z (Rz = R in 3-D, 0 in 2-D)
  int perx, pery;   // periodic flags (non-periodic neighbours are class 255 and never dereferenced)
  long long plane;   // NX*NY
  long long nnodes;  // NZl*plane (owned nodes)
  long long nE;      // (NZl+2Rz)*plane (extended slab)
  long long fs;      // stride of every position-indexed array (>= nstore + 1)
  long long own0, own1;  // positions of the owned fluid nodes: [own0, own1)
  const uint32_t *P;     // [nE+1] extended node index -> position (nullptr: identity)
  const uint32_t *list;  // [nstore] position -> extended node index (nullptr: identity)
  int cnx, cny;      // NX+2R, NY+2R
};

// position of extended node index oe (for a solid node: the position of the next fluid node)
__device__ __forceinline__ long long pos_of(const Grid &g, long long oe) {
  return g.P ? (long long)__ldg(g.P + oe) : oe;
}

// One face of the box in the numbering of lbm_definitions.h:45-50 (0-based: XM, XP, YM, YP, ZM, ZP).
// The node routines of lbm_bc.F90 are written for one boundary and rotated onto the others with
// DiscSetLocalDirections (lbm_discretization_d3q19.F90:566-714, lbm_discretization_d2q9.F90:372-434);
// all they take from the rotation is ci(directions(local_normal), normal axis) -- the inward normal
// sign on every boundary -- and the set of tangential axes, each handled independently.
struct FaceDesc {
  int axis, sign, coord;  // normal axis, inward sign, coordinate of the face plane (owned, local in z)
  int t1, t2, n1, n2;     // tangential axes (t1 fastest in the face array) and their local extents
  int type;               // TXG_BC_DIRICHLET / NEUMANN / VELOCITY (lbm_definitions.h:32-34)
};

// (owned dense index, position) of face entry idx; false for a solid node or past the end
__device__ __forceinline__ bool face_node(const Grid &g, const FaceDesc &fd, const uint32_t *__restrict__ nbmask,
                                          long long idx, long long &pos) {
  if (idx >= (long long)fd.n1 * fd.n2) return false;
  int x[3] = {0, 0, 0};
  x[fd.axis] = fd.coord;
  x[fd.t1] = (int)(idx % fd.n1);
  x[fd.t2] = (int)(idx / fd.n1);
  const long long o = (long long)x[2] * g.plane + (long long)x[1] * g.NX + x[0];
  if (nbmask[o] >> 31) return false;
  pos = pos_of(g, o + (long long)g.Rz * g.plane);
  return true;
}

constexpr int MAXS = 5;  // instantiated component counts: 1..5 (NMAX_COMPONENTS of the reference, lbm_definitions.h:71)

struct Phys {
  // per component
  double inv_tau[MAXS];        // SRT
  double mrt_rate[MAXS][19];   // MRT: s_r / ||M_r||^2 per moment row
  double mm[MAXS], d_k[MAXS], mmot[MAXS];
  double gf[MAXS][MAXS];
  // EOSApply (lbm_eos.F90:149-349), active when eos != 0.  eos_kind: TXG_EOS_DENSITY 1 (psi = rho), SC 2, PR 3,
  // THERMO 4.  SC / THERMO: eos_rho0, eos_psi0.  PR: a, b, R, T, alpha (host, :331-332) and c0g = c_0 * g_mm.
  int eos_kind[MAXS];
  double eos_rho0[MAXS], eos_psi0[MAXS];
  double pr_a[MAXS], pr_b[MAXS], pr_R[MAXS], pr_T[MAXS], pr_alpha[MAXS], pr_c0g[MAXS];
  int *eos_bad;                // device counter: PR inner square root went negative (lbm_eos.F90:337-341)
  double gvt[3];
  int nminerals;
  int fluidfluid, fluidsolid, body, eos;
  const double *gw;            // device [nminerals][S]
};

// ------------------------------------------------------------------ addressing helpers
struct NodeIdx {
  int x, y, z;      // owned coordinates
  long long o;      // z*plane + y*NX + x  (nbmask / export index)
  long long oe;     // (z+Rz)*plane + y*NX + x  (extended index)
  long long pos;    // position in the fluid list (valid for fluid nodes)
};

__device__ __forceinline__ bool node_of_thread(const Grid &g, int z0, int nz, NodeIdx &nd) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nz * g.plane) return false;
  int zz = (int)(idx / g.plane);
  int r = (int)(idx - (long long)zz * g.plane);
  nd.y = r / g.NX;
  nd.x = r - nd.y * g.NX;
  nd.z = z0 + zz;
  nd.o = (long long)nd.z * g.plane + r;
  nd.oe = nd.o + (long long)g.Rz * g.plane;
  nd.pos = pos_of(g, nd.oe);
  return true;
}

// wrapped neighbour coordinate (clamped when the direction is not periodic; such
// neighbours are solid in the masks and never read)
__device__ __forceinline__ int wrapc(int v, int N, int per) {
  if (v < 0) return per ? v + N : 0;
  if (v >= N) return per ? v - N : N - 1;
  return v;
}

// ------------------------------------------------------------------ populations of a node
// fi(:,:,X) after DistributionStreamD*/DistributionBouncebackD* (lbm_distribution_function.F90:560-784):
// with push storage simply the node's own slots.
template <class L, int S>
__device__ __forceinline__ void load_node(const Grid &g, const double *__restrict__ fA, const NodeIdx &nd,
                                          double (&f)[S][L::Q]) {
#pragma unroll
  for (int m = 0; m < S; ++m)
#pragma unroll
    for (int n = 0; n < L::Q; ++n) f[m][n] = __ldg(fA + (long long)(m * L::Q + n) * g.fs + nd.pos);
}

constexpr uint32_t MASK_SOLID = 0x80000000u;    // the node itself is solid
constexpr uint32_t MASK_WALLREC = 0x40000000u;  // the node has a wall record
constexpr uint32_t MASK_XLO = 0x10000000u;      // x = 0 and x periodic: the -x neighbours wrap
constexpr uint32_t MASK_XHI = 0x20000000u;      // x = NX-1 and x periodic: the +x neighbours wrap
// a fluid node of a BC_REFLECTING face (and of no Dirichlet / Neumann / velocity face): BCUpdateRho skips it
// (lbm_bc.F90:436-456), so collision, flux and exports take the density FlowCalcRhoForces stored before BCApply
constexpr uint32_t MASK_STALE = 0x08000000u;
constexpr uint32_t MASK_DIRS = 0x07ffffffu;     // bit n: neighbour X + c_n is solid

// ------------------------------------------------------------------ forces
// FlowCalcForces (lbm_flow.F90:760-808): F = 0; fluid-solid (LBMAddFluidSolidForcesD*,
// lbm_forcing.F90:1326-1421); body (LBMAddBodyForcesD*, :1440-1496); fluid-fluid
// (LBMAddFluidFluidForcesD*, :51-1299) -- in that order, entries in the reference's order.
// `rho_here` is this node's density; `psi` the stencil field (rho, or psi(rho) with an EOS).
template <class L, int S, int ISO>
__device__ __forceinline__ void forces(const Grid &g, const Phys &p, const double *__restrict__ psi,
                                       const uint8_t *__restrict__ cls, const uint32_t *__restrict__ ffmask,
                                       const NodeIdx &nd, uint32_t mask, const double (&rho_here)[S],
                                       double (&F)[S][L::D]) {
  constexpr int D = L::D;
#pragma unroll
  for (int m = 0; m < S; ++m)
#pragma unroll
    for (int d = 0; d < D; ++d) F[m][d] = 0.;

  if (p.fluidsolid && (mask & MASK_DIRS)) {
    const long long cbase = ((long long)(nd.z + g.Rz) * g.cny + (nd.y + g.R)) * g.cnx + (nd.x + g.R);
    static_for<1, L::Q>([&](auto n_) {
      constexpr int n = decltype(n_)::value;
      if ((mask >> n) & 1u) {
        const long long coff = ((long long)L::c(n, 2) * g.cny + L::c(n, 1)) * g.cnx + L::c(n, 0);
        const int id = cls[cbase + coff];
        // 0 < walls < 998 and a valid mineral (lbm_forcing.F90:1352); 800/900-902 index out of
        // bounds in the reference and contribute nothing here
        if (id >= 1 && id <= p.nminerals) {
          constexpr double w = L::fs_weight(n);
          static_for<0, D>([&](auto d_) {
            constexpr int d = decltype(d_)::value;
            if constexpr (L::c(n, d) != 0) {
#pragma unroll
              for (int m = 0; m < S; ++m)
                F[m][d] = F[m][d] - w * rho_here[m] * p.gw[(id - 1) * S + m] * (double)L::c(n, d);
            }
          });
        }
      }
    });
  }

  if (p.body) {
#pragma unroll
    for (int m = 0; m < S; ++m)
#pragma unroll
      for (int d = 0; d < D; ++d) F[m][d] = F[m][d] + p.gvt[d] * p.mm[m] * rho_here[m];
  }

  if (p.fluidfluid) {
    using FF = typename L::FF;
    constexpr int E = ff_entries<L>(ISO);
    constexpr int RAD = stencil_radius(ISO);
    double G[D][S], W[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      W[d] = 0.;
#pragma unroll
      for (int m = 0; m < S; ++m) G[d][m] = 0.;
    }
    // wrapped coordinates for every offset in [-RAD, RAD]
    int xi[2 * RAD + 1], yi[2 * RAD + 1];
#pragma unroll
    for (int a = -RAD; a <= RAD; ++a) {
      xi[a + RAD] = wrapc(nd.x + a, g.NX, g.perx);
      yi[a + RAD] = wrapc(nd.y + a, g.NY, g.pery);
    }
    double psi_here[S];
#pragma unroll
    for (int m = 0; m < S; ++m) psi_here[m] = p.eos ? psi[m * g.fs + nd.pos] : rho_here[m];

    uint32_t words[(E + 31) / 32];
    if constexpr (ISO != 4) {
#pragma unroll
      for (int w = 0; w < (E + 31) / 32; ++w) words[w] = ffmask[(long long)w * g.nnodes + nd.o];
    }
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      constexpr int dx = FF::off[e][0], dy = FF::off[e][1], dz = FF::off[e][2];
      bool active;
      if constexpr (ISO == 4) {
        constexpr int n = dir_of<L>(dx, dy, dz);  // order-4 offsets are the lattice directions
        active = !((mask >> n) & 1u);
      } else {
        active = (words[e / 32] >> (e % 32)) & 1u;
      }
      if (active) {
        constexpr double wgt = L::ffw(ISO, FF::L[e]);
        const long long nb = pos_of(g, (long long)(nd.z + g.Rz + dz) * g.plane + (long long)yi[dy + RAD] * g.NX + xi[dx + RAD]);
        double diff[S];
#pragma unroll
        for (int m = 0; m < S; ++m) diff[m] = __ldg(psi + m * g.fs + nb) - psi_here[m];
        if constexpr (dx != 0) {
#pragma unroll
          for (int m = 0; m < S; ++m) G[0][m] = G[0][m] + ((double)dx * wgt) * diff[m];
          W[0] = W[0] + wgt * (double)(dx * dx);
        }
        if constexpr (dy != 0) {
#pragma unroll
          for (int m = 0; m < S; ++m) G[1][m] = G[1][m] + ((double)dy * wgt) * diff[m];
          W[1] = W[1] + wgt * (double)(dy * dy);
        }
        if constexpr (D == 3 && dz != 0) {
#pragma unroll
          for (int m = 0; m < S; ++m) G[D - 1][m] = G[D - 1][m] + ((double)dz * wgt) * diff[m];
          W[D - 1] = W[D - 1] + wgt * (double)(dz * dz);
        }
      }
    });
    const double eps = (double)1.e-12f;  // default-real literal, lbm_forcing.F90:69
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (W[d] > eps) {
#pragma unroll
        for (int m = 0; m < S; ++m) {
          double acc = 0.;
#pragma unroll
          for (int mp = 0; mp < S; ++mp) acc += p.gf[m][mp] * (G[d][mp] / W[d]);
          F[m][d] = F[m][d] - 6.0 * psi_here[m] * acc;  // c_0 = 6 on both lattices
        }
      }
  }
}

// ------------------------------------------------------------------ moments
// density of the streamed populations, ascending n like sum(fi(:,:,i,j,k),2)
// (DistributionCalcDensityD*, lbm_distribution_function.F90:379-428)
template <class L, int S>
__device__ __forceinline__ void density(const double (&f)[S][L::Q], double (&rho)[S]) {
#pragma unroll
  for (int m = 0; m < S; ++m) {
    double a = 0.;
#pragma unroll
    for (int n = 0; n < L::Q; ++n) a += f[m][n];
    rho[m] = a;
  }
}

// momentum j (DistributionCalcFluxD*, :451-508) and the common velocity u'
// (FlowUpdateUED*, lbm_flow.F90:494-574)
template <class L, int S>
__device__ __forceinline__ void common_velocity(const Phys &p, const double (&f)[S][L::Q], const double (&rho)[S],
                                                const double (&F)[S][L::D], double (&up)[L::D]) {
  constexpr int D = L::D;
  double ue[S][D];
#pragma unroll
  for (int m = 0; m < S; ++m)
    static_for<0, D>([&](auto d_) {
      constexpr int d = decltype(d_)::value;
      double a = 0.;
      static_for<0, L::Q>([&](auto n_) {
        constexpr int n = decltype(n_)::value;
        if constexpr (L::c(n, d) != 0) a += f[m][n] * (double)L::c(n, d);
      });
      ue[m][d] = a + .5 * F[m][d];
    });
  double den = 0.;
#pragma unroll
  for (int m = 0; m < S; ++m) den += rho[m] * p.mmot[m];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double num = 0.;
#pragma unroll
    for (int m = 0; m < S; ++m) num += ue[m][d] * p.mmot[m];
    up[d] = num / den;
  }
}

// equilibrium (DiscretizationEquilf_D3Q19/D2Q9) for one component.  The reference divides by c_s2 = 1/3 and 2 c_s2^2; here the
// factors 3 and 4.5 are multiplied in, as collide1 does (a correctly rounded fp64 division is ~20 instructions and a branch: with the
// 37 divisions of the literal form FlowFiInit took 20 ms at 512^3, profiles/r2aj_e2e_sync_marks_before.json -> r2ak_e2e_sync_marks.json; the difference is one rounding).
template <class L>
__device__ __forceinline__ void equilibrium(double rho, double d_k, const double (&u)[L::D], double (&feq)[L::Q]) {
  double usqr = 0.;
#pragma unroll
  for (int d = 0; d < L::D; ++d) usqr += u[d] * u[d];
  feq[0] = rho * L::feq0(d_k, usqr);
  const double base = 1.5 * (1. - d_k) - 1.5 * usqr;
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    double udote = 0.;
    static_for<0, L::D>([&](auto d_) {
      constexpr int d = decltype(d_)::value;
      if constexpr (L::c(n, d) != 0) udote += (double)L::c(n, d) * u[d];
    });
    feq[n] = L::w(n) * rho * (base + 3. * udote + 4.5 * (udote * udote));
  });
}

// prefactor(m,n) = sum_d F(m,d)(c_n,d - u_d) / (rho_m c_s2)  (FlowFiBarEqPrefactor, lbm_flow.F90:836-851)
template <class L>
__device__ __forceinline__ void prefactor(double rho, const double (&F)[L::D], const double (&u)[L::D],
                                          double (&pref)[L::Q]) {
  const double inv = 3.0 / rho;  // 1 / (rho c_s2)
  static_for<0, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    double a = 0.;
#pragma unroll
    for (int d = 0; d < L::D; ++d) a += F[d] * ((double)L::c(n, d) - u[d]);
    pref[n] = a * inv;
  });
}

// EOSApply_Rho / _SC / _Thermo / _PR (lbm_eos.F90:183-349) for one value
__device__ __forceinline__ double eos_psi(const Phys &p, int m, double rho) {
  const int kind = p.eos_kind[m];
  if (kind == 2) return p.eos_rho0[m] * (1. - exp(-rho / p.eos_rho0[m]));
  if (kind == 4) return p.eos_psi0[m] * exp(-p.eos_rho0[m] / rho);
  if (kind == 3) {
    const double b = p.pr_b[m];
    const double tmp = 2. * (rho * p.pr_R[m] * p.pr_T[m] / (1. - b * rho) -
                             (p.pr_a[m] * p.pr_alpha[m] * (rho * rho)) / (1. + 2. * b * rho - (b * rho) * (b * rho)) - rho / 3.) /
                       p.pr_c0g[m];
    if (tmp < 0.) atomicAdd(p.eos_bad, 1);  // "PR EOS inner sqrt went negative": reported at the next synchronising call
    return sqrt(tmp);
  }
  return rho;
}

// ================================================================== kernels

// K3 fi_init (FlowFiInit lbm_flow.F90:923-934, FlowFeqBarD* :867-921): F from rho0, feq(rho0, u0),
// f = (1 - prefactor/2) feq, written into the node's own slots.  u0 is [S][D][nnodes] or null (= 0).
template <class L, int S, int ISO>
__global__ void __launch_bounds__(128) k_fi_init(Grid g, Phys p, double *__restrict__ fN,
                                                 const double *__restrict__ rho, const double *__restrict__ rho_true,
                                                 const double *__restrict__ u0, const uint32_t *__restrict__ nbmask,
                                                 const uint32_t *__restrict__ ffmask, const uint8_t *__restrict__ cls,
                                                 int z0, int nz) {
  NodeIdx nd;
  if (!node_of_thread(g, z0, nz, nd)) return;
  const uint32_t mask = nbmask[nd.o];
  if (mask >> 31) return;
  constexpr int Q = L::Q, D = L::D;
  double r[S], F[S][D];
#pragma unroll
  for (int m = 0; m < S; ++m) r[m] = rho_true[m * g.fs + nd.pos];
  forces<L, S, ISO>(g, p, rho, cls, ffmask, nd, mask, r, F);
#pragma unroll
  for (int m = 0; m < S; ++m) {
    double u[D], feq[Q], pref[Q];
#pragma unroll
    for (int d = 0; d < D; ++d) u[d] = u0 ? u0[(long long)(m * D + d) * g.nnodes + nd.o] : 0.;
    equilibrium<L>(r[m], p.d_k[m], u, feq);
    prefactor<L>(r[m], F[m], u, pref);
#pragma unroll
    for (int n = 0; n < Q; ++n) fN[(long long)(m * Q + n) * g.fs + nd.pos] = (1. - 0.5 * pref[n]) * feq[n];
  }
}

// K6 state/diagnostics export: from the populations and the current rho field compute
// what the reference holds after FlowUpdateMoments (rho, forces, common velocity u') and what
// FlowUpdateDiagnosticsD* (lbm_flow.F90:654-758) derives (rhot, prs, velt).  Any output may be null.
// Fsrc != null: the forces are read from the step's force buffer instead of being re-formed.
template <class L, int S, int ISO>
__global__ void __launch_bounds__(128) k_export(Grid g, Phys p, const double *__restrict__ fA,
                                                const double *__restrict__ rho, const uint32_t *__restrict__ nbmask,
                                                const uint32_t *__restrict__ ffmask, const uint8_t *__restrict__ cls,
                                                const double *__restrict__ Fsrc /*[S*D][fs] or null*/,
                                                const double *__restrict__ rho_stale /*[S][fs] or null: MASK_STALE nodes*/,
                                                int fsrc_faces /*bit b: Fsrc holds the nodes of face b only (0: every node)*/,
                                                double *__restrict__ rho_out /*[S][nnodes]*/,
                                                double *__restrict__ u_out /*[S][D][nnodes]*/,
                                                double *__restrict__ F_out /*[S][D][nnodes]*/,
                                                double *__restrict__ rhot, double *__restrict__ prs,
                                                double *__restrict__ velt /*[nnodes][D]: the host array's own layout*/, double null_pressure,
                                                int z0, int nz) {
  NodeIdx nd;
  if (!node_of_thread(g, z0, nz, nd)) return;
  const uint32_t mask = nbmask[nd.o];
  constexpr int Q = L::Q, D = L::D;
  if (mask >> 31) {
#pragma unroll
    for (int m = 0; m < S; ++m) {
      if (rho_out) rho_out[m * g.nnodes + nd.o] = 0.;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (u_out) u_out[(long long)(m * D + d) * g.nnodes + nd.o] = 0.;
        if (F_out) F_out[(long long)(m * D + d) * g.nnodes + nd.o] = 0.;
      }
    }
    if (rhot) rhot[nd.o] = 0.;
    if (prs) prs[nd.o] = null_pressure;
    if (velt)
#pragma unroll
      for (int d = 0; d < D; ++d) velt[(long long)nd.o * D + d] = 0.;
    return;
  }
  double f[S][Q], r[S], F[S][D], up[D];
  load_node<L, S>(g, fA, nd, f);
  density<L, S>(f, r);
  if (rho_stale && (mask & MASK_STALE)) {
#pragma unroll
    for (int m = 0; m < S; ++m) r[m] = __ldg(rho_stale + (long long)m * g.fs + nd.pos);
  }
  bool stored = Fsrc != nullptr;
  if (stored && fsrc_faces) {
    // fused step with face BCs: only the nodes of the BC faces have stored forces; elsewhere the sum of the populations
    // IS the density the forces were formed from, and they are re-formed like without BCs
    const int c[3] = {nd.x, nd.y, nd.z}, n[3] = {g.NX, g.NY, g.NZl};
    stored = false;
    for (int b = 0; b < 2 * D; ++b)
      if ((fsrc_faces >> b) & 1) stored = stored || c[b / 2] == ((b & 1) ? n[b / 2] - 1 : 0);
  }
  if (stored) {
    // with external face BCs the forces of the step are the ones FlowCalcRhoForces formed BEFORE
    // BCApply / BCUpdateRho changed the face nodes (lbm_flow.F90:1958-1991): take the stored ones
#pragma unroll
    for (int m = 0; m < S; ++m)
#pragma unroll
      for (int d = 0; d < D; ++d) F[m][d] = __ldg(Fsrc + (long long)(m * D + d) * g.fs + nd.pos);
  } else {
    forces<L, S, ISO>(g, p, rho, cls, ffmask, nd, mask, r, F);
  }
  common_velocity<L, S>(p, f, r, F, up);
#pragma unroll
  for (int m = 0; m < S; ++m) {
    if (rho_out) rho_out[m * g.nnodes + nd.o] = r[m];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (u_out) u_out[(long long)(m * D + d) * g.nnodes + nd.o] = up[d];
      if (F_out) F_out[(long long)(m * D + d) * g.nnodes + nd.o] = F[m][d];
    }
  }
  if (rhot || prs || velt) {
    double rt = 0.;
#pragma unroll
    for (int m = 0; m < S; ++m) rt += r[m] * p.mm[m];
    if (rhot) rhot[nd.o] = rt;
    if (prs) {
      double pr = rt / 3.;
      if (p.eos || S > 1) {
        double ps[S];
#pragma unroll
        for (int m = 0; m < S; ++m) ps[m] = p.eos ? eos_psi(p, m, r[m]) : r[m];
#pragma unroll
        for (int m = 0; m < S; ++m) {
          double acc = 0.;
#pragma unroll
          for (int mp = 0; mp < S; ++mp) acc += p.gf[m][mp] * ps[mp];
          pr = pr + 6.0 / 2. * ps[m] * acc;
        }
      }
      prs[nd.o] = pr;
    }
    if (velt) {
      static_for<0, D>([&](auto d_) {
        constexpr int d = decltype(d_)::value;
        double a = 0.;
#pragma unroll
        for (int m = 0; m < S; ++m) {
          double j = 0.;
          static_for<0, Q>([&](auto n_) {
            constexpr int n = decltype(n_)::value;
            if constexpr (L::c(n, d) != 0) j += f[m][n] * (double)L::c(n, d);
          });
          a += (j + .5 * F[m][d]) * p.mm[m];
        }
        velt[(long long)nd.o * D + d] = a / rt;
      });
    }
  }
}

// ------------------------------------------------------------------ set-up kernels

// nbmask and the fluid-fluid activity bits, evaluated once per walls upload.
// Line-of-sight rule: SURVEY.md 8a-15 / ff_stencil.cuh.
template <class L, int ISO>
__global__ void k_build_masks(Grid g, const uint8_t *__restrict__ cls, uint32_t *__restrict__ nbmask,
                              uint32_t *__restrict__ ffmask, int *__restrict__ specular) {
  NodeIdx nd;
  if (!node_of_thread(g, 0, g.NZl, nd)) return;
  const long long cbase = ((long long)(nd.z + g.Rz) * g.cny + (nd.y + g.R)) * g.cnx + (nd.x + g.R);
  auto solid = [&](int dx, int dy, int dz) -> bool {
    return cls[cbase + ((long long)dz * g.cny + dy) * g.cnx + dx] != 0;
  };
  uint32_t mask = solid(0, 0, 0) ? 0x80000000u : 0u;
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    const uint8_t c = cls[cbase + ((long long)L::c(n, 2) * g.cny + L::c(n, 1)) * g.cnx + L::c(n, 0)];
    if (c != 0) mask |= 1u << n;
    if (c >= 250 && c <= 252 && !(mask >> 31)) atomicAdd(specular, 1);
  });
  bool wallrec = (mask & MASK_DIRS) != 0;
  if constexpr (ISO != 4) {
    using FF = typename L::FF;
    constexpr int E = ff_entries<L>(ISO);
    uint32_t words[(E + 31) / 32];
#pragma unroll
    for (int w = 0; w < (E + 31) / 32; ++w) words[w] = 0u;
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      constexpr int ox = FF::off[e][0], oy = FF::off[e][1], oz = FF::off[e][2];
      bool ok = !solid(ox, oy, oz);
      bool any = false;
      static_for<0, FF::MAXALT>([&](auto a_) {
        constexpr int a = decltype(a_)::value;
        if constexpr (a < FF::nalt[e]) {
          bool all = true;
          static_for<0, FF::MAXLEN>([&](auto j_) {
            constexpr int j = decltype(j_)::value;
            if constexpr (j < FF::altlen[e][a]) {
              constexpr int lx = FF::los[e][a][j][0], ly = FF::los[e][a][j][1], lz = FF::los[e][a][j][2];
              all = all && !solid(lx, ly, lz);
            }
          });
          any = any || all;
        }
      });
      if (ok && any)
        words[e / 32] |= 1u << (e % 32);
      else
        wallrec = true;
    });
#pragma unroll
    for (int w = 0; w < (E + 31) / 32; ++w) ffmask[(long long)w * g.nnodes + nd.o] = words[w];
  }
  if (wallrec && !(mask >> 31)) mask |= MASK_WALLREC;
  if (g.perx && nd.x == 0) mask |= MASK_XLO;
  if (g.perx && nd.x == g.NX - 1) mask |= MASK_XHI;
  nbmask[nd.o] = mask;
}

}  // namespace txg

// hot_kernels.cuh -- the two per-timestep kernels of the flow hot path (sm_100a).
//
// Work decomposition: one lane per (fluid node, component).  A warp carries NPW = 32/S consecutive
// entries of the fluid-node list for all S components: lanes [m*NPW, (m+1)*NPW) hold component m, so
// every population load/store of a half-warp (S = 2) is one contiguous 128-byte run.  Quantities
// that couple the components (Shan-Chen gradient of the other phases, common velocity) cross lanes
// with warp shuffles, summed in ascending component order like the reference.  Splitting by
// component halves the live register set (19 populations instead of 38) -- the fp64 collision is
// otherwise register- and latency-bound on B200 long before HBM is.
//
// Only fluid nodes occupy lanes: `list` is the ascending list of fluid node indices of the slab
// (built once per walls upload), so a porous medium does not waste fp64 issue slots on solid
// voxels.  list == nullptr means "no solid node anywhere": entry i is node first + i.
#pragma once
#include "kernels.cuh"

namespace txg {

template <int S>
struct Lanes {
  static constexpr int NPW = 32 / S;  // nodes per warp
};

// (node, component) of this lane.  Returns false if the whole warp is beyond the range.  Lanes past
// the end of the range (or the 32 - S*NPW spare lanes when S does not divide 32) replay a valid
// item with active = false: they take part in the shuffles and store nothing.
template <int S>
__device__ __forceinline__ bool item_of_lane(const Grid &g, const uint32_t *__restrict__ list, long long first,
                                             long long count, NodeIdx &nd, int &m, int &j, bool &active) {
  constexpr int NPW = Lanes<S>::NPW;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long base = warp * NPW;
  if (base >= count) return false;
  m = lane / NPW;
  j = lane - m * NPW;
  active = true;
  if (m >= S) {
    m = S - 1;
    active = false;
  }
  long long i = base + j;
  if (i >= count) {
    i = count - 1;
    active = false;
  }
  const unsigned o = list ? __ldg(list + first + i) : (unsigned)(first + i);
  const unsigned plane = (unsigned)g.plane;
  const unsigned z = o / plane;
  const unsigned r = o - z * plane;
  const unsigned y = r / (unsigned)g.NX;
  nd.x = (int)(r - y * (unsigned)g.NX);
  nd.y = (int)y;
  nd.z = (int)z;
  nd.o = (long long)o;
  return true;
}

// value of `v` held by the lane of component k for the same node
template <int S>
__device__ __forceinline__ double from_component(double v, int k, int j) {
  if constexpr (S == 1) return v;
  return __shfl_sync(0xffffffffu, v, k * Lanes<S>::NPW + j);
}

// pull streaming with bounce-back for ONE component (see pull<> in kernels.cuh)
template <class L>
__device__ __forceinline__ void pull1(const Grid &g, const double *__restrict__ fAm /* fA + m*Q*fstride */,
                                      const NodeIdx &nd, uint32_t mask, double (&f)[L::Q]) {
  const int xm = wrapc(nd.x - 1, g.NX, g.perx), xp = wrapc(nd.x + 1, g.NX, g.perx);
  const int ym = wrapc(nd.y - 1, g.NY, g.pery), yp = wrapc(nd.y + 1, g.NY, g.pery);
  const long long here = (long long)(nd.z + 1) * g.plane + (long long)nd.y * g.NX + nd.x;
  static_for<0, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    constexpr int on = opp<L>(n);
    const int sx = L::c(n, 0) == 0 ? nd.x : (L::c(n, 0) > 0 ? xm : xp);
    const int sy = L::c(n, 1) == 0 ? nd.y : (L::c(n, 1) > 0 ? ym : yp);
    const int sz = nd.z + 1 - L::c(n, 2);
    const bool bounce = n != 0 && ((mask >> on) & 1u);
    const long long src = bounce ? here + (long long)on * g.fstride
                                 : ((long long)sz * g.plane + (long long)sy * g.NX + sx) + (long long)n * g.fstride;
    f[n] = __ldg(fAm + src);
  });
}

// FlowCalcForces for ONE component (see forces<> in kernels.cuh for the reference citations).
// rho_m / psi_m: this lane's component at this node; psi_field = psi + m*rstride.
template <class L, int S, int ISO>
__device__ __forceinline__ void forces1(const Grid &g, const Phys &p, const double *__restrict__ psi_field,
                                        const uint8_t *__restrict__ cls, const uint32_t *__restrict__ ffmask,
                                        const NodeIdx &nd, uint32_t mask, int m, int j, double rho_m, double psi_m,
                                        double (&F)[L::D]) {
  constexpr int D = L::D;
#pragma unroll
  for (int d = 0; d < D; ++d) F[d] = 0.;

  if (p.fluidsolid && (mask & 0x7fffffffu)) {
    const long long cbase = ((long long)(nd.z + g.Rz) * g.cny + (nd.y + g.R)) * g.cnx + (nd.x + g.R);
    static_for<1, L::Q>([&](auto n_) {
      constexpr int n = decltype(n_)::value;
      if ((mask >> n) & 1u) {
        const long long coff = ((long long)L::c(n, 2) * g.cny + L::c(n, 1)) * g.cnx + L::c(n, 0);
        const int id = cls[cbase + coff];
        if (id >= 1 && id <= p.nminerals) {
          constexpr double w = L::fs_weight(n);
          const double t = w * rho_m * __ldg(p.gw + (id - 1) * S + m);
          static_for<0, D>([&](auto d_) {
            constexpr int d = decltype(d_)::value;
            if constexpr (L::c(n, d) != 0) F[d] = F[d] - t * (double)L::c(n, d);
          });
        }
      }
    });
  }

  if (p.body) {
#pragma unroll
    for (int d = 0; d < D; ++d) F[d] = F[d] + p.gvt[d] * p.mm[m] * rho_m;
  }

  if (p.fluidfluid) {
    using FF = typename L::FF;
    constexpr int E = ff_entries<L>(ISO);
    constexpr int RAD = stencil_radius(ISO);
    double G[D], W[D];
#pragma unroll
    for (int d = 0; d < D; ++d) G[d] = W[d] = 0.;
    int xi[2 * RAD + 1], yi[2 * RAD + 1];
#pragma unroll
    for (int a = -RAD; a <= RAD; ++a) {
      xi[a + RAD] = wrapc(nd.x + a, g.NX, g.perx);
      yi[a + RAD] = wrapc(nd.y + a, g.NY, g.pery);
    }
    uint32_t words[(E + 31) / 32];
    if constexpr (ISO != 4) {
#pragma unroll
      for (int w = 0; w < (E + 31) / 32; ++w) words[w] = __ldg(ffmask + (long long)w * g.nnodes + nd.o);
    }
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      constexpr int dx = FF::off[e][0], dy = FF::off[e][1], dz = FF::off[e][2];
      bool on;
      if constexpr (ISO == 4) {
        constexpr int n = dir_of<L>(dx, dy, dz);
        on = !((mask >> n) & 1u);
      } else {
        on = (words[e / 32] >> (e % 32)) & 1u;
      }
      if (on) {
        constexpr double wgt = L::ffw(ISO, FF::L[e]);
        const long long nb = (long long)(nd.z + g.R + dz) * g.plane + (long long)yi[dy + RAD] * g.NX + xi[dx + RAD];
        const double diff = __ldg(psi_field + nb) - psi_m;
        if constexpr (dx != 0) {
          G[0] = G[0] + ((double)dx * wgt) * diff;
          W[0] = W[0] + wgt * (double)(dx * dx);
        }
        if constexpr (dy != 0) {
          G[1] = G[1] + ((double)dy * wgt) * diff;
          W[1] = W[1] + wgt * (double)(dy * dy);
        }
        if constexpr (D == 3 && dz != 0) {
          G[D - 1] = G[D - 1] + ((double)dz * wgt) * diff;
          W[D - 1] = W[D - 1] + wgt * (double)(dz * dz);
        }
      }
    });
    const double eps = (double)1.e-12f;  // default-real literal, lbm_forcing.F90:69
#pragma unroll
    for (int d = 0; d < D; ++d) {
      // gradient of this lane's component, normalised; every lane of the node sees the same W
      const double q = W[d] > eps ? G[d] / W[d] : 0.;
      double acc = 0.;
#pragma unroll
      for (int k = 0; k < S; ++k) acc += p.gf[m][k] * from_component<S>(q, k, j);
      if (W[d] > eps) F[d] = F[d] - 6.0 * psi_m * acc;  // c_0 = 6 on both lattices
    }
  }
}

// is moment row r even under n -> opp(n)?  (every row of both lattices is either even or odd)
template <class L>
TXG_HD constexpr bool row_even(int r) {
  for (int n = 0; n < L::Q; ++n)
    if (L::M(r, n) != L::M(r, opp<L>(n))) return false;
  return true;
}
template <class L>
TXG_HD constexpr bool row_odd(int r) {
  for (int n = 0; n < L::Q; ++n)
    if (L::M(r, n) != -L::M(r, opp<L>(n))) return false;
  return true;
}
template <class L>
TXG_HD constexpr bool rows_have_parity() {
  for (int r = 0; r < L::Q; ++r)
    if (!row_even<L>(r) && !row_odd<L>(r)) return false;
  return true;
}

// Equilibrium, forcing prefactor and SRT/MRT relaxation of ONE component, fused:
//   feq_n  (DiscretizationEquilf_*),  pref_n (FlowFiBarEqPrefactor, lbm_flow.F90:836-851),
//   f <- f - relax(f - (1 - pref/2) feq) + pref feq   (FlowCollisionD*, RelaxationCollide*).
// Opposite directions share (c.u)^2 and differ in the sign of c.u and c.F, so they are evaluated
// in pairs; the MRT transform is split into the even rows acting on pair sums and the odd rows on
// pair differences (M^-1 = M^T diag(1/|M_r|^2), lbm_relaxation.F90:182-200) -- the same arithmetic
// as the reference's 19 rank-1 updates up to summation order.
template <class L, bool MRT>
__device__ __forceinline__ void collide1(const Phys &p, int m, double rho, const double (&F)[L::D],
                                         const double (&u)[L::D], double (&f)[L::Q]) {
  constexpr int Q = L::Q, D = L::D;
  static_assert(rows_have_parity<L>(), "moment rows must be even or odd under direction reversal");
  double usqr = 0., Fu = 0.;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    usqr += u[d] * u[d];
    Fu += F[d] * u[d];
  }
  const double d_k = p.d_k[m];
  const double inv = 3.0 / rho;                       // 1 / (rho c_s2)
  const double base = 1.5 * (1. - d_k) - 1.5 * usqr;  // 1.5(1-d_k) - usqr/(2 c_s2)
  double dn[Q];                                       // f - fbar_eq
  {
    const double feq = rho * L::feq0(d_k, usqr);
    const double gg = (-Fu * inv) * feq;  // pref_0 feq_0
    dn[0] = (f[0] - feq) + .5 * gg;
    f[0] = f[0] + gg;
  }
  static_for<1, Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    constexpr int o = opp<L>(n);
    if constexpr (n < o) {
      double cu = 0., cF = 0.;
      static_for<0, D>([&](auto d_) {
        constexpr int d = decltype(d_)::value;
        if constexpr (L::c(n, d) != 0) {
          cu += (double)L::c(n, d) * u[d];
          cF += (double)L::c(n, d) * F[d];
        }
      });
      const double wr = L::w(n) * rho;
      const double t = base + 4.5 * (cu * cu);
      const double b = 3. * cu;
      const double feq_p = wr * (t + b), feq_m = wr * (t - b);
      const double g_p = ((cF - Fu) * inv) * feq_p, g_m = ((-cF - Fu) * inv) * feq_m;
      dn[n] = (f[n] - feq_p) + .5 * g_p;
      dn[o] = (f[o] - feq_m) + .5 * g_m;
      f[n] = f[n] + g_p;
      f[o] = f[o] + g_m;
    }
  });
  if constexpr (!MRT) {
    const double it = p.inv_tau[m];
#pragma unroll
    for (int n = 0; n < Q; ++n) f[n] = f[n] - dn[n] * it;
  } else {
    // pair sums / differences (in place: dn[n] <- sum, dn[opp] <- difference)
    static_for<1, Q>([&](auto n_) {
      constexpr int n = decltype(n_)::value;
      constexpr int o = opp<L>(n);
      if constexpr (n < o) {
        const double s = dn[n] + dn[o], a = dn[n] - dn[o];
        dn[n] = s;
        dn[o] = a;
      }
    });
    double cr[Q];
    static_for<0, Q>([&](auto r_) {
      constexpr int r = decltype(r_)::value;
      double mom = 0.;
      if constexpr (row_even<L>(r)) {
        if constexpr (L::M(r, 0) != 0) mom = (double)L::M(r, 0) * dn[0];
        static_for<1, Q>([&](auto n_) {
          constexpr int n = decltype(n_)::value;
          if constexpr (n < opp<L>(n) && L::M(r, n) != 0) mom += (double)L::M(r, n) * dn[n];
        });
      } else {
        static_for<1, Q>([&](auto n_) {
          constexpr int n = decltype(n_)::value;
          if constexpr (n < opp<L>(n) && L::M(r, n) != 0) mom += (double)L::M(r, n) * dn[opp<L>(n)];
        });
      }
      cr[r] = p.mrt_rate[m][r] * mom;
    });
    {
      double e0 = 0.;
      static_for<0, Q>([&](auto r_) {
        constexpr int r = decltype(r_)::value;
        if constexpr (row_even<L>(r) && L::M(r, 0) != 0) e0 += (double)L::M(r, 0) * cr[r];
      });
      f[0] = f[0] - e0;
    }
    static_for<1, Q>([&](auto n_) {
      constexpr int n = decltype(n_)::value;
      constexpr int o = opp<L>(n);
      if constexpr (n < o) {
        double ev = 0., od = 0.;
        static_for<0, Q>([&](auto r_) {
          constexpr int r = decltype(r_)::value;
          if constexpr (L::M(r, n) != 0) {
            if constexpr (row_even<L>(r))
              ev += (double)L::M(r, n) * cr[r];
            else
              od += (double)L::M(r, n) * cr[r];
          }
        });
        f[n] = f[n] - (ev + od);
        f[o] = f[o] - (ev - od);
      }
    });
  }
}

// ================================================================== the two hot kernels

// K1 moments: stream + bounce-back folded into the read, rho_m = sum_n f_n (ascending n); writes rho
// (psi with an EOS).  Replaces DistributionStreamD*, DistributionBouncebackD*,
// DistributionCalcDensityD* (lbm_distribution_function.F90:379-428,560-784) and EOSApply.
template <class L, int S>
__global__ void __launch_bounds__(128) k_moments(Grid g, Phys p, const double *__restrict__ fA,
                                                 double *__restrict__ rho, const uint32_t *__restrict__ nbmask,
                                                 const uint32_t *__restrict__ list, long long first, long long count) {
  NodeIdx nd;
  int m, j;
  bool active;
  if (!item_of_lane<S>(g, list, first, count, nd, m, j, active)) return;
  const uint32_t mask = __ldg(nbmask + nd.o);
  if (mask >> 31) return;  // only reachable through the dense (list == nullptr) path
  double f[L::Q];
  pull1<L>(g, fA + (long long)m * L::Q * g.fstride, nd, mask, f);
  double a = 0.;
#pragma unroll
  for (int n = 0; n < L::Q; ++n) a += f[n];
  if (!active) return;
  const long long o = (long long)m * g.rstride + (long long)(nd.z + g.R) * g.plane + (long long)nd.y * g.NX + nd.x;
  rho[o] = p.eos ? eos_psi(p, m, a) : a;
}

// K2 collide: pull again, forces from the rho stencil, momentum, common velocity, equilibrium,
// prefactor, SRT/MRT relaxation, forcing term; writes the post-collision populations.
// Replaces LBMAddFluidFluid/FluidSolid/BodyForcesD* (lbm_forcing.F90), DistributionCalcFluxD*
// (lbm_distribution_function.F90:451-508), FlowUpdateUED* (lbm_flow.F90:494-574),
// DiscretizationEquilf_*, FlowFiBarEqPrefactor, FlowCollisionD* (lbm_flow.F90:836-1029),
// RelaxationCollide* (lbm_relaxation.F90:171-200).
template <class L, int S, bool MRT, int ISO>
__global__ void __launch_bounds__(128, 3)
    k_collide(Grid g, Phys p, const double *__restrict__ fA, double *__restrict__ fB, const double *__restrict__ rho,
              const uint32_t *__restrict__ nbmask, const uint32_t *__restrict__ ffmask,
              const uint8_t *__restrict__ cls, const uint32_t *__restrict__ list, long long first, long long count) {
  NodeIdx nd;
  int m, j;
  bool active;
  if (!item_of_lane<S>(g, list, first, count, nd, m, j, active)) return;
  constexpr int Q = L::Q, D = L::D;
  const uint32_t mask = __ldg(nbmask + nd.o);
  const bool solid = mask >> 31;  // dense path only; the lane keeps running for the shuffles
  double f[Q];
  pull1<L>(g, fA + (long long)m * Q * g.fstride, nd, solid ? 0u : mask, f);
  double r = 0.;
#pragma unroll
  for (int n = 0; n < Q; ++n) r += f[n];
  if (solid) r = 1.;  // keeps the arithmetic finite; nothing is stored
  const long long ro = (long long)(nd.z + g.R) * g.plane + (long long)nd.y * g.NX + nd.x;
  const double psi_m = p.eos ? __ldg(rho + (long long)m * g.rstride + ro) : r;
  double F[D];
  forces1<L, S, ISO>(g, p, rho + (long long)m * g.rstride, cls, ffmask, nd, solid ? 0x7fffffffu : mask, m, j, r, psi_m,
                     F);
  // momentum j_m (DistributionCalcFluxD*) and the common velocity u' (FlowUpdateUED*)
  double up[D];
  {
    double num[D], den = 0.;
    const double mmot = p.mmot[m];
    double ue[D];
    static_for<0, D>([&](auto d_) {
      constexpr int d = decltype(d_)::value;
      double a = 0.;
      static_for<0, Q>([&](auto n_) {
        constexpr int n = decltype(n_)::value;
        if constexpr (L::c(n, d) != 0) a += f[n] * (double)L::c(n, d);
      });
      ue[d] = (a + .5 * F[d]) * mmot;
      num[d] = 0.;
    });
    const double rm = r * mmot;
#pragma unroll
    for (int k = 0; k < S; ++k) {
      den += from_component<S>(rm, k, j);
#pragma unroll
      for (int d = 0; d < D; ++d) num[d] += from_component<S>(ue[d], k, j);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) up[d] = num[d] / den;
  }
  collide1<L, MRT>(p, m, r, F, up, f);
  if (!active || solid) return;
  const long long o = (long long)(nd.z + 1) * g.plane + (long long)nd.y * g.NX + nd.x;
  double *out = fB + (long long)m * Q * g.fstride + o;
#pragma unroll
  for (int n = 0; n < Q; ++n) out[(long long)n * g.fstride] = f[n];
}

}  // namespace txg

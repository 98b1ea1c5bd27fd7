// hot_kernels.cuh -- the per-timestep kernels of the flow hot path (sm_100a).
//
// Work decomposition: one lane per (fluid node, component).  A warp carries NPW = 32/S consecutive
// entries of the fluid-node list for all S components: lanes [m*NPW, (m+1)*NPW) hold component m, so
// every population load of a half-warp (S = 2) is one contiguous 128-byte run when the nodes are
// x-neighbours.  Quantities that couple the components (Shan-Chen gradient of the other phases,
// common velocity) cross lanes with warp shuffles, summed in ascending component order like the
// reference.  Splitting by component halves the live register set (19 populations instead of 38):
// the fp64 collision is otherwise register- and latency-bound on B200 long before HBM is.
//
// Only fluid nodes exist in storage (kernels.cuh), so a porous medium wastes neither fp64 issue slots
// nor DRAM sectors on solid voxels.  A launch covers the positions [first, first + count).
#pragma once
#include "kernels.cuh"

#ifndef TXG_ABL
#define TXG_ABL 0  // ablation builds (tools/): 1 no stores, 2 no collision arithmetic, 3 no density gather, 4 node-aligned stores, 5 = 2 + 4
#endif

namespace txg {

template <int S>
struct Lanes {
  static constexpr int NPW = 32 / S;  // nodes per warp
};

struct Item {
  long long pos;  // position in the fluid list
  int m, j;       // component, node slot inside the warp
  bool active;    // false: replayed item, stores nothing
};

// (position, component) of this lane.  Returns false if the whole warp is beyond the range.  Lanes past
// the end of the range (or the 32 - S*NPW spare lanes when S does not divide 32) replay a valid
// item with active = false: they take part in the shuffles and store nothing.
template <int S>
__device__ __forceinline__ bool item_of_lane(long long first, long long count, long long warp, Item &it) {
  constexpr int NPW = Lanes<S>::NPW;
  const int lane = threadIdx.x & 31;
  const long long base = warp * NPW;
  if (base >= count) return false;
  it.m = lane / NPW;
  it.j = lane - it.m * NPW;
  it.active = true;
  if (it.m >= S) {
    it.m = S - 1;
    it.active = false;
  }
  long long i = base + it.j;
  if (i >= count) {
    i = count - 1;
    it.active = false;
  }
  it.pos = first + i;
  return true;
}
template <int S>
__device__ __forceinline__ bool item_of_lane(long long first, long long count, Item &it) {
  return item_of_lane<S>(first, count, ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, it);
}

// in-plane coordinates of extended node index oe
__device__ __forceinline__ void xy_of(const Grid &g, unsigned oe, int &x, int &y) {
  const unsigned plane = (unsigned)g.plane;
  const unsigned r = oe % plane;
  const unsigned yy = r / (unsigned)g.NX;
  x = (int)(r - yy * (unsigned)g.NX);
  y = (int)yy;
}

// value of `v` held by the lane of component k for the same node
template <int S>
__device__ __forceinline__ double from_component(double v, int k, int j) {
  if constexpr (S == 1) return v;
  return __shfl_sync(0xffffffffu, v, k * Lanes<S>::NPW + j);
}

// element offset from x to its periodic (or clamped) neighbour x + a
__device__ __forceinline__ int wrap_delta(int v, int a, int N, int per) { return wrapc(v + a, N, per) - v; }

// bulk value of the gradient normalisation W[d] = sum_e ffw(L_e) c_e,d^2 over all entries (reference order)
template <class L, int ISO>
TXG_HD constexpr double bulk_weight_sum(int d) {
  double w = 0.;
  for (int e = 0; e < L::FF::E; ++e)
    if (L::FF::gate[e] <= ISO) {
      const int c = L::FF::off[e][d];
      if (c != 0) w = w + L::ffw(ISO, L::FF::L[e]) * (double)(c * c);
    }
  return w;
}

// FlowCalcForces (lbm_flow.F90:760-808) for ONE component: fluid-solid, body, fluid-fluid, in the
// reference's order.  psi_m: pointer to this component's psi array (position-indexed).
// The geometry-only factors come from the wall record: A[d] = sum_n w_n gw(mineral(X+c_n), m) c_n,d
// (LBMAddFluidSolidForcesD*, lbm_forcing.F90:1326-1421, float literals 1./6. etc. folded in by
// k_build_wallrec) and rW[d] = 1/weightsum_d (lbm_forcing.F90:946-953); bulk nodes use the
// compile-time weight sum.  Neighbour densities are loaded unconditionally (a solid neighbour's
// position is that of the next fluid node -- some valid, finite value) and masked afterwards, so that
// all loads of a lane are in flight together.  npos[n]: position of X + c_n (order 4 re-uses them).
// Adjacency of a fluid node.  Positions run along x, so P(x+1,y',z') = P(x,y',z') + fluid(x,y',z') for any
// row (P of a solid node = position of the next fluid node): the table holds the position of X + c for
// the NCEN centre directions (c_x = 0) only, and the c_x = +-1 neighbours follow from those and the
// mask bits.  Nodes on a periodic x face (mask bits 28/29) look their wrapped neighbours up through
// the node -> position map instead.  For a solid neighbour the result is some valid position that
// the mask bit keeps from being used.
template <class L>
struct Adjacency {
  static constexpr int NCEN = num_centres<L>();
  unsigned cen[NCEN > 0 ? NCEN : 1];
  unsigned here;
  uint32_t mask;

  __device__ __forceinline__ void load(const Grid &g, const uint32_t *__restrict__ nbr, long long pos) {
    here = (unsigned)pos;
#pragma unroll
    for (int k = 0; k < NCEN; ++k) cen[k] = __ldg(nbr + (long long)k * g.fs + pos);
  }

  // the same row out of the shared-memory columns of this thread (k_collide)
  __device__ __forceinline__ void load_staged(const uint32_t (*col)[128], long long pos) {
    here = (unsigned)pos;
#pragma unroll
    for (int k = 0; k < NCEN; ++k) cen[k] = col[k][threadIdx.x];
    mask = col[NCEN][threadIdx.x];
  }

  // wrapped neighbour of a node on a periodic x face (rare path: two dependent loads).  Takes the
  // grid fields by value: a reference would force a local copy of the kernel parameter.
  template <int n>
  __device__ __forceinline__ static unsigned wrapped(const uint32_t *__restrict__ list, const uint32_t *__restrict__ P,
                                                  int NX, int NY, int pery, unsigned here) {
    const unsigned oe = list ? __ldg(list + here) : here;
    const unsigned plane = (unsigned)(NX * NY);
    const unsigned r = oe % plane;
    const int y = (int)(r / (unsigned)NX), x = (int)(r - (unsigned)y * (unsigned)NX);
    const int delta = wrap_delta(x, L::c(n, 0), NX, 1) + wrap_delta(y, L::c(n, 1), NY, pery) * NX +
                      L::c(n, 2) * (int)plane;
    const long long t = (long long)oe + delta;
    return P ? __ldg(P + t) : (unsigned)t;
  }

  // position of X + c_n
  template <int n>
  __device__ __forceinline__ unsigned at(const Grid &g) const {
    constexpr int cx = L::c(n, 0);
    if constexpr (n == 0) return here;
    if constexpr (cx == 0) return cen[centre_rank<L>(n)];
    constexpr int nc = dir_of<L>(0, L::c(n, 1), L::c(n, 2));  // centre of the row of X + c_n (0: own row)
    unsigned base = here, centre_solid = 0u;
    if constexpr (nc != 0) {
      base = cen[centre_rank<L>(nc)];
      centre_solid = (mask >> nc) & 1u;
    }
    unsigned v;
    if constexpr (cx > 0)
      v = base + 1u - centre_solid;
    else
      v = base - 1u + ((mask >> n) & 1u);
    if (mask & (cx > 0 ? MASK_XHI : MASK_XLO)) v = wrapped<n>(g.list, g.P, g.NX, g.NY, g.pery, here);
    return v;
  }
};

// Second-round operands of a node, loaded by the caller as soon as the adjacency row and the mask
// have arrived and before the populations are touched (so that one wait covers both rounds):
//   A[d], rW[d]: the wall record (0 / bulk weight for a node without one);  vn[n]: psi(X + c_n) for the
//   order-4 stencil, whose offsets are the lattice directions (wider stencils load inside forces1).
template <class L>
struct Round2 {
  double A[L::D], rW[L::D], vn[L::Q];
};

template <class L, int S, int ISO>
__device__ __forceinline__ void load_round2(const Grid &g, const Phys &p, const double *__restrict__ psi_field,
                                            const double *__restrict__ wallrec, const Item &it, uint32_t mask,
                                            const unsigned (&npos)[L::Q], Round2<L> &r2) {
  constexpr int D = L::D;
  const bool rec = (mask & MASK_WALLREC) != 0;
  if (p.fluidsolid) {
#pragma unroll
    for (int d = 0; d < D; ++d) r2.A[d] = rec ? __ldg(wallrec + (long long)(it.m * D + d) * g.fs + it.pos) : 0.;
  }
  if (p.fluidfluid) {
    static_for<0, D>([&](auto d_) {
      constexpr int d = decltype(d_)::value;
      constexpr double bulk = 1.0 / bulk_weight_sum<L, ISO>(d);
      r2.rW[d] = rec ? __ldg(wallrec + (long long)(S * D + d) * g.fs + it.pos) : bulk;
    });
    if constexpr (ISO == 4) {
#pragma unroll
      for (int n = 1; n < L::Q; ++n) {
#if TXG_ABL == 3
        r2.vn[n] = (double)(npos[n] & 3u);
#else
        r2.vn[n] = __ldg(psi_field + npos[n]);
#endif
      }
    }
  }
}

// WideLookup: how a stencil wider than the lattice (orders 8, 10) fetches psi(X + (dx, dy, dz)): the default walks the
// node -> position map (two dependent loads per entry); k_forces_tile reads a dense shared-memory tile instead.
struct WideThroughMap {
  static constexpr bool tile = false;
  template <int dx, int dy, int dz>
  __device__ __forceinline__ double get() const {
    return 0.;
  }
};

template <class L, int S, int ISO, class WideLookup = WideThroughMap>
__device__ __forceinline__ void forces1(const Grid &g, const Phys &p, const double *__restrict__ psi_field,
                                        const uint32_t *__restrict__ ffmask, const Round2<L> &r2,
                                        const Item &it, unsigned oe, int x, int y, uint32_t mask,
                                        double rho_m, double psi_m, double (&F)[L::D], const WideLookup wide = WideLookup()) {
  constexpr int D = L::D;
  const int m = it.m;
#pragma unroll
  for (int d = 0; d < D; ++d) F[d] = 0.;

  if (p.fluidsolid) {
#pragma unroll
    for (int d = 0; d < D; ++d) F[d] = F[d] - rho_m * r2.A[d];
  }

  if (p.body) {
#pragma unroll
    for (int d = 0; d < D; ++d) F[d] = F[d] + p.gvt[d] * p.mm[m] * rho_m;
  }

  if (p.fluidfluid) {
    using FF = typename L::FF;
    constexpr int E = ff_entries<L>(ISO);
    constexpr int RAD = stencil_radius(ISO);
    int dxo[2 * RAD + 1], dyo[2 * RAD + 1];
    uint32_t words[(E + 31) / 32];
    if constexpr (ISO != 4) {
      if constexpr (!WideLookup::tile) {
#pragma unroll
        for (int a = -RAD; a <= RAD; ++a) {
          dxo[a + RAD] = wrap_delta(x, a, g.NX, g.perx);
          dyo[a + RAD] = wrap_delta(y, a, g.NY, g.pery) * g.NX;
        }
      }
      const long long o = (long long)oe - (long long)g.Rz * g.plane;  // owned dense index
#pragma unroll
      for (int w = 0; w < (E + 31) / 32; ++w) words[w] = __ldg(ffmask + (long long)w * g.nnodes + o);
    }
    const int plane = (int)g.plane;
    double G[D];
#pragma unroll
    for (int d = 0; d < D; ++d) G[d] = 0.;
    static_for<0, E>([&](auto e_) {
      constexpr int e = decltype(e_)::value;
      constexpr int dx = FF::off[e][0], dy = FF::off[e][1], dz = FF::off[e][2];
      bool on;
      double v;
      if constexpr (ISO == 4) {
        constexpr int n = dir_of<L>(dx, dy, dz);
        on = !((mask >> n) & 1u);
        v = r2.vn[n];
      } else if constexpr (WideLookup::tile) {
        on = (words[e / 32] >> (e % 32)) & 1u;
        v = wide.template get<dx, dy, dz>();
      } else {
        on = (words[e / 32] >> (e % 32)) & 1u;
        v = __ldg(psi_field + pos_of(g, (long long)oe + (dz * plane + dyo[dy + RAD] + dxo[dx + RAD])));
      }
      constexpr double wgt = L::ffw(ISO, FF::L[e]);
      const double diff = on ? v - psi_m : 0.;
      if constexpr (dx != 0) G[0] = G[0] + ((double)dx * wgt) * diff;
      if constexpr (dy != 0) G[1] = G[1] + ((double)dy * wgt) * diff;
      if constexpr (D == 3 && dz != 0) G[D - 1] = G[D - 1] + ((double)dz * wgt) * diff;
    });
#pragma unroll
    for (int d = 0; d < D; ++d) {
      // normalised gradient of this lane's component; rW = 0 where the reference skips the direction
      const double q = G[d] * r2.rW[d];
      double acc = 0.;
#pragma unroll
      for (int k = 0; k < S; ++k) acc += p.gf[m][k] * from_component<S>(q, k, it.j);
      F[d] = F[d] - 6.0 * psi_m * acc;  // c_0 = 6 on both lattices
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

// momentum j_m = sum_n f_n c_n of this lane's component (DistributionCalcFluxD*,
// lbm_distribution_function.F90:451-508) and the common velocity u' shared by all components
// (FlowUpdateUED*, lbm_flow.F90:494-574); the component sums run in ascending component order
// over the lanes that hold the same node.
template <class L, int S>
__device__ __forceinline__ void common_velocity1(const Phys &p, const Item &it, const double (&f)[L::Q], double r,
                                                 const double (&F)[L::D], double (&up)[L::D]) {
  constexpr int Q = L::Q, D = L::D;
  double num[D], den = 0.;
  const double mmot = p.mmot[it.m];
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
    den += from_component<S>(rm, k, it.j);
#pragma unroll
    for (int d = 0; d < D; ++d) num[d] += from_component<S>(ue[d], k, it.j);
  }
  const double rden = 1. / den;
#pragma unroll
  for (int d = 0; d < D; ++d) up[d] = num[d] * rden;
}

// ================================================================== the hot kernels

// K1 moments: rho_m = sum_n f_n (ascending n) of the streamed populations; writes rho (psi with an
// EOS).  Replaces DistributionCalcDensityD* (lbm_distribution_function.F90:379-428) and EOSApply;
// streaming and bounce-back happened in the push of the previous collide.
// PAIR: the launch covers TWO runs of positions -- [first, first + split_at) and, `jump` positions further on, the rest
// of `count` (the bottom and top boundary planes of a z-slab in one launch).
template <class L, int S, bool PAIR = false>
__global__ void __launch_bounds__(128, 16) k_moments(Grid g, Phys p, const double *__restrict__ fA,
                                                 double *__restrict__ rho, double *__restrict__ rho_true,
                                                 long long first, long long count, long long split_at, long long jump) {
  Item it;
  if (!item_of_lane<S>(first, count, it)) return;
  if constexpr (PAIR) {
    if (it.pos - first >= split_at) it.pos += jump;
  }
  const double *src = fA + (long long)it.m * L::Q * g.fs + it.pos;
  double f[L::Q];
#pragma unroll
  for (int n = 0; n < L::Q; ++n) f[n] = __ldg(src + (long long)n * g.fs);
  double a = 0.;
#pragma unroll
  for (int n = 0; n < L::Q; ++n) a += f[n];
  if (!it.active) return;
  // with a non-ideal EOS the stencil field is psi(rho) and the density proper is kept beside it
  // (rho_true == rho otherwise)
  if (p.eos) {
    rho_true[(long long)it.m * g.fs + it.pos] = a;
    a = eos_psi(p, it.m, a);
  }
  rho[(long long)it.m * g.fs + it.pos] = a;
}

// asynchronous 4-byte copy global -> shared (LDGSTS): no register holds the value in flight
__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// K2a forces: F_m(X) of FlowCalcForces (lbm_flow.F90:760-808) for every fluid node, written to
// Fbuf[(m*D + d)*fs + pos].  All data-dependent gathers of the step live here -- the neighbour
// densities (18 for the order-4 stencil, up to 92 for order 8) and the wall record -- in a kernel that
// needs few registers and runs at full occupancy; inside the register-heavy collide kernel the same
// gathers cost 3.5 ms per step against 1.4 ms here (tools/membench/layoutbench.cu, 512^3 porous).
// Replaces LBMAddFluidSolidForcesD*, LBMAddBodyForcesD*, LBMAddFluidFluidForcesD* (lbm_forcing.F90).
// FACE: the launch covers the fluid nodes of ONE box face (entry i of the face array -> node, face_node) instead of a run
// of positions: with external face BCs on the fused step only the face nodes need their forces stored (BCApply reads them,
// and the face nodes are collided again with them, flow.cu one_step_bc).
template <class L, int S, int ISO, bool FACE = false>
__global__ void __launch_bounds__(128, 8) k_forces(Grid g, Phys p, const double *__restrict__ rho,
                                                const double *__restrict__ rho_true, const uint32_t *__restrict__ lmask,
                                                const uint32_t *__restrict__ nbr, const uint32_t *__restrict__ ffmask,
                                                const double *__restrict__ wallrec, double *__restrict__ Fbuf,
                                                long long first, long long count, FaceDesc fd, const uint32_t *__restrict__ nbmask) {
  constexpr int Q = L::Q, D = L::D;
  Item it;
  if (!item_of_lane<S>(first, count, it)) return;
  if constexpr (FACE) {  // (first = 0, count = entries of the face; a solid entry replays some fluid node and stores nothing)
    long long pos = g.own0;
    const bool fluid = face_node(g, fd, nbmask, it.pos, pos);
    it.active = it.active && fluid;
    it.pos = pos;
  }
  Adjacency<L> adj;
  adj.mask = __ldg(lmask + it.pos);
  adj.load(g, nbr, it.pos);
  const uint32_t mask = adj.mask;
  const double *psi_field = rho + (long long)it.m * g.fs;
  const double psi_m = __ldg(psi_field + it.pos);
  // density proper (differs from psi only with a non-ideal EOS)
  const double r = p.eos ? __ldg(rho_true + (long long)it.m * g.fs + it.pos) : psi_m;
  unsigned npos[Q];
  static_for<0, Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    npos[n] = adj.template at<n>(g);
  });
  unsigned oe = 0;
  int x = 0, y = 0;
  if constexpr (ISO != 4) {  // wider stencils look their extra neighbours up through P
    oe = g.list ? __ldg(g.list + it.pos) : (unsigned)it.pos;
    xy_of(g, oe, x, y);
  }
  Round2<L> r2;
  load_round2<L, S, ISO>(g, p, psi_field, wallrec, it, mask, npos, r2);
  double F[D];
  forces1<L, S, ISO>(g, p, psi_field, ffmask, r2, it, oe, x, y, mask, r, psi_m, F);
  if (!it.active) return;
#pragma unroll
  for (int d = 0; d < D; ++d) Fbuf[(long long)(it.m * D + d) * g.fs + it.pos] = F[d];
}

// K2a for the wide stencils (orders 8, 10), tile-staged: the 92 (D3Q19 order 8) gathers of a node through the node ->
// position map are two dependent loads each (k_forces: 13.8 ms per launch at 512^3, profiles/r1_bench_iso8_split_path.json).
// Here a block owns a dense column of TX x TY nodes and MARCHES along z over `zc` planes.  It keeps psi of the column plus a
// halo of RAD nodes in x and y for the 2 RAD + 1 planes around the current one -- every component -- in shared memory as a
// ring of dense planes, filling ONE new plane per step (one coalesced look-up of the position map per box node:
// (TX + 2 RAD)(TY + 2 RAD) / (TX TY) = 1.7 look-ups per tile node and plane instead of 92 per fluid node; the first form
// of this kernel re-filled the whole 5-plane box per plane: 8.4).  Its lanes then take the plane's FLUID nodes only (found
// through the row starts of the position map: no lane is spent on a solid node) and read every stencil entry from the
// ring at a compile-time offset.  Same arithmetic and summation order as k_forces: bit-identical forces.  Replaces the
// same reference procedures (LBMAddFluidFluidForcesD* with the 8th / 10th order stencils, lbm_forcing.F90:51-1299;
// LBMAddFluidSolidForcesD*, LBMAddBodyForcesD*).
template <class L, int ISO>
struct ForceTile {
  static constexpr int RAD = stencil_radius(ISO);
  static constexpr int TX = 32, TY = 8;
  static constexpr int BX = TX + 2 * RAD, BY = TY + 2 * RAD, BZ = L::D == 3 ? 2 * RAD + 1 : 1;
  static constexpr int PLANE = BX * BY;
  static constexpr int RING = L::D == 3 ? BZ + 1 : 1;  // one plane more than the stencil reads: the next plane lands meanwhile
  static constexpr int BOX = PLANE * RING;
  static constexpr int NT = 256;
  static constexpr int NIT = (PLANE + NT - 1) / NT;    // box nodes of one plane per thread
};

template <class L, int ISO>
struct WideFromTile {
  static constexpr bool tile = true;
  const double *pl[ForceTile<L, ISO>::BZ];  // psi of this lane's component at (x, y) in the ring planes z - RAD .. z + RAD
  template <int dx, int dy, int dz>
  __device__ __forceinline__ double get() const {
    using T = ForceTile<L, ISO>;
    return pl[L::D == 3 ? dz + T::RAD : 0][dy * T::BX + dx];
  }
};

// FUSE: the block does not write the forces; it goes on to collide and push its nodes (k_step_tile below) -- the collide
// kernel of the split path then neither re-reads the forces nor is launched.
template <class L, int S, int ISO, bool FUSE, bool MRT>
__device__ __forceinline__ void forces_tile_body(const Grid &g, const Phys &p, const double *__restrict__ rho,
                                                 const double *__restrict__ rho_true, const uint32_t *__restrict__ lmask,
                                                 const uint32_t *__restrict__ ffmask, const double *__restrict__ wallrec,
                                                 double *__restrict__ Fbuf, int z0, int nz, int zc, const double *__restrict__ fA,
                                                 double *__restrict__ fB, const uint32_t *__restrict__ nbr) {
  using T = ForceTile<L, ISO>;
  constexpr int D = L::D, RAD = T::RAD, NPW = Lanes<S>::NPW;
  extern __shared__ __align__(16) double box[];  // [S][RING][BY][BX]: ring slot of extended plane e = e % RING
  __shared__ unsigned row_pos[2][T::TY + 1];      // position of the first fluid node of every tile row (plane parity)
  __shared__ unsigned row_cnt[2][T::TY + 1];      // exclusive prefix of the fluid-node counts of the tile rows; [TY]: total
  const int x0 = blockIdx.x * T::TX, y0 = blockIdx.y * T::TY;
  const int zb = z0 + (int)blockIdx.z * zc, ze = min(z0 + nz, zb + zc);  // owned planes [zb, ze) of this block
  if (zb >= ze) return;
  const int nx = min(T::TX, g.NX - x0), ny = min(T::TY, g.NY - y0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // One dense plane of psi for the ring, in two halves so that the loads travel while the block computes: look-ups of the
  // position map + loads of psi into registers (wrapped or clamped coordinates: the entry masks exclude what lies beyond
  // a closed face), and later the stores into the ring slot.
  auto plane_load = [&](int e, double (&v)[S][T::NIT]) {
    unsigned bpos[T::NIT];
#pragma unroll
    for (int k = 0; k < T::NIT; ++k) {
      const int t = min((int)threadIdx.x + k * T::NT, T::PLANE - 1);
      const int bx = t % T::BX, by = t / T::BX;
      const int x = wrapc(x0 - RAD + bx, g.NX, g.perx), y = wrapc(y0 - RAD + by, g.NY, g.pery);
      bpos[k] = (unsigned)pos_of(g, ((long long)e * g.NY + y) * g.NX + x);
    }
#pragma unroll
    for (int mm = 0; mm < S; ++mm)
#pragma unroll
      for (int k = 0; k < T::NIT; ++k) v[mm][k] = __ldg(rho + (long long)mm * g.fs + bpos[k]);
  };
  auto plane_store = [&](int e, const double (&v)[S][T::NIT]) {
    const int slot = D == 3 ? e % T::RING : 0;
#pragma unroll
    for (int mm = 0; mm < S; ++mm)
#pragma unroll
      for (int k = 0; k < T::NIT; ++k) {
        const int t = (int)threadIdx.x + k * T::NT;
        if (t < T::PLANE) box[(mm * T::RING + slot) * T::PLANE + t] = v[mm][k];
      }
  };
  // rows of the tile in owned plane z: fluid nodes of row y are the positions [P(x0, y), P(x0 + nx, y)); exclusive prefix
  // of the counts by a shuffle scan inside warp 0 (no block-wide step)
  auto rows_of = [&](int z) {
    const int par = z & 1;
    unsigned first_pos = 0u, cnt = 0u;
    if (lane < ny) {
      const long long oe = ((long long)(z + g.Rz) * g.NY + (y0 + lane)) * g.NX + x0;
      first_pos = g.P ? __ldg(g.P + oe) : (unsigned)oe;
      cnt = (g.P ? __ldg(g.P + oe + nx) : (unsigned)(oe + nx)) - first_pos;
    }
    unsigned incl = cnt;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane <= T::TY) {
      row_pos[par][lane] = first_pos;
      row_cnt[par][lane] = incl - cnt;  // lanes >= ny hold cnt = 0: entry [TY] is the total
    }
  };
  {
    double v[S][T::NIT];
    if constexpr (D == 3) {
      for (int dz = -RAD; dz <= RAD; ++dz) {  // (Rz >= RAD: the ghost planes exist)
        plane_load(zb + g.Rz + dz, v);
        plane_store(zb + g.Rz + dz, v);
      }
    } else {
      plane_load(0, v);
      plane_store(0, v);
    }
  }
  if (warp == 0) rows_of(zb);
  __syncthreads();
  int m = lane / NPW;
  const int j = lane - m * NPW;
  bool lane_ok = true;
  if (m >= S) {
    m = S - 1;
    lane_ok = false;
  }
  for (int z = zb; z < ze; ++z) {
    const int zz = z + g.Rz;  // extended plane
    const int par = z & 1;
    // the next step's plane and row tables start travelling now; they land after this plane's arithmetic
    double vnext[S][T::NIT];
    const bool more = z + 1 < ze;
    if (more) {
      plane_load(zz + RAD + 1, vnext);
      if (warp == 0) rows_of(z + 1);
    }
    const unsigned nfluid = row_cnt[par][T::TY];
    for (unsigned c0 = (unsigned)warp * NPW; c0 < nfluid; c0 += (T::NT / 32) * NPW) {
      Item it;
      it.m = m;
      it.j = j;
      unsigned c = c0 + (unsigned)j;
      it.active = lane_ok && c < nfluid;
      c = min(c, nfluid - 1u);
      int ry = 0;
#pragma unroll
      for (int r = 1; r < T::TY; ++r) ry += (c >= row_cnt[par][r]) ? 1 : 0;
      it.pos = (long long)row_pos[par][ry] + (c - row_cnt[par][ry]);
      const unsigned oe = g.list ? __ldg(g.list + it.pos) : (unsigned)it.pos;
      const int x = (int)(oe % (unsigned)g.NX) - x0;  // column inside the tile
      const uint32_t mask = __ldg(lmask + it.pos);
      Round2<L> r2;
      const bool rec = (mask & MASK_WALLREC) != 0;
      if (p.fluidsolid) {
#pragma unroll
        for (int d = 0; d < D; ++d) r2.A[d] = rec ? __ldg(wallrec + (long long)(m * D + d) * g.fs + it.pos) : 0.;
      }
      if (p.fluidfluid) {
        static_for<0, D>([&](auto d_) {
          constexpr int d = decltype(d_)::value;
          constexpr double bulk = 1.0 / bulk_weight_sum<L, ISO>(d);
          r2.rW[d] = rec ? __ldg(wallrec + (long long)(S * D + d) * g.fs + it.pos) : bulk;
        });
      }
      WideFromTile<L, ISO> wide;
      const int inplane = (ry + RAD) * T::BX + (x + RAD);
#pragma unroll
      for (int k = 0; k < T::BZ; ++k) {
        const int slot = D == 3 ? (zz + k - RAD) % T::RING : 0;
        wide.pl[k] = box + (m * T::RING + slot) * T::PLANE + inplane;
      }
      const double psi_m = *wide.pl[D == 3 ? RAD : 0];
      const double r = p.eos ? __ldg(rho_true + (long long)m * g.fs + it.pos) : psi_m;
      double F[D];
      forces1<L, S, ISO>(g, p, rho + (long long)m * g.fs, ffmask, r2, it, oe, 0, 0, mask, r, psi_m, F, wide);
      if constexpr (!FUSE) {
        if (it.active) {
#pragma unroll
          for (int d = 0; d < D; ++d) Fbuf[(long long)(m * D + d) * g.fs + it.pos] = F[d];
        }
      } else {
        // K2b on the same lane: populations, common velocity, collision, push (k_collide, same arithmetic and order)
        constexpr int Q = L::Q;
        Adjacency<L> adj;
        adj.mask = mask;
        adj.load(g, nbr, it.pos);
        double f[Q];
        {
          const double *src = fA + (long long)m * Q * g.fs + it.pos;
#pragma unroll
          for (int n = 0; n < Q; ++n) f[n] = __ldg(src + (long long)n * g.fs);
        }
        double rr = 0.;
#pragma unroll
        for (int n = 0; n < Q; ++n) rr += f[n];
        double up[D];
        common_velocity1<L, S>(p, it, f, rr, F, up);
        collide1<L, MRT>(p, m, rr, F, up, f);
        if (it.active) {
          double *out = fB + (long long)m * Q * g.fs;
          const unsigned fs = (unsigned)g.fs, here = (unsigned)it.pos;
          out[here] = f[0];
          static_for<1, Q>([&](auto n_) {
            constexpr int n = decltype(n_)::value;
            constexpr int on = opp<L>(n);
            const bool bounce = (mask >> n) & 1u;
            const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + adj.template at<n>(g);
            out[e] = f[n];
          });
        }
      }
    }
    // the next plane goes into the one ring slot this step did not read (RING = stencil planes + 1)
    if (more) plane_store(zz + RAD + 1, vnext);
    __syncthreads();  // ONE barrier per plane: ring and row tables of the next step are complete, this step's are free
  }
}

template <class L, int S, int ISO>
__global__ void __launch_bounds__(ForceTile<L, ISO>::NT, 3) k_forces_tile(Grid g, Phys p, const double *__restrict__ rho,
                                                                     const double *__restrict__ rho_true,
                                                                     const uint32_t *__restrict__ lmask,
                                                                     const uint32_t *__restrict__ ffmask,
                                                                     const double *__restrict__ wallrec, double *__restrict__ Fbuf,
                                                                     int z0, int nz, int zc) {
  forces_tile_body<L, S, ISO, false, false>(g, p, rho, rho_true, lmask, ffmask, wallrec, Fbuf, z0, nz, zc, nullptr, nullptr, nullptr);
}

// K2 of the wide stencils in ONE kernel: forces out of the dense psi tile, then collision and push of the same nodes
// (k_forces_tile + k_collide without the force buffer in between).  Not used with external face BCs, whose BCApply
// sits between the forces and the collision.
template <class L, int S, bool MRT, int ISO>
__global__ void __launch_bounds__(ForceTile<L, ISO>::NT, 2) k_step_tile(Grid g, Phys p, const double *__restrict__ fA,
                                                                      double *__restrict__ fB, const double *__restrict__ rho,
                                                                      const double *__restrict__ rho_true,
                                                                      const uint32_t *__restrict__ lmask,
                                                                      const uint32_t *__restrict__ nbr,
                                                                      const uint32_t *__restrict__ ffmask,
                                                                      const double *__restrict__ wallrec, int z0, int nz, int zc) {
  forces_tile_body<L, S, ISO, true, MRT>(g, p, rho, rho_true, lmask, ffmask, wallrec, nullptr, z0, nz, zc, fA, fB, nbr);
}

// K2b collide + push: node populations and forces in, momentum, common velocity, equilibrium,
// prefactor, SRT/MRT relaxation, forcing term; the post-collision populations are streamed by the
// store (bounce-back folded in).  Every load address follows from the position alone.
// Replaces DistributionCalcFluxD* (lbm_distribution_function.F90:451-508), FlowUpdateUED*
// (lbm_flow.F90:494-574), DiscretizationEquilf_*, FlowFiBarEqPrefactor, FlowCollisionD*
// (lbm_flow.F90:836-1029), RelaxationCollide* (lbm_relaxation.F90:171-200), DistributionStreamD*,
// DistributionBouncebackD* (lbm_distribution_function.F90:560-784).
#ifndef TXG_COLLIDE_MIN_BLOCKS
#define TXG_COLLIDE_MIN_BLOCKS 4
#endif
template <class L, int S, bool MRT, bool FACE = false>
__global__ void __launch_bounds__(128, TXG_COLLIDE_MIN_BLOCKS)
    k_collide(Grid g, Phys p, const double *__restrict__ fA, double *__restrict__ fB, const double *__restrict__ Fbuf,
              const uint32_t *__restrict__ lmask, const uint32_t *__restrict__ nbr, long long first, long long count,
              const double *__restrict__ rho_stale /*[S][fs] density of FlowCalcRhoForces for MASK_STALE nodes, or null*/,
              FaceDesc fd, const uint32_t *__restrict__ nbmask /*FACE (k_forces has the convention)*/) {
  constexpr int Q = L::Q, D = L::D, NCEN = num_centres<L>();
  // The adjacency row and the mask are needed by the push only.  They travel into shared memory
  // asynchronously (no register is tied up while the collision runs, and their round trip hides
  // behind the population loads and the arithmetic); column threadIdx.x belongs to this lane alone.
  __shared__ uint32_t adj_col[NCEN + 1][128];
  Item it;
  if (!item_of_lane<S>(first, count, it)) return;
  if constexpr (FACE) {
    long long pos = g.own0;
    const bool fluid = face_node(g, fd, nbmask, it.pos, pos);
    it.active = it.active && fluid;
    it.pos = pos;
  }
#pragma unroll
  for (int k = 0; k < NCEN; ++k) cp_async4(&adj_col[k][threadIdx.x], nbr + (long long)k * g.fs + it.pos);
  cp_async4(&adj_col[NCEN][threadIdx.x], lmask + it.pos);
  cp_async_commit();
  double f[Q];
  {
    const double *src = fA + (long long)it.m * Q * g.fs + it.pos;
#pragma unroll
    for (int n = 0; n < Q; ++n) f[n] = __ldg(src + (long long)n * g.fs);
  }
  double F[D];
#pragma unroll
  for (int d = 0; d < D; ++d) F[d] = __ldg(Fbuf + (long long)(it.m * D + d) * g.fs + it.pos);
  double r = 0.;
#pragma unroll
  for (int n = 0; n < Q; ++n) r += f[n];
  if (rho_stale) {  // (handles with a BC_REFLECTING face only)
    if (__ldg(lmask + it.pos) & MASK_STALE) r = __ldg(rho_stale + (long long)it.m * g.fs + it.pos);
  }
  double up[D];
  common_velocity1<L, S>(p, it, f, r, F, up);
#if TXG_ABL != 2 && TXG_ABL != 5
  collide1<L, MRT>(p, it.m, r, F, up, f);
#else
  f[0] += up[0] + up[D - 1];
#endif
  if (!it.active) return;
#if TXG_ABL == 1
  {
    double acc = 0.;
#pragma unroll
    for (int n = 0; n < Q; ++n) acc += f[n];
    if (acc == 1.2345e300) fB[it.pos] = acc;
    return;
  }
#endif
  // push: slot (n, pos(X + c_n)), or slot (opp(n), pos(X)) when X + c_n is solid
  // (element indices inside one component's Q*fs block fit 32 bits: checked in txg_set_walls)
  cp_async_wait_all();
  Adjacency<L> adj;
  adj.load_staged(adj_col, it.pos);
  const uint32_t mask = adj.mask;
  double *out = fB + (long long)it.m * Q * g.fs;
  const unsigned fs = (unsigned)g.fs, here = (unsigned)it.pos;
  out[here] = f[0];
  static_for<1, Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    constexpr int on = opp<L>(n);
    const bool bounce = (mask >> n) & 1u;
#if TXG_ABL == 4 || TXG_ABL == 5
    const unsigned e = (unsigned)n * fs + here + ((bounce && adj.cen[0] == 0xffffffffu) ? 1u : 0u);
#else
    const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + adj.template at<n>(g);
#endif
    out[e] = f[n];
  });
}

// Adjacency table (one thread per owned position): nbr[k*fs + pos] = position of X + c_n for the k-th
// centre direction (c_x = 0), with the periodic wrap in y applied.  For a solid or out-of-domain
// neighbour the entry is the position of the next fluid node (see Adjacency).
template <class L>
__global__ void k_build_nbr(Grid g, uint32_t *__restrict__ nbr) {
  const long long pos = g.own0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= g.own1) return;
  const unsigned oe = g.list ? g.list[pos] : (unsigned)pos;
  int x, y;
  xy_of(g, oe, x, y);
  const int dym = wrap_delta(y, -1, g.NY, g.pery) * g.NX, dyp = wrap_delta(y, 1, g.NY, g.pery) * g.NX;
  const int plane = (int)g.plane;
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    constexpr int k = centre_rank<L>(n);
    if constexpr (k >= 0) {
      const int delta = (L::c(n, 1) == 0 ? 0 : (L::c(n, 1) > 0 ? dyp : dym)) + L::c(n, 2) * plane;
      nbr[(long long)k * g.fs + pos] = (uint32_t)pos_of(g, (long long)oe + delta);
    }
  });
}

// Wall records (one thread per owned position with MASK_WALLREC): the geometry-only factors of the
// fluid-solid force and of the gradient normalisation, evaluated once per walls upload.
//   A[m][d]  = sum over lattice directions n (ascending) with a mineral neighbour of
//              w_n * gw(mineral, m) * c_n,d, w_n the reference's default-real literals
//              (lbm_forcing.F90:1355,1364,1406,1414; 0 < walls < 998 and a valid mineral id)
//   rW[d]    = 1 / weightsum_d if weightsum_d > 1e-12 else 0   (lbm_forcing.F90:69,946-953)
template <class L, int S, int ISO>
__global__ void k_build_wallrec(Grid g, Phys p, const uint8_t *__restrict__ cls, const uint32_t *__restrict__ lmask,
                                const uint32_t *__restrict__ ffmask, double *__restrict__ rec) {
  const long long pos = g.own0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= g.own1) return;
  constexpr int D = L::D;
  const uint32_t mask = lmask[pos];
  if (!(mask & MASK_WALLREC)) return;
  const unsigned oe = g.list ? g.list[pos] : (unsigned)pos;
  int x, y;
  xy_of(g, oe, x, y);
  const int zz = (int)(oe / (unsigned)g.plane);  // z + Rz
  const long long o = (long long)oe - (long long)g.Rz * g.plane;
  const long long cbase = ((long long)zz * g.cny + (y + g.R)) * g.cnx + (x + g.R);
  double A[S][D];
#pragma unroll
  for (int m = 0; m < S; ++m)
#pragma unroll
    for (int d = 0; d < D; ++d) A[m][d] = 0.;
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    if ((mask >> n) & 1u) {
      const long long coff = ((long long)L::c(n, 2) * g.cny + L::c(n, 1)) * g.cnx + L::c(n, 0);
      const int id = cls[cbase + coff];
      if (id >= 1 && id <= p.nminerals) {
        constexpr double w = L::fs_weight(n);
#pragma unroll
        for (int m = 0; m < S; ++m)
          static_for<0, D>([&](auto d_) {
            constexpr int d = decltype(d_)::value;
            if constexpr (L::c(n, d) != 0) A[m][d] = A[m][d] + w * p.gw[(id - 1) * S + m] * (double)L::c(n, d);
          });
      }
    }
  });
  using FF = typename L::FF;
  constexpr int E = ff_entries<L>(ISO);
  double W[D];
#pragma unroll
  for (int d = 0; d < D; ++d) W[d] = 0.;
  static_for<0, E>([&](auto e_) {
    constexpr int e = decltype(e_)::value;
    constexpr int dx = FF::off[e][0], dy = FF::off[e][1], dz = FF::off[e][2];
    bool on;
    if constexpr (ISO == 4) {
      constexpr int n = dir_of<L>(dx, dy, dz);
      on = !((mask >> n) & 1u);
    } else {
      on = (ffmask[(long long)(e / 32) * g.nnodes + o] >> (e % 32)) & 1u;
    }
    if (on) {
      constexpr double wgt = L::ffw(ISO, FF::L[e]);
      if constexpr (dx != 0) W[0] = W[0] + wgt * (double)(dx * dx);
      if constexpr (dy != 0) W[1] = W[1] + wgt * (double)(dy * dy);
      if constexpr (D == 3 && dz != 0) W[D - 1] = W[D - 1] + wgt * (double)(dz * dz);
    }
  });
  const double eps = (double)1.e-12f;  // default-real literal, lbm_forcing.F90:69
#pragma unroll
  for (int m = 0; m < S; ++m)
#pragma unroll
    for (int d = 0; d < D; ++d) rec[(long long)(m * D + d) * g.fs + pos] = A[m][d];
#pragma unroll
  for (int d = 0; d < D; ++d) rec[(long long)(S * D + d) * g.fs + pos] = W[d] > eps ? 1. / W[d] : 0.;
}

// Halo unpack (the receiving half of DistributionCommunicateFi, lbm_distribution_function.F90:309-334,
// reduced to the populations that cross the face).  The pushes that left the neighbour slab through
// its ghost plane arrive in `src`; population n of boundary-plane node Y is taken iff its source
// Y - c_n is fluid -- otherwise slot (n, Y) already holds Y's own bounce-back value.  The fluid nodes
// of the sender's ghost plane are the fluid nodes of this boundary plane, in the same order.
//   up = 1: fills the bottom owned plane with the directions c_z = +1 (from the slab below)
//   up = 0: fills the top owned plane    with the directions c_z = -1 (from the slab above)
// [first, first+count): positions of the boundary plane.  Entry i of direction n, component m is
// src[(m*Q + n) * fs + src_first + i] when !packed (a ghost plane of an f buffer), else
// src[(m*NCROSS + k) * count + i], k the rank of n among the crossing directions (NCCL staging).
template <class L, int S>
__global__ void k_halo_unpack(Grid g, double *__restrict__ f, const double *__restrict__ src, long long src_first,
                              int packed, const uint32_t *__restrict__ lmask, long long first, long long count,
                              int up) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const long long pos = first + i;
  const uint32_t mask = lmask[pos];
  int k = 0;
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    if constexpr (L::c(n, 2) != 0) {
      if ((L::c(n, 2) > 0) == (up != 0)) {
        constexpr int on = opp<L>(n);
        if (!((mask >> on) & 1u)) {
#pragma unroll
          for (int m = 0; m < S; ++m) {
            const double v = packed ? src[(long long)(m * L::NCROSS + k) * count + i]
                                    : src[(long long)(m * L::Q + n) * g.fs + src_first + i];
            f[(long long)(m * L::Q + n) * g.fs + pos] = v;
          }
        }
        ++k;
      }
    }
  });
}

// The crossing populations of a ghost plane as ONE contiguous message: dst[(m*NCROSS + k) * count + i] =
// f[(m*Q + n) * fs + first + i], k the rank of n among the directions with c_z > 0 (up != 0) or < 0 -- the layout
// k_halo_unpack reads on the other side.  One send and one receive per neighbour instead of S * NCROSS of each.
template <class L, int S>
__global__ void k_halo_pack(Grid g, const double *__restrict__ f, double *__restrict__ dst, long long first, long long count, int up) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int k = 0;
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    if constexpr (L::c(n, 2) != 0) {
      if ((L::c(n, 2) > 0) == (up != 0)) {
#pragma unroll
        for (int m = 0; m < S; ++m) dst[(long long)(m * L::NCROSS + k) * count + i] = f[(long long)(m * L::Q + n) * g.fs + first + i];
        ++k;
      }
    }
  });
}

}  // namespace txg

// pipe_kernel.cuh -- software-pipelined collide + push kernel (isotropy order 4, the configuration
// of every shipped case and of the headline benchmark).
//
// Same arithmetic as k_collide (hot_kernels.cuh), different schedule.  The collision is a long fp64
// dependency chain that needs ~130 registers per lane, so only 12-16 warps fit on an SM -- far too few
// to hide the two dependent memory round trips of an item (populations + adjacency, then the
// neighbour densities the adjacency points at).  Here every warp is persistent and walks over its
// share of the fluid nodes; while it collides item k, the loads of item k+1 are already in flight as
// per-lane asynchronous copies (cp.async / LDGSTS) into the warp's private shared-memory stage:
//
//   top of iteration k     wait: everything of item k has landed; read it into registers
//                          issue A(k+1): populations, adjacency row, mask      -> stage
//   after the forces       wait A(k+1); issue B(k+1): neighbour densities (through the adjacency
//                          that just landed), wall record                      -> stage
//   rest of iteration k    velocity, equilibrium, relaxation, push stores
//
// Each lane stages and reads back only its own data, so no warp or block barrier is needed.
#pragma once
#include "hot_kernels.cuh"

namespace txg {

__device__ __forceinline__ void cp_async8(void *smem, const void *gptr) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gptr) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// one warp's stage; [..][32]: one column per lane, conflict-free
template <class L>
struct PipeStage {
  double f[L::Q][32];         // populations of the item being fetched / just fetched
  double psi[L::Q][32];       // [0]: own psi (EOS only), [n]: psi at X + c_n
  double rec[2 * L::D][32];   // wall record: A[d], rW[d]
  uint32_t nbr[2][L::Q][32];  // adjacency rows of item k and k+1 ([.][0] unused)
  uint32_t mask[2][32];
};

constexpr int PIPE_WARPS = 4;  // warps per block

template <class L, int S, bool MRT>
__global__ void __launch_bounds__(32 * PIPE_WARPS, 3)
    k_collide_pipe(Grid g, Phys p, const double *__restrict__ fA, double *__restrict__ fB,
                   const double *__restrict__ rho, const uint32_t *__restrict__ lmask,
                   const uint32_t *__restrict__ nbr, const double *__restrict__ wallrec, long long first,
                   long long count) {
  constexpr int Q = L::Q, D = L::D, NPW = Lanes<S>::NPW, ISO = 4;
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  PipeStage<L> &st = reinterpret_cast<PipeStage<L> *>(pipe_smem)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  int m = lane / NPW;
  const int j = lane - m * NPW;
  const bool lane_ok = m < S;
  if (!lane_ok) m = S - 1;
  const long long nitems = (count + NPW - 1) / NPW;
  const long long nwarps = (long long)gridDim.x * PIPE_WARPS;
  long long item = (long long)blockIdx.x * PIPE_WARPS + (threadIdx.x >> 5);
  if (item >= nitems) return;

  const double *fsrc = fA + (long long)m * Q * g.fs;
  const double *psi_field = rho + (long long)m * g.fs;
  double *out = fB + (long long)m * Q * g.fs;
  const unsigned fs = (unsigned)g.fs;

  auto position = [&](long long it_, bool &act) -> long long {
    long long i = it_ * NPW + j;
    act = lane_ok;
    if (i >= count) {
      i = count - 1;
      act = false;
    }
    return first + i;
  };
  // A: everything whose address follows from the position alone
  auto issue_a = [&](long long pos, int slot) {
#pragma unroll
    for (int n = 0; n < Q; ++n) cp_async8(&st.f[n][lane], fsrc + (long long)n * g.fs + pos);
#pragma unroll
    for (int n = 1; n < Q; ++n) cp_async4(&st.nbr[slot][n][lane], nbr + (long long)(n - 1) * g.fs + pos);
    cp_async4(&st.mask[slot][lane], lmask + pos);
    cp_async_commit();
  };
  // B: what the adjacency row and the mask point at
  auto issue_b = [&](long long pos, int slot) {
#pragma unroll
    for (int n = 1; n < Q; ++n) cp_async8(&st.psi[n][lane], psi_field + st.nbr[slot][n][lane]);
    if (p.eos) cp_async8(&st.psi[0][lane], psi_field + pos);
    if (st.mask[slot][lane] & MASK_WALLREC) {
      if (p.fluidsolid) {
#pragma unroll
        for (int d = 0; d < D; ++d) cp_async8(&st.rec[d][lane], wallrec + (long long)(m * D + d) * g.fs + pos);
      }
      if (p.fluidfluid) {
#pragma unroll
        for (int d = 0; d < D; ++d) cp_async8(&st.rec[D + d][lane], wallrec + (long long)(S * D + d) * g.fs + pos);
      }
    }
    cp_async_commit();
  };

  bool act;
  long long pos = position(item, act);
  int slot = 0;
  issue_a(pos, slot);
  cp_async_wait_all();
  issue_b(pos, slot);

  while (true) {
    cp_async_wait_all();
    double f[Q];
#pragma unroll
    for (int n = 0; n < Q; ++n) f[n] = st.f[n][lane];
    const uint32_t mask = st.mask[slot][lane];
    const bool rec = (mask & MASK_WALLREC) != 0;
    double r = 0.;
#pragma unroll
    for (int n = 0; n < Q; ++n) r += f[n];
    const double psi_m = p.eos ? st.psi[0][lane] : r;

    // ---- forces (FlowCalcForces, lbm_flow.F90:760-808; see forces1 in hot_kernels.cuh)
    double F[D];
#pragma unroll
    for (int d = 0; d < D; ++d) F[d] = 0.;
    if (p.fluidsolid) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double A = rec ? st.rec[d][lane] : 0.;
        F[d] = F[d] - r * A;
      }
    }
    if (p.body) {
#pragma unroll
      for (int d = 0; d < D; ++d) F[d] = F[d] + p.gvt[d] * p.mm[m] * r;
    }
    if (p.fluidfluid) {
      using FF = typename L::FF;
      constexpr int E = ff_entries<L>(ISO);
      double G[D];
#pragma unroll
      for (int d = 0; d < D; ++d) G[d] = 0.;
      static_for<0, E>([&](auto e_) {
        constexpr int e = decltype(e_)::value;
        constexpr int dx = FF::off[e][0], dy = FF::off[e][1], dz = FF::off[e][2];
        constexpr int n = dir_of<L>(dx, dy, dz);
        constexpr double wgt = L::ffw(ISO, FF::L[e]);
        const bool on = !((mask >> n) & 1u);
        const double diff = on ? st.psi[n][lane] - psi_m : 0.;
        if constexpr (dx != 0) G[0] = G[0] + ((double)dx * wgt) * diff;
        if constexpr (dy != 0) G[1] = G[1] + ((double)dy * wgt) * diff;
        if constexpr (D == 3 && dz != 0) G[D - 1] = G[D - 1] + ((double)dz * wgt) * diff;
      });
      static_for<0, D>([&](auto d_) {
        constexpr int d = decltype(d_)::value;
        constexpr double bulk = 1.0 / bulk_weight_sum<L, ISO>(d);
        const double rW = rec ? st.rec[D + d][lane] : bulk;
        const double q = G[d] * rW;
        double acc = 0.;
#pragma unroll
        for (int k = 0; k < S; ++k) acc += p.gf[m][k] * from_component<S>(q, k, j);
        F[d] = F[d] - 6.0 * psi_m * acc;
      });
    }

    // ---- the stage is consumed: start fetching the next item of this warp
    const long long next = item + nwarps;
    const bool has_next = next < nitems;  // warp-uniform
    bool act_next = false;
    long long pos_next = 0;
    if (has_next) {
      pos_next = position(next, act_next);
      issue_a(pos_next, slot ^ 1);
    }

    // ---- momentum j_m (DistributionCalcFluxD*) and the common velocity u' (FlowUpdateUED*)
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
      const double rden = 1. / den;
#pragma unroll
      for (int d = 0; d < D; ++d) up[d] = num[d] * rden;
    }

    if (has_next) {
      cp_async_wait_all();  // adjacency row and mask of the next item
      issue_b(pos_next, slot ^ 1);
    }

    collide1<L, MRT>(p, m, r, F, up, f);

    // ---- push: slot (n, pos(X + c_n)), or slot (opp(n), pos(X)) when X + c_n is solid
    if (act) {
      const unsigned here = (unsigned)pos;
      out[here] = f[0];
      static_for<1, Q>([&](auto n_) {
        constexpr int n = decltype(n_)::value;
        constexpr int on = opp<L>(n);
        const bool bounce = (mask >> n) & 1u;
        const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + st.nbr[slot][n][lane];
        out[e] = f[n];
      });
    }
    if (!has_next) break;
    item = next;
    pos = pos_next;
    act = act_next;
    slot ^= 1;
  }
}

}  // namespace txg

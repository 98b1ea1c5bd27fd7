// fused_kernel.cuh -- the one-kernel form of forces + collision + push (K2) for the order-4 stencil.
//
// With the order-4 Shan-Chen stencil the 18 gathered densities are the lattice neighbours the push needs
// anyway, and folding the forces into the collide kernel measures ~0.9 ms per step faster at 512^3 than
// the split k_forces + k_collide pair (11.4 ms against 3.1 + 9.3 ms; profiles/), although the gathers
// are individually more expensive inside the 126-register kernel.  Wider stencils (92 gathers at order
// 8) always take the split path.  Neighbour positions come from the full adjacency table nbr_all
// ((Q-1) entries per node); the split path's compact centre table is not used here.
#pragma once
#include "hot_kernels.cuh"

namespace txg {

// FlowCalcForces (lbm_flow.F90:760-808) for ONE component: fluid-solid, body, fluid-fluid, in the
// reference's order.  psi_m: pointer to this component's psi array (position-indexed).
// The geometry-only factors come from the wall record: A[d] = sum_n w_n gw(mineral(X+c_n), m) c_n,d
// (LBMAddFluidSolidForcesD*, lbm_forcing.F90:1326-1421, float literals 1./6. etc. folded in by
// k_build_wallrec) and rW[d] = 1/weightsum_d (lbm_forcing.F90:946-953); bulk nodes use the
// compile-time weight sum.  Neighbour densities are loaded unconditionally (a solid neighbour's
// position is that of the next fluid node -- some valid, finite value) and masked afterwards, so that
// all loads of a lane are in flight together.  npos[n]: position of X + c_n (order 4 re-uses them).
// Gather: how the order-4 stencil fetches psi of the lattice neighbour n at position np; the default loads it from global
// memory, k_step_fused_tile serves it from a shared-memory tile (TileGather below).
struct GlobalGather {
  template <int n>
  __device__ __forceinline__ double get(const double *__restrict__ psi_field, long long np) const {
    return __ldg(psi_field + np);
  }
};

template <class L, int S, int ISO, class Gather = GlobalGather>
__device__ __forceinline__ void forces1_inline(const Grid &g, const Phys &p, const double *__restrict__ psi_field,
                                        const uint32_t *__restrict__ ffmask, const double *__restrict__ wallrec,
                                        const Item &it, unsigned oe, int x, int y, uint32_t mask,
                                        const unsigned (&npos)[L::Q], double rho_m, double psi_m, double (&F)[L::D],
                                        const Gather gather = Gather()) {
  constexpr int D = L::D;
  const bool rec = (mask & MASK_WALLREC) != 0;
  const int m = it.m;
#pragma unroll
  for (int d = 0; d < D; ++d) F[d] = 0.;

  if (p.fluidsolid) {
    double A[D];
#pragma unroll
    for (int d = 0; d < D; ++d) A[d] = rec ? __ldg(wallrec + (long long)(m * D + d) * g.fs + it.pos) : 0.;
#pragma unroll
    for (int d = 0; d < D; ++d) F[d] = F[d] - rho_m * A[d];
  }

  if (p.body) {
#pragma unroll
    for (int d = 0; d < D; ++d) F[d] = F[d] + p.gvt[d] * p.mm[m] * rho_m;
  }

  if (p.fluidfluid) {
    using FF = typename L::FF;
    constexpr int E = ff_entries<L>(ISO);
    constexpr int RAD = stencil_radius(ISO);
    double rW[D];
    static_for<0, D>([&](auto d_) {
      constexpr int d = decltype(d_)::value;
      constexpr double bulk = 1.0 / bulk_weight_sum<L, ISO>(d);
      rW[d] = rec ? __ldg(wallrec + (long long)(S * D + d) * g.fs + it.pos) : bulk;
    });
    int dxo[2 * RAD + 1], dyo[2 * RAD + 1];
    uint32_t words[(E + 31) / 32];
    if constexpr (ISO != 4) {
#pragma unroll
      for (int a = -RAD; a <= RAD; ++a) {
        dxo[a + RAD] = wrap_delta(x, a, g.NX, g.perx);
        dyo[a + RAD] = wrap_delta(y, a, g.NY, g.pery) * g.NX;
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
      long long np;
      double v;
      if constexpr (ISO == 4) {
        constexpr int n = dir_of<L>(dx, dy, dz);
        on = !((mask >> n) & 1u);
        np = npos[n];
        v = gather.template get<n>(psi_field, np);
      } else {
        on = (words[e / 32] >> (e % 32)) & 1u;
        np = pos_of(g, (long long)oe + (dz * plane + dyo[dy + RAD] + dxo[dx + RAD]));
        v = __ldg(psi_field + np);
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
      const double q = G[d] * rW[d];
      double acc = 0.;
#pragma unroll
      for (int k = 0; k < S; ++k) acc += p.gf[m][k] * from_component<S>(q, k, it.j);
      F[d] = F[d] - 6.0 * psi_m * acc;  // c_0 = 6 on both lattices
    }
  }
}

// L2 prefetch of the rows a block of 128 lanes (4 warps x NPW positions) reads at the start of
// k_step_fused: S*Q population rows, Q-1 adjacency rows and the mask row of positions
// [first + blk*PB, first + (blk+1)*PB), PB = 4*NPW.  One 128-byte line per lane and round.
__device__ __forceinline__ void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// Cache-policy experiments on the population traffic of the fused kernels (build-time, default = plain):
//   TXG_LDF_MODE 1: loads do not allocate in L1 (the 19 rows are read once; L1 is left to the rho gathers)
//   TXG_STF_MODE 1: streaming stores (st.global.cs: evict-first), 2: st.global.cg
#ifndef TXG_LDF_MODE
#define TXG_LDF_MODE 0
#endif
#ifndef TXG_STF_MODE
#define TXG_STF_MODE 0
#endif
__device__ __forceinline__ double load_population(const double *p) {
#if TXG_LDF_MODE == 1
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ void store_population(double *p, double v) {
#if TXG_STF_MODE == 1
  __stcs(p, v);
#elif TXG_STF_MODE == 2
  __stcg(p, v);
#else
  *p = v;
#endif
}

template <class L, int S>
__device__ __forceinline__ void prefetch_block_rows(const Grid &g, const double *__restrict__ fA,
                                                    const uint32_t *__restrict__ lmask,
                                                    const uint32_t *__restrict__ nbr,
                                                    const double *__restrict__ wallrec, long long first,
                                                    long long count, long long blk) {
  constexpr int Q = L::Q, PB = 4 * Lanes<S>::NPW;
  constexpr int FL = (PB * 8 + 127) / 128, NL = (PB * 4 + 127) / 128;  // lines per row
  constexpr int NF = S * Q * FL, NN = (Q - 1) * NL, NW = (S * L::D + L::D) * FL;
  const int total = NF + NN + NL + (wallrec ? NW : 0);
  const long long p0 = blk * PB;
  if (p0 >= count) return;
  const long long pos = first + p0;
  for (int t = threadIdx.x; t < total; t += 128) {
    if (t < NF) {
      const int row = t / FL, seg = t - row * FL;
      prefetch_l2(fA + (long long)row * g.fs + pos + seg * 16);
    } else if (t < NF + NN) {
      const int u = t - NF, row = u / NL, seg = u - row * NL;
      prefetch_l2(nbr + (long long)row * g.fs + pos + seg * 32);
    } else if (t < NF + NN + NL) {
      prefetch_l2(lmask + pos + (t - NF - NN) * 32);
    } else {
      const int u = t - NF - NN - NL, row = u / FL, seg = u - row * FL;
      prefetch_l2(wallrec + (long long)row * g.fs + pos + seg * 16);
    }
  }
}

// K2 forces + collide + push: node populations, forces from the rho stencil, momentum, common velocity,
// equilibrium, prefactor, SRT/MRT relaxation, forcing term; the post-collision populations are
// streamed by the store (bounce-back folded in).
// Replaces LBMAddFluidFluid/FluidSolid/BodyForcesD* (lbm_forcing.F90), DistributionCalcFluxD*
// (lbm_distribution_function.F90:451-508), FlowUpdateUED* (lbm_flow.F90:494-574),
// DiscretizationEquilf_*, FlowFiBarEqPrefactor, FlowCollisionD* (lbm_flow.F90:836-1029),
// RelaxationCollide* (lbm_relaxation.F90:171-200), DistributionStreamD*, DistributionBouncebackD*
// (lbm_distribution_function.F90:560-784).
// (order-4 stencil only: its offsets are the lattice directions, so the gathers re-use npos)
// PAIR: two runs of positions in one launch (k_moments has the convention); no prefetch then.
#ifndef TXG_FUSED_THREADS
#define TXG_FUSED_THREADS 128
#endif
template <class L, int S, bool MRT, bool PAIR = false>
__global__ void __launch_bounds__(TXG_FUSED_THREADS, 512 / TXG_FUSED_THREADS)
    k_step_fused(Grid g, Phys p, const double *__restrict__ fA, double *__restrict__ fB, const double *__restrict__ rho,
                 const uint32_t *__restrict__ lmask, const uint32_t *__restrict__ nbr_all,
                 const double *__restrict__ wallrec, long long first, long long count, int pf_blocks, long long split_at,
                 long long jump) {
  constexpr int Q = L::Q, D = L::D, ISO = 4;
  // pf_blocks > 0: first ask L2 for the rows (populations, adjacency, mask, wall record) of the block
  // pf_blocks further on -- about one wave of resident blocks ahead -- so that its demand loads hit L2
  if (!PAIR && pf_blocks > 0) prefetch_block_rows<L, S>(g, fA, lmask, nbr_all, wallrec, first, count, (long long)blockIdx.x + pf_blocks);
  Item it;
  if (!item_of_lane<S>(first, count, it)) return;
  if constexpr (PAIR) {
    if (it.pos - first >= split_at) it.pos += jump;
  }
  // adjacency row and mask first: the second round of loads (neighbour densities, wall record)
  // hangs on them, the populations are not needed until the arithmetic starts
  const uint32_t mask = __ldg(lmask + it.pos);
  // positions of the lattice neighbours X + c_n (adjacency table, built once per walls upload)
  unsigned npos[Q];
  npos[0] = (unsigned)it.pos;
#pragma unroll
  for (int n = 1; n < Q; ++n) npos[n] = __ldg(nbr_all + (long long)(n - 1) * g.fs + it.pos);
  const long long mo = (long long)it.m * Q * g.fs + it.pos;
  double f[Q];
  {
    const double *src = fA + mo;
#pragma unroll
    for (int n = 0; n < Q; ++n) f[n] = load_population(src + (long long)n * g.fs);
  }
  const double *psi_field = rho + (long long)it.m * g.fs;
  double r = 0.;
#pragma unroll
  for (int n = 0; n < Q; ++n) r += f[n];
  const double psi_m = p.eos ? __ldg(psi_field + it.pos) : r;
  double F[D];
  forces1_inline<L, S, ISO>(g, p, psi_field, nullptr, wallrec, it, 0u, 0, 0, mask, npos, r, psi_m, F);
  double up[D];
  common_velocity1<L, S>(p, it, f, r, F, up);
  collide1<L, MRT>(p, it.m, r, F, up, f);
  if (!it.active) return;
  // push: slot (n, pos(X + c_n)), or slot (opp(n), pos(X)) when X + c_n is solid
  // (element indices inside one component's Q*fs block fit 32 bits: checked in txg_set_walls)
  double *out = fB + (long long)it.m * Q * g.fs;
  const unsigned fs = (unsigned)g.fs, here = (unsigned)it.pos;
  store_population(out + here, f[0]);
  static_for<1, Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    constexpr int on = opp<L>(n);
    const bool bounce = (mask >> n) & 1u;
    const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + npos[n];
    store_population(out + e, f[n]);
  });
}

// ------------------------------------------------------------------ rho tiles in shared memory (opt-in: TXG_RHOTILE=1)
// k_step_fused with the neighbour densities of the Shan-Chen stencil staged in shared memory by bulk copies (TMA,
// cp.async.bulk + mbarrier) issued at block start, so that they travel while the populations do instead of in a second,
// dependent round trip after the adjacency has arrived (57 % of K2's warp time is long-scoreboard, a quarter of it on the
// gathers: profiles/r1k_step_fused_ncu_summary.txt; issuing the gathers early costs registers the kernel does not have).
// Positions ascend in (z, y, x) and X -> X + (0, dy, dz) keeps that order, so the neighbours of a block of consecutive
// positions in one (dy, dz) row group are (nearly) one run of positions: rtab[blk][r] is the start of that run (set-up
// kernel k_build_rtab), CAP entries of it are copied per component, and a neighbour outside the window falls back to the
// global load (the position arrays are padded by 256 entries so that the window may overrun the last position).  The two same-row neighbours (x +- 1) stay global loads (adjacent lanes: L1 hits).
// NOT YET RUN ON A GPU (written in a session without GPU minutes).
template <class L>
TXG_HD constexpr int row_group(int n) {  // 0 .. NG-1 for the (dy, dz) != (0, 0) row groups, -1 for the node's own row
  const int cy = L::c(n, 1), cz = L::D == 3 ? L::c(n, 2) : 0;
  if (cy == 0 && cz == 0) return -1;
  const int k = L::D == 3 ? (cy + 1) * 3 + (cz + 1) : (cy + 1) * 3 + 1;  // 0..8 without 4
  return L::D == 3 ? (k < 4 ? k : k - 1) : (cy < 0 ? 0 : 1);
}
// Window length.  On the C4 geometry (192 x 192 x 48 sample, blocks of 64 positions) a window of 64 / 96 / 128 / 192 / 256
// positions covers 76.2 / 86.2 / 93.7 / 99.5 / 99.8 % of the fluid neighbours (the misses are blocks that straddle two x-rows).
#ifndef TXG_TILE_CAP
#define TXG_TILE_CAP 192
#endif
template <class L>
struct RhoTile {
  static constexpr int NG = L::D == 3 ? 8 : 2, CAP = TXG_TILE_CAP;
  static_assert(CAP % 2 == 0 && CAP >= 64 && CAP <= 256, "window of 16-byte granules inside the padding of the position arrays");
};

// every direction off the node's own x-row falls in exactly one of the NG row groups, and every group is used
template <class L>
TXG_HD constexpr bool row_groups_consistent() {
  int members[RhoTile<L>::NG] = {};
  for (int n = 0; n < L::Q; ++n) {
    const int r = row_group<L>(n);
    const bool own_row = L::c(n, 1) == 0 && (L::D == 2 || L::c(n, 2) == 0);
    if ((r < 0) != own_row || r >= RhoTile<L>::NG) return false;
    if (r >= 0) {
      ++members[r];
      for (int k = 0; k < n; ++k)  // same group <=> same (dy, dz)
        if (row_group<L>(k) >= 0 &&
            (row_group<L>(k) == r) != (L::c(k, 1) == L::c(n, 1) && (L::D == 2 || L::c(k, 2) == L::c(n, 2))))
          return false;
    }
  }
  for (int r = 0; r < RhoTile<L>::NG; ++r)
    if (members[r] == 0) return false;
  return true;
}
static_assert(row_groups_consistent<D3Q19>() && row_groups_consistent<D2Q9>(), "row groups of the density tiles");

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <class L, int S>
struct TileGather {
  const double *tile;      // [S][NG][CAP] in shared memory
  const uint32_t *starts;  // [NG] in shared memory
  int m;
  template <int n>
  __device__ __forceinline__ double get(const double *__restrict__ psi_field, long long np) const {
    constexpr int r = row_group<L>(n);
    if constexpr (r < 0) {
      return __ldg(psi_field + np);
    } else {
      const unsigned idx = (unsigned)np - starts[r];
      if (idx < (unsigned)RhoTile<L>::CAP) return tile[(m * RhoTile<L>::NG + r) * RhoTile<L>::CAP + idx];
      return __ldg(psi_field + np);
    }
  }
};

// warp 0 of a block: read the window starts of this block, then one bulk copy per (component, row group) into `tile`;
// completion is counted on `bar` (initialised by thread 0 + __syncthreads before the call)
template <class L, int S>
__device__ __forceinline__ void tile_issue(double *tile, uint32_t *starts, uint64_t *bar, const double *__restrict__ rho,
                                           long long fs, const uint32_t *__restrict__ rtab_block) {
  constexpr int NG = RhoTile<L>::NG, CAP = RhoTile<L>::CAP;
  const int lane = threadIdx.x;
  if (lane < NG) starts[lane] = __ldg(rtab_block + lane);
  __syncwarp();
  if (lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((unsigned)(S * NG * CAP * 8)) : "memory");
  __syncwarp();
  for (int idx = lane; idx < S * NG; idx += 32) {
    const int m = idx / NG, r = idx - m * NG;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(tile + (m * NG + r) * CAP)),
                 "l"(rho + (long long)m * fs + starts[r]), "r"((unsigned)(CAP * 8)), "r"(smem_addr(bar))
                 : "memory");
  }
}
// the tiles must have landed (acquire on the mbarrier also publishes `starts`); bounded, reports instead of hanging
__device__ __forceinline__ void tile_wait(uint64_t *bar, int *gave_up) {
  unsigned ok = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_addr(bar)), "r"(0u) : "memory");
    if (ok) break;
    ++spins;
    if (spins > (1u << 16) || ((spins & 255u) == 0 && *(volatile int *)gave_up != 0)) {  // or someone else gave up
      atomicAdd(gave_up, 1);
      break;
    }
  }
}

// rtab[blk * NG + r] = even-aligned start of the window of row group r for the block of PB positions blk (relative to
// own0): the smallest neighbour position of the block in that group (a solid neighbour's table entry is the position
// of the next fluid node, close by; the window is a heuristic, the fallback load keeps every value exact).
template <class L>
__global__ void k_build_rtab(Grid g, const uint32_t *__restrict__ nbr_all, int PB, long long nblocks, uint32_t *__restrict__ rtab) {
  constexpr int NG = RhoTile<L>::NG;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nblocks * NG) return;
  const long long blk = t / NG;
  const int r = (int)(t - blk * NG);
  const long long p0 = g.own0 + blk * PB, p1 = min(g.own1, p0 + PB);
  uint32_t lo = 0xffffffffu;
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    if (row_group<L>(n) == r)
      for (long long pos = p0; pos < p1; ++pos) lo = min(lo, __ldg(nbr_all + (long long)(n - 1) * g.fs + pos));
  });
  if (lo == 0xffffffffu) lo = 0;
  rtab[t] = lo & ~1u;
}

// the same for the blocks of the one-pass launch: block index = row * row_blocks + x, C block x of schedule row `row`
template <class L>
__global__ void k_build_rtab_lag(Grid g, const uint32_t *__restrict__ nbr_all, int PB, const uint32_t *__restrict__ crows /*[nrows][2]*/,
                                 long long nrows, int row_blocks, uint32_t *__restrict__ rtab) {
  constexpr int NG = RhoTile<L>::NG;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nrows * row_blocks * NG) return;
  const long long blk = t / NG;
  const int r = (int)(t - blk * NG);
  const long long row = blk / row_blocks, x = blk - row * row_blocks;
  const long long cfirst = crows[2 * row], ccount = crows[2 * row + 1];
  uint32_t lo = 0xffffffffu;
  if (x * PB < ccount) {
    const long long p0 = cfirst + x * PB, p1 = min(cfirst + ccount, p0 + PB);
    static_for<1, L::Q>([&](auto n_) {
      constexpr int n = decltype(n_)::value;
      if (row_group<L>(n) == r)
        for (long long pos = p0; pos < p1; ++pos) lo = min(lo, __ldg(nbr_all + (long long)(n - 1) * g.fs + pos));
    });
  }
  if (lo == 0xffffffffu) lo = 0;
  rtab[t] = lo & ~1u;
}

template <class L, int S, bool MRT>
__global__ void __launch_bounds__(128, 4)
    k_step_fused_tile(Grid g, Phys p, const double *__restrict__ fA, double *__restrict__ fB, const double *__restrict__ rho,
                      const uint32_t *__restrict__ lmask, const uint32_t *__restrict__ nbr_all,
                      const double *__restrict__ wallrec, const uint32_t *__restrict__ rtab, int *__restrict__ gave_up,
                      long long first, long long count, int pf_blocks) {
  constexpr int Q = L::Q, D = L::D, ISO = 4, NG = RhoTile<L>::NG, CAP = RhoTile<L>::CAP;
  __shared__ __align__(128) double tile[S * NG * CAP];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t starts[NG];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) tile_issue<L, S>(tile, starts, &bar, rho, g.fs, rtab + (long long)blockIdx.x * NG);
  if (pf_blocks > 0) prefetch_block_rows<L, S>(g, fA, lmask, nbr_all, wallrec, first, count, (long long)blockIdx.x + pf_blocks);
  Item it;
  const bool have = item_of_lane<S>(first, count, it);
  if (have) {
    const uint32_t mask = __ldg(lmask + it.pos);
    unsigned npos[Q];
    npos[0] = (unsigned)it.pos;
#pragma unroll
    for (int n = 1; n < Q; ++n) npos[n] = __ldg(nbr_all + (long long)(n - 1) * g.fs + it.pos);
    const long long mo = (long long)it.m * Q * g.fs + it.pos;
    double f[Q];
    {
      const double *src = fA + mo;
#pragma unroll
      for (int n = 0; n < Q; ++n) f[n] = load_population(src + (long long)n * g.fs);
    }
    const double *psi_field = rho + (long long)it.m * g.fs;
    double r = 0.;
#pragma unroll
    for (int n = 0; n < Q; ++n) r += f[n];
    const double psi_m = p.eos ? __ldg(psi_field + it.pos) : r;
    tile_wait(&bar, gave_up);
    double F[D];
    const TileGather<L, S> gather{tile, starts, it.m};
    forces1_inline<L, S, ISO>(g, p, psi_field, nullptr, wallrec, it, 0u, 0, 0, mask, npos, r, psi_m, F, gather);
    double up[D];
    common_velocity1<L, S>(p, it, f, r, F, up);
    collide1<L, MRT>(p, it.m, r, F, up, f);
    if (it.active) {
      double *out = fB + (long long)it.m * Q * g.fs;
      const unsigned fs = (unsigned)g.fs, here = (unsigned)it.pos;
      store_population(out + here, f[0]);
      static_for<1, Q>([&](auto n_) {
        constexpr int n = decltype(n_)::value;
        constexpr int on = opp<L>(n);
        const bool bounce = (mask >> n) & 1u;
        const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + npos[n];
        store_population(out + e, f[n]);
      });
    }
  }
  // a block must not retire while its bulk copies are in flight: warp 0 always has work (and waited above); the other
  // warps of a short last block leave early, which is harmless -- shared memory lives until the whole block is gone
}

// ------------------------------------------------------------------ one-pass step (opt-in: TXG_LAG=1)
// k_step_fused with the density sum of the NEXT step folded into the same launch (lag_schedule.h has the
// schedule and the dependency argument).  Block index = row * row_blocks + x: schedule row, block in the row.
//   x <  nC: C block -- exactly k_step_fused on PB positions of the row's C range, then fence + done[row] += 1;
//   x >= nC: M block -- wait until the <= 9 dependency rows are complete, then rho_next[m][pos] = sum_n fB[m][n][pos]
//            (ascending n, DistributionCalcDensityD*, lbm_distribution_function.F90:379-428) for MB positions of the
//            row's M ranges, read through L2 (ld.global.cg: other SMs wrote them in this launch).
// The C ranges of the schedule rows sit in constant memory so that range starts and counts stay in uniform registers
// like the kernel parameters of k_step_fused (a per-block table in global memory costs ~100 bytes of spills at the
// 128-register cap); the M ranges are read from global memory by the M blocks, which have registers to spare.  The
// constant table is per module and device: the host uploads it before every launch.
// Not used with a non-ideal EOS (psi needs its own pass), free-slip walls (the mirrors rewrite slots after
// the push) or external face BCs.  NOT YET RUN ON A GPU (written in a session without
// GPU minutes; the schedule is CPU-tested, tests/test_lag_schedule.py).
struct LagRowDev {
  uint32_t cfirst, ccount, m0first, m0count, m1first, m1count;  // lag_schedule.h LagRow (global memory: the M blocks read it)
};
struct LagCRow {
  uint32_t cfirst, ccount;  // the C part of a row, in constant memory
};
constexpr int LAG_MAX_ROWS = 7680;  // 60 KB of the 64 KB constant bank: 14 bands of a 512-plane slab
struct LagMeta {
  int rows_per_band, lag, MB;
  int nrows, row_blocks;  // the launch: nrows * row_blocks blocks
  int depbands[16][3];
};
__constant__ LagCRow c_lag_rows[LAG_MAX_ROWS];

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// L2 residency hints of the one-pass step (build-time, -DTXG_LAG_HINTS=1; default off): the streamed input is marked
// evict-first, the pushed populations evict-last until their density sum has read them (which demotes them again).
#ifndef TXG_LAG_HINTS
#define TXG_LAG_HINTS 0
#endif
__device__ __forceinline__ double lag_load_input(const double *p) {
#if TXG_LAG_HINTS
  unsigned long long pol;
  double v;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ void lag_store_pushed(double *p, double v) {
#if TXG_LAG_HINTS
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
#else
  *p = v;
#endif
}
__device__ __forceinline__ double lag_load_pushed(const double *p) {
#if TXG_LAG_HINTS
  unsigned long long pol;
  double v;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
#else
  return __ldcg(p);
#endif
}

// the M block of k_step_fused_lag
template <class L, int S>
__device__ __forceinline__ void lag_m_block(const Grid &g, const LagMeta &meta, const double *__restrict__ fB,
                                         double *__restrict__ rho_next, const LagRowDev *__restrict__ rows,
                                         unsigned *__restrict__ done, unsigned *__restrict__ gave_up, unsigned y, unsigned xm) {
  constexpr int Q = L::Q, NPW = Lanes<S>::NPW, PB = 4 * NPW;
  const LagRowDev row = rows[y];
  const unsigned MB = (unsigned)meta.MB;
  const unsigned n0 = (row.m0count + MB - 1) / MB;
  long long first, count;
  if (xm < n0) {
    first = (long long)row.m0first + (long long)xm * MB;
    count = min((long long)MB, (long long)row.m0count - (long long)xm * MB);
  } else {
    const unsigned q = xm - n0;
    if ((long long)q * MB >= (long long)row.m1count) return;  // padding block of the grid
    first = (long long)row.m1first + (long long)q * MB;
    count = min((long long)MB, (long long)row.m1count - (long long)q * MB);
  }
  // wait for the collisions that push into these positions: rows (b', zm + dz), b' in depbands[b]
  if (threadIdx.x < 9) {
    const int b = (int)y / meta.rows_per_band, k = (int)y - b * meta.rows_per_band;
    const int bb = meta.depbands[b][threadIdx.x / 3];
    if (bb >= 0) {
      const int dep = bb * meta.rows_per_band + (k - 1 - meta.lag) + ((int)threadIdx.x % 3 - 1);
      const unsigned want = (c_lag_rows[dep].ccount + PB - 1) / PB;
      // bounded: a schedule bug or out-of-order block dispatch must not hang the device (the host reports gave_up)
      unsigned spins = 0;
      while (ld_acquire_u32(done + dep) < want) {
        ++spins;
        if (spins > (1u << 20) || ((spins & 255u) == 0 && ld_acquire_u32(gave_up) != 0)) {  // ~1 s, or someone else gave up
          atomicAdd(gave_up, 1u);
          break;
        }
        __nanosleep(200);
      }
    }
  }
  __syncthreads();
  // two warps of positions per round: 2 x Q loads in flight per lane
  for (long long w = threadIdx.x >> 5; w * NPW < count; w += 8) {
    Item it0, it1;
    item_of_lane<S>(first, count, w, it0);  // (never beyond the range: the loop condition is its test)
    const bool two = item_of_lane<S>(first, count, w + 4, it1);
    if (!two) {
      it1 = it0;
      it1.active = false;
    }
    const double *src0 = fB + (long long)it0.m * Q * g.fs + it0.pos;
    const double *src1 = fB + (long long)it1.m * Q * g.fs + it1.pos;
    double v0[Q], v1[Q];
#pragma unroll
    for (int n = 0; n < Q; ++n) v0[n] = lag_load_pushed(src0 + (long long)n * g.fs);
#pragma unroll
    for (int n = 0; n < Q; ++n) v1[n] = lag_load_pushed(src1 + (long long)n * g.fs);
    double a0 = 0., a1 = 0.;
#pragma unroll
    for (int n = 0; n < Q; ++n) a0 += v0[n];
#pragma unroll
    for (int n = 0; n < Q; ++n) a1 += v1[n];
    if (it0.active) rho_next[(long long)it0.m * g.fs + it0.pos] = a0;
    if (it1.active) rho_next[(long long)it1.m * g.fs + it1.pos] = a1;
  }
}

// the C block of k_step_fused_lag: the body of k_step_fused on positions [first, first + count), warp `warp` of them
template <class L, int S, bool MRT, bool TILE>
__device__ __forceinline__ void lag_c_warp(const Grid &g, const Phys &p, const double *__restrict__ fA, double *__restrict__ fB,
                                           const double *__restrict__ rho, const uint32_t *__restrict__ lmask,
                                           const uint32_t *__restrict__ nbr_all, const double *__restrict__ wallrec,
                                           long long first, long long count, long long warp, const double *tile,
                                           const uint32_t *starts, uint64_t *bar, int *tile_gave_up) {
  constexpr int Q = L::Q, D = L::D, ISO = 4;
  Item it;
  if (!item_of_lane<S>(first, count, warp, it)) return;
  const uint32_t mask = __ldg(lmask + it.pos);
  unsigned npos[Q];
  npos[0] = (unsigned)it.pos;
#pragma unroll
  for (int n = 1; n < Q; ++n) npos[n] = __ldg(nbr_all + (long long)(n - 1) * g.fs + it.pos);
  const long long mo = (long long)it.m * Q * g.fs + it.pos;
  double f[Q];
  {
    const double *src = fA + mo;
#pragma unroll
    for (int n = 0; n < Q; ++n) f[n] = lag_load_input(src + (long long)n * g.fs);
  }
  const double *psi_field = rho + (long long)it.m * g.fs;
  double r = 0.;
#pragma unroll
  for (int n = 0; n < Q; ++n) r += f[n];
  double F[D];
  const double psi_m = p.eos ? __ldg(psi_field + it.pos) : r;  // (p.eos is 0 on this path; kept so that the code generated is k_step_fused's)
  if constexpr (TILE) {  // neighbour densities from the shared-memory windows (k_step_fused_tile)
    tile_wait(bar, tile_gave_up);
    const TileGather<L, S> gather{tile, starts, it.m};
    forces1_inline<L, S, ISO>(g, p, psi_field, nullptr, wallrec, it, 0u, 0, 0, mask, npos, r, psi_m, F, gather);
  } else {
    forces1_inline<L, S, ISO>(g, p, psi_field, nullptr, wallrec, it, 0u, 0, 0, mask, npos, r, psi_m, F);
  }
  double up[D];
  common_velocity1<L, S>(p, it, f, r, F, up);
  collide1<L, MRT>(p, it.m, r, F, up, f);
  if (!it.active) return;
  double *out = fB + (long long)it.m * Q * g.fs;
  const unsigned fs = (unsigned)g.fs, here = (unsigned)it.pos;
  lag_store_pushed(out + here, f[0]);
  static_for<1, Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    constexpr int on = opp<L>(n);
    const bool bounce = (mask >> n) & 1u;
    const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + npos[n];
    lag_store_pushed(out + e, f[n]);
  });
}

// TILE: the C blocks also stage the stencil's neighbour densities in shared memory (k_step_fused_tile); rtab then holds the
// window starts per block index of THIS launch (k_build_rtab_lag), tile_gave_up its give-up counter
template <class L, int S, bool MRT, bool TILE>
__global__ void __launch_bounds__(128, 4)
    k_step_fused_lag(Grid g, Phys p, LagMeta meta, const double *__restrict__ fA, double *__restrict__ fB,
                     const double *__restrict__ rho, double *__restrict__ rho_next, const uint32_t *__restrict__ lmask,
                     const uint32_t *__restrict__ nbr_all, const double *__restrict__ wallrec,
                     const LagRowDev *__restrict__ rows, unsigned *__restrict__ done, unsigned *__restrict__ gave_up,
                     const uint32_t *__restrict__ rtab, int *__restrict__ tile_gave_up, int pf_blocks) {
  constexpr int PB = 4 * Lanes<S>::NPW;
  constexpr int NG = RhoTile<L>::NG, CAP = RhoTile<L>::CAP;
  __shared__ __align__(128) double tile[TILE ? S * NG * CAP : 1];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t starts[NG];
  // 1-D launch of nrows * row_blocks blocks (a 1-D grid is dispatched in index order): row y, block x of the row
  const unsigned y = blockIdx.x / (unsigned)meta.row_blocks, x = blockIdx.x - y * (unsigned)meta.row_blocks;
  const LagCRow &row = c_lag_rows[y];
  const unsigned nC = (row.ccount + PB - 1) / PB;
  if (x >= nC) {
    lag_m_block<L, S>(g, meta, fB, rho_next, rows, done, gave_up, y, x - nC);
    return;
  }
  if constexpr (TILE) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) tile_issue<L, S>(tile, starts, &bar, rho, g.fs, rtab + (long long)blockIdx.x * NG);
  }
  // L2 prefetch for the C block pf_blocks further on in launch order: in this row, or at the start of the next one
  if (pf_blocks > 0) {
    const unsigned ahead = x + (unsigned)pf_blocks;
    const bool next = ahead >= nC && y + 1 < (unsigned)meta.nrows;
    const LagCRow &pr = c_lag_rows[y + (next ? 1u : 0u)];
    prefetch_block_rows<L, S>(g, fA, lmask, nbr_all, wallrec, (long long)pr.cfirst, (long long)pr.ccount, (long long)(next ? ahead - nC : ahead));
  }
  lag_c_warp<L, S, MRT, TILE>(g, p, fA, fB, rho, lmask, nbr_all, wallrec, (long long)row.cfirst, (long long)row.ccount,
                              ((long long)x * 128 + threadIdx.x) >> 5, tile, starts, &bar, tile_gave_up);
  // every thread's stores are ordered before the row count: release fence (acq_rel, lighter than __threadfence's
  // sequentially consistent one), block barrier, one release-add
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) red_release_add_u32(done + blockIdx.x / (unsigned)meta.row_blocks, 1u);  // (recomputed: no register held across the collision)
}

// FlowFiInit for the fused path (FlowFiInit lbm_flow.F90:923-934, FlowFeqBarD* :867-921), one lane per
// (fluid node, component) like the step kernel: F from rho0 through the same forces routine, then
// f = (1 - prefactor/2) feq(rho0, u0) into the node's own slots.  u0 is [S][D][nnodes] (dense) or null.
// (The generic k_fi_init walks the dense box and looks every neighbour up through P: 55 ms at 512^3
// against one step's 14 ms; this one streams.)
template <class L, int S>
__global__ void __launch_bounds__(128, 4)
    k_fi_init_fused(Grid g, Phys p, double *__restrict__ fN, const double *__restrict__ psi,
                    const double *__restrict__ rho_true, const double *__restrict__ u0,
                    const uint32_t *__restrict__ lmask, const uint32_t *__restrict__ nbr_all,
                    const double *__restrict__ wallrec, long long first, long long count) {
  constexpr int Q = L::Q, D = L::D;
  Item it;
  if (!item_of_lane<S>(first, count, it)) return;
  const uint32_t mask = __ldg(lmask + it.pos);
  unsigned npos[Q];
  npos[0] = (unsigned)it.pos;
#pragma unroll
  for (int n = 1; n < Q; ++n) npos[n] = __ldg(nbr_all + (long long)(n - 1) * g.fs + it.pos);
  const double *psi_field = psi + (long long)it.m * g.fs;
  const double r = __ldg(rho_true + (long long)it.m * g.fs + it.pos);
  const double psi_m = __ldg(psi_field + it.pos);
  double F[D];
  forces1_inline<L, S, 4>(g, p, psi_field, nullptr, wallrec, it, 0u, 0, 0, mask, npos, r, psi_m, F);
  double u[D];
#pragma unroll
  for (int d = 0; d < D; ++d) u[d] = 0.;
  if (u0) {
    const long long o = (g.list ? (long long)__ldg(g.list + it.pos) : it.pos) - (long long)g.Rz * g.plane;
#pragma unroll
    for (int d = 0; d < D; ++d) u[d] = __ldg(u0 + (long long)(it.m * D + d) * g.nnodes + o);
  }
  double feq[Q], pref[Q];
  equilibrium<L>(r, p.d_k[it.m], u, feq);
  prefactor<L>(r, F, u, pref);
  if (!it.active) return;
  double *dst = fN + (long long)it.m * Q * g.fs + it.pos;
#pragma unroll
  for (int n = 0; n < Q; ++n) dst[(long long)n * g.fs] = (1. - 0.5 * pref[n]) * feq[n];
}

// FlowUpdateDiagnosticsD* (lbm_flow.F90:654-758) for the fused path: rhot, prs and velt of the FLUID nodes, one lane per (fluid node,
// component) like the step kernel -- populations as position-aligned rows, forces through the adjacency table and the wall records, the
// component sums by shuffles in ascending component order.  The solid nodes of the dense output arrays are filled by
// k_export_fill_solid.  (k_export walks the dense box, holds both components of a node in one thread and looks every neighbour up
// through the position map and the class bytes: 27 ms at 512^3 against this kernel's few; it stays the path of the face-BC modes, of
// the wide stencils and of rho / u / forces exports.)  Same formulas in the same order as k_export.
template <class L, int S>
__global__ void __launch_bounds__(128, 4)
    k_export_diag_fused(Grid g, Phys p, const double *__restrict__ fA, const double *__restrict__ psi, const uint32_t *__restrict__ lmask,
                        const uint32_t *__restrict__ nbr_all, const double *__restrict__ wallrec, long long first, long long count,
                        double *__restrict__ rhot, double *__restrict__ prs, double *__restrict__ velt /*[nnodes][D]*/) {
  constexpr int Q = L::Q, D = L::D;
  Item it;
  if (!item_of_lane<S>(first, count, it)) return;
  const uint32_t mask = __ldg(lmask + it.pos);
  unsigned npos[Q];
  npos[0] = (unsigned)it.pos;
#pragma unroll
  for (int n = 1; n < Q; ++n) npos[n] = __ldg(nbr_all + (long long)(n - 1) * g.fs + it.pos);
  double f[Q];
  const double *src = fA + (long long)it.m * Q * g.fs + it.pos;
#pragma unroll
  for (int n = 0; n < Q; ++n) f[n] = load_population(src + (long long)n * g.fs);
  double r = 0.;
#pragma unroll
  for (int n = 0; n < Q; ++n) r += f[n];
  const double *psi_field = psi + (long long)it.m * g.fs;
  const double psi_m = p.eos ? __ldg(psi_field + it.pos) : r;
  double F[D];
  forces1_inline<L, S, 4>(g, p, psi_field, nullptr, wallrec, it, 0u, 0, 0, mask, npos, r, psi_m, F);
  // rhot = sum_m rho_m mm_m
  const double rmm = r * p.mm[it.m];
  double rt = 0.;
#pragma unroll
  for (int k = 0; k < S; ++k) rt += from_component<S>(rmm, k, it.j);
  // prs = rhot / 3 + (c_0 / 2) sum_m psi_m sum_m' g_mm' psi_m'
  double pr = rt / 3.;
  if (p.eos || S > 1) {
    const double ps_own = p.eos ? eos_psi(p, it.m, r) : r;
    double ps[S];
#pragma unroll
    for (int k = 0; k < S; ++k) ps[k] = from_component<S>(ps_own, k, it.j);
#pragma unroll
    for (int k = 0; k < S; ++k) {
      double acc = 0.;
#pragma unroll
      for (int kp = 0; kp < S; ++kp) acc += p.gf[k][kp] * ps[kp];
      pr = pr + 6.0 / 2. * ps[k] * acc;
    }
  }
  // velt_d = sum_m (j_m,d + F_m,d / 2) mm_m / rhot
  double vt[D];
  static_for<0, D>([&](auto d_) {
    constexpr int d = decltype(d_)::value;
    double j = 0.;
    static_for<0, Q>([&](auto n_) {
      constexpr int n = decltype(n_)::value;
      if constexpr (L::c(n, d) != 0) j += f[n] * (double)L::c(n, d);
    });
    const double term = (j + .5 * F[d]) * p.mm[it.m];
    double a = 0.;
#pragma unroll
    for (int k = 0; k < S; ++k) a += from_component<S>(term, k, it.j);
    vt[d] = a / rt;
  });
  if (!it.active) return;
  const long long o = (g.list ? (long long)__ldg(g.list + it.pos) : it.pos) - (long long)g.Rz * g.plane;
  if (it.m == 0) {
    if (rhot) rhot[o] = rt;
    if (prs) prs[o] = pr;
  }
  if (it.m == S - 1 && velt) {
#pragma unroll
    for (int d = 0; d < D; ++d) velt[o * D + d] = vt[d];
  }
}

// Adjacency table (one thread per owned position): nbr[(n-1)*fs + pos] = position of X + c_n, with
// the periodic wrap in x and y applied.  For a solid or out-of-domain neighbour the entry is some
// valid position that the mask bit keeps from being used.
template <class L>
__global__ void k_build_nbr_all(Grid g, uint32_t *__restrict__ nbr) {
  const long long pos = g.own0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= g.own1) return;
  const unsigned oe = g.list ? g.list[pos] : (unsigned)pos;
  int x, y;
  xy_of(g, oe, x, y);
  const int dxm = wrap_delta(x, -1, g.NX, g.perx), dxp = wrap_delta(x, 1, g.NX, g.perx);
  const int dym = wrap_delta(y, -1, g.NY, g.pery) * g.NX, dyp = wrap_delta(y, 1, g.NY, g.pery) * g.NX;
  const int plane = (int)g.plane;
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    const int delta = (L::c(n, 0) == 0 ? 0 : (L::c(n, 0) > 0 ? dxp : dxm)) +
                      (L::c(n, 1) == 0 ? 0 : (L::c(n, 1) > 0 ? dyp : dym)) + L::c(n, 2) * plane;
    nbr[(long long)(n - 1) * g.fs + pos] = (uint32_t)pos_of(g, (long long)oe + delta);
  });
}

}  // namespace txg

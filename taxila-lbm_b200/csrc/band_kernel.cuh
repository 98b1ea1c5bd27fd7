// band_kernel.cuh -- K2 of the order-4 step as "band" blocks: forces + collide + push with
//   * no per-node adjacency table: neighbour positions and the solid mask come from the bit rows (bitrows.cuh),
//   * the neighbour densities of the Shan-Chen stencil served from shared memory: a block owns LB consecutive
//     positions of ONE plane (several x-rows) and walks them in chunks; the densities its stencil can touch are, per
//     plane z-1, z, z+1, one contiguous run of positions (rows y0-1 .. y1+1), fetched once per block by bulk copies
//     (cp.async.bulk + mbarrier) -- about (LB + 2 rows) / LB x 3 x 8 S bytes per node instead of 18 gathers of 8 S
//     bytes through L2,
//   * the populations of the block's NEXT chunk requested from L2 while the current one is computed.
// Why: the fused kernel is bound by the number of 32-byte sectors it moves through L2 (profiles/r2a_lag_tile_results.txt:
// 1.32 G read + 0.73 G write sectors per launch at 512^3, ~190 G sectors/s whatever the DRAM traffic is); the gathers are
// 40 % of its read sectors and the adjacency 10 %.
// Same arithmetic as k_step_fused (forces1_inline, common_velocity1, collide1): results are bit-identical.
// Replaces the same reference procedures as k_step_fused (fused_kernel.cuh).
#pragma once
#include "bitrows.cuh"
#include "fused_kernel.cuh"

namespace txg {

struct BandBlock {
  uint32_t first, count;  // positions [first, first + count) of one owned plane
  uint32_t win[3];        // even-aligned start of the density window of plane z-1, z, z+1 (2-D: win[0] only)
  uint32_t len[3];        // its length in doubles (even, <= BandParams::cap): rows y0-1 .. y1+1 of that plane
};

struct BandParams {
  const BitrowEntry *rows;  // [(NZl + 2 Rz) * (NY + 2)][NW]
  const uint32_t *rowend;   // [(NZl + 2 Rz) * (NY + 2)] position one past the last fluid node of the (real) row
  const uint32_t *xrow;     // [fs] x | rowid << 11
  const BandBlock *blocks;
  int NW;              // words per row
  int rows_per_plane;  // NY + 2
  int cap;             // window length in doubles per (plane, component)
};

template <class L>
struct BandGeom {
  static constexpr int NP = L::D == 3 ? 3 : 1;  // density windows per component
};

// does the lattice have a direction with (c_y, c_z) = (dy, dz)?  does one of them have c_x != 0?
template <class L>
TXG_HD constexpr bool band_row_used(int dy, int dz) {
  for (int n = 0; n < L::Q; ++n)
    if (L::c(n, 1) == dy && (L::D == 3 ? L::c(n, 2) : 0) == dz) return true;
  return false;
}
template <class L>
TXG_HD constexpr bool band_row_has_x(int dy, int dz) {
  for (int n = 0; n < L::Q; ++n)
    if (L::c(n, 1) == dy && (L::D == 3 ? L::c(n, 2) : 0) == dz && L::c(n, 0) != 0) return true;
  return false;
}

// neighbour densities out of the block's windows; outside a window (the wrapped rows of a periodic box, rows longer than
// the window) the global load gives the same value
template <class L, int S>
struct BandGather {
  const double *tile;  // [NP][S][cap] in shared memory
  const uint32_t *win;  // [NP] in shared memory
  const uint32_t *len;  // [NP] in shared memory
  unsigned cap;
  int m;
  template <int n>
  __device__ __forceinline__ double get(const double *__restrict__ psi_field, long long np) const {
    constexpr int pl = L::D == 3 ? L::c(n, 2) + 1 : 0;
    const unsigned idx = (unsigned)np - win[pl];
    if (idx < len[pl]) return tile[(size_t)(pl * S + m) * cap + idx];
    return __ldg(psi_field + np);
  }
};

// positions of the lattice neighbours of the node at (x, rowid) and its solid mask (bit n: X + c_n is solid), from the
// bit rows: one 8-byte entry per neighbour row
template <class L>
__device__ __forceinline__ void band_neighbours(const BandParams &bp, int NX, int perx, unsigned x, unsigned rowid,
                                                unsigned here, unsigned (&npos)[L::Q], uint32_t &mask) {
  const int w = (int)(x >> 5), b = (int)(x & 31u);
  const bool xlo = perx && x == 0u, xhi = perx && x == (unsigned)(NX - 1);
  mask = 0u;
  npos[0] = here;
  static_for<0, 9>([&](auto k_) {
    constexpr int k = decltype(k_)::value;
    constexpr int dy = k % 3 - 1, dz = k / 3 - 1;
    if constexpr (band_row_used<L>(dy, dz)) {
      const long long r = (long long)rowid + dz * bp.rows_per_plane + dy;
      const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(bp.rows + r * bp.NW + w));
      RowTriple t = bitrow_triple(BitrowEntry{raw.x, raw.y}, b);
      if constexpr (band_row_has_x<L>(dy, dz)) {
        // periodic x faces: the flag of the wrapped node is in the entry, its position is at the other end of the row
        if (xlo && (t.win & 1u)) t.pm = __ldg(bp.rowend + r) - 1u;
        if (xhi && (t.win & 4u)) t.pp = __ldg(&bp.rows[r * bp.NW].se) & BITROW_POSMASK;
      }
      static_for<1, L::Q>([&](auto n_) {
        constexpr int n = decltype(n_)::value;
        if constexpr (L::c(n, 1) == dy && (L::D == 3 ? L::c(n, 2) : 0) == dz) {
          constexpr int cx = L::c(n, 0);
          npos[n] = cx < 0 ? t.pm : (cx > 0 ? t.pp : t.p0);
          mask |= ((~t.win >> (cx + 1)) & 1u) << n;
        }
      });
    }
  });
  if (mask) mask |= MASK_WALLREC;  // order-4 stencil: a wall record exists iff some lattice neighbour is solid (k_build_masks)
}

// request the lines of positions [pos, pos + npos) of every population row and of xrow from L2 (one line per thread and round)
template <class L, int S>
__device__ __forceinline__ void band_prefetch(const Grid &g, const double *__restrict__ fA, const uint32_t *__restrict__ xrow,
                                              long long pos, int npos) {
  constexpr int NR = S * L::Q;
  const int fl = (npos * 8 + 127) / 128 + 1, xl = (npos * 4 + 127) / 128 + 1;  // + 1: the range need not start on a line
  const int total = NR * fl + xl;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    if (t < NR * fl) {
      const int row = t / fl, seg = t - row * fl;
      prefetch_l2(fA + (long long)row * g.fs + pos + seg * 16);
    } else {
      prefetch_l2(xrow + pos + (t - NR * fl) * 32);
    }
  }
}

// the streamed populations of the node at `here` out of the collided ones of its neighbours (pull form of stream +
// half-way bounce-back): fi_n(X) = g_n(X - c_n), or g_opp(n)(X) when X - c_n = X + c_opp(n) is solid
template <class L>
__device__ __forceinline__ void band_pull(const double *__restrict__ src, unsigned fs, unsigned here, const unsigned (&npos)[L::Q],
                                          uint32_t mask, double (&f)[L::Q]) {
  f[0] = load_population(src + here);
  static_for<1, L::Q>([&](auto n_) {
    constexpr int n = decltype(n_)::value;
    constexpr int on = opp<L>(n);
    const bool bounce = (mask >> on) & 1u;
    const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + npos[on];
    f[n] = load_population(src + e);
  });
}

#ifndef TXG_BAND_THREADS
#define TXG_BAND_THREADS 256
#endif

// PULL = false: fA holds the streamed populations of the step (the reference's fi); the collided ones are pushed into
//   fB (bounce-back folded into the store) -- stores of a warp split between the pushed rows and the bounce-back rows.
// PULL = true: fA holds the COLLIDED populations g of the previous step at their own nodes; the node gathers
//   fi_n(X) = g_n(X - c_n), or g_opp(n)(X) when X - c_n is solid (DistributionStreamD* + DistributionBouncebackD*,
//   lbm_distribution_function.F90:560-784, read from the other side), collides, and stores its 19 values at its own
//   position: every store of a half-warp is one aligned 128-byte run, the fragmentation sits on the load side where a
//   partly used sector costs nothing but its fetch.  `identity` != 0 (warp-uniform): fA already holds the streamed
//   populations (first step after FlowFiInit / a restart / an export) and is read in place.
template <class L, int S, bool MRT, bool PULL>
__global__ void __launch_bounds__(TXG_BAND_THREADS, 512 / TXG_BAND_THREADS)
    k_step_band(Grid g, Phys p, BandParams bp, const double *__restrict__ fA, double *__restrict__ fB,
                const double *__restrict__ rho, const double *__restrict__ wallrec, int block0, int prefetch, int identity,
                int *__restrict__ gave_up) {
  constexpr int Q = L::Q, D = L::D, ISO = 4, NP = BandGeom<L>::NP, NPW = Lanes<S>::NPW;
  constexpr int CH = (TXG_BAND_THREADS / 32) * NPW;  // positions per chunk
  extern __shared__ __align__(128) double tile[];     // [NP][S][cap]
  __shared__ __align__(8) uint64_t bar;
  __shared__ BandBlock sb;
  if (threadIdx.x == 0) {
    sb = bp.blocks[block0 + blockIdx.x];
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned cap = (unsigned)bp.cap;
  if (threadIdx.x == 0) {
    unsigned total = 0;
#pragma unroll
    for (int pl = 0; pl < NP; ++pl) total += sb.len[pl];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"((unsigned)S * total * 8u) : "memory");
#pragma unroll
    for (int pl = 0; pl < NP; ++pl)
#pragma unroll
      for (int m = 0; m < S; ++m)
        if (sb.len[pl])  // (a window without fluid nodes: nothing to copy)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(tile + (size_t)(pl * S + m) * cap)),
                     "l"(rho + (long long)m * g.fs + sb.win[pl]), "r"(sb.len[pl] * 8u), "r"(smem_addr(&bar))
                     : "memory");
  }
  const long long first = sb.first, count = sb.count;
  const int warp = threadIdx.x >> 5;
  bool landed = false;
  // xrow of this lane's node, requested one chunk ahead (the gathers of the pull form wait for it and for the bit rows)
  const int lane_j = (threadIdx.x & 31) % NPW;
  auto xrow_of_chunk = [&](long long c0) -> uint32_t {
    const long long i = c0 + (long long)warp * NPW;
    if (i >= count) return 0u;
    return __ldg(bp.xrow + first + min(i + lane_j, count - 1));
  };
  uint32_t xr_next = xrow_of_chunk(0);
  for (long long c0 = 0; c0 < count; c0 += CH) {
    if (prefetch && (!PULL || identity) && c0 + CH < count)
      band_prefetch<L, S>(g, fA, bp.xrow, first + c0 + CH, (int)min((long long)CH, count - c0 - CH));
    Item it;
    if (!item_of_lane<S>(first, count, (c0 / NPW) + warp, it)) continue;
    const uint32_t xr = xr_next;
    xr_next = xrow_of_chunk(c0 + CH);
    double f[Q];
    if (!PULL || identity) {
      const double *src = fA + (long long)it.m * Q * g.fs + it.pos;
#pragma unroll
      for (int n = 0; n < Q; ++n) f[n] = load_population(src + (long long)n * g.fs);
    }
    unsigned npos[Q];
    uint32_t mask;
    band_neighbours<L>(bp, g.NX, g.perx, xr & ((1u << BITROW_XBITS) - 1u), xr >> BITROW_XBITS, (unsigned)it.pos, npos, mask);
    if (PULL && !identity) band_pull<L>(fA + (long long)it.m * Q * g.fs, (unsigned)g.fs, (unsigned)it.pos, npos, mask, f);
    const double *psi_field = rho + (long long)it.m * g.fs;
    double r = 0.;
#pragma unroll
    for (int n = 0; n < Q; ++n) r += f[n];
    const double psi_m = p.eos ? __ldg(psi_field + it.pos) : r;
    if (!landed) {
      tile_wait(&bar, gave_up);
      landed = true;
    }
    double F[D];
    const BandGather<L, S> gather{tile, sb.win, sb.len, cap, it.m};
    forces1_inline<L, S, ISO>(g, p, psi_field, nullptr, wallrec, it, 0u, 0, 0, mask, npos, r, psi_m, F, gather);
    double up[D];
    common_velocity1<L, S>(p, it, f, r, F, up);
    collide1<L, MRT>(p, it.m, r, F, up, f);
    if (it.active) {
      double *out = fB + (long long)it.m * Q * g.fs;
      const unsigned fs = (unsigned)g.fs, here = (unsigned)it.pos;
      if constexpr (PULL) {
#pragma unroll
        for (int n = 0; n < Q; ++n) store_population(out + (unsigned)n * fs + here, f[n]);
      } else {
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
  }
  // a block must not retire while its bulk copies are in flight (shared memory would be re-assigned under them)
  if (!landed && threadIdx.x < 32) tile_wait(&bar, gave_up);
}

// K1 of the pull form: rho_m(X) = sum_n fi_n(X) (ascending n, like k_moments) with fi gathered from the collided
// populations of the neighbours.  WRITE_F: also (materialise) store the gathered populations at their own node in fOut
// -- the reference's fi after DistributionStream + Bounceback -- for the exports, the restart files and every path that
// works on streamed populations; rho is not written then.
// Replaces DistributionStreamD*, DistributionBouncebackD* (lbm_distribution_function.F90:560-784), DistributionCalcDensityD*
// (:379-428) and EOSApply.
template <class L, int S, bool WRITE_F>
__global__ void __launch_bounds__(128, 8) k_moments_pull(Grid g, Phys p, BandParams bp, const double *__restrict__ gA,
                                                        double *__restrict__ fOut, double *__restrict__ rho,
                                                        double *__restrict__ rho_true, long long first, long long count) {
  constexpr int Q = L::Q;
  Item it;
  if (!item_of_lane<S>(first, count, it)) return;
  const uint32_t xr = __ldg(bp.xrow + it.pos);
  unsigned npos[Q];
  uint32_t mask;
  band_neighbours<L>(bp, g.NX, g.perx, xr & ((1u << BITROW_XBITS) - 1u), xr >> BITROW_XBITS, (unsigned)it.pos, npos, mask);
  double f[Q];
  band_pull<L>(gA + (long long)it.m * Q * g.fs, (unsigned)g.fs, (unsigned)it.pos, npos, mask, f);
  if (!it.active) return;
  if constexpr (WRITE_F) {
    double *out = fOut + (long long)it.m * Q * g.fs + it.pos;
#pragma unroll
    for (int n = 0; n < Q; ++n) out[(long long)n * g.fs] = f[n];
  } else {
    double a = 0.;
#pragma unroll
    for (int n = 0; n < Q; ++n) a += f[n];
    if (p.eos) {
      rho_true[(long long)it.m * g.fs + it.pos] = a;
      a = eos_psi(p, it.m, a);
    }
    rho[(long long)it.m * g.fs + it.pos] = a;
  }
}

}  // namespace txg

// stream_kernel.cuh -- the collide + push kernel fed by the TMA engine (sm_100a).
//
// Same arithmetic and the same lane mapping as k_collide (hot_kernels.cuh); what changes is who
// moves the bytes.  The fp64 collision keeps ~126 registers per lane live, so an SM holds 16 warps
// at most, and the loads those few warps can keep in flight from registers bound k_collide at
// ~4.7 TB/s of DRAM traffic with neither HBM, L2 nor the fp64 pipe saturated (ncu: 60 % of the
// warp time is long-scoreboard stall, profiles/).  Here each block is persistent and walks over
// chunks of PB = STREAM_WARPS * NPW consecutive positions of the fluid list; every row the chunk reads at
// addresses that follow from the position alone -- the S*Q population rows, the Q-1 adjacency
// rows and the mask row -- arrives in shared memory as one bulk asynchronous copy per row
// (cp.async.bulk, completion counted on an mbarrier), issued one to two chunks ahead by a single
// elected lane.  The copies hold no registers and need no resident warp, so the bytes in flight
// per SM no longer depend on the occupancy.  What is left as ordinary loads are the gathers whose
// addresses come out of the adjacency row (neighbour densities) and the wall record.
//
//   iteration k of a warp (stage s = k % 2)
//     wait full[s]                          chunk k has landed
//     mask, adjacency, populations -> registers; arrive on empty[s]
//     warp 0 only: wait empty[s] (all warps of the block have taken chunk k), then one bulk copy per lane
//                  and round refills stage s with chunk k+2 while k and k+1 are worked on
//     density gathers, forces, velocity, collision, push stores
//   There is no block-wide barrier in the loop: a warp waits only for data.
//
// A block takes its chunks in strips of `strip` consecutive chunks (then jumps ahead by
// gridDim.x strips), so that it sweeps a few consecutive x-rows of the lattice: the density rows
// gathered as y- and z-neighbours of one x-row are the ones the next x-row gathers again, and
// they are then still in the SM's L1.
//
// Chunks are aligned in absolute position space (chunk c = positions [c*PB, (c+1)*PB)), so every
// copy starts on a 256-byte boundary; lanes whose position falls outside [first, first+count)
// compute on whatever the row holds there and store nothing.
#pragma once
#include "hot_kernels.cuh"

namespace txg {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy (TMA engine); bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 bulk prefetch (TMA engine, no destination): bytes a multiple of 16, address 16-byte aligned
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// L2 eviction-priority policies: the population / adjacency streams are touched once per step
// (evict first); the density rows are re-read by the neighbouring planes (evict last)
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, unsigned bytes, uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void st_hint(double *ptr, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(ptr), "d"(v), "l"(pol) : "memory");
}

enum StreamOpts : int {
  STREAM_HINTS = 1,        // L2 eviction hints on the streams and the push stores
  STREAM_PF_WALLREC = 2,   // L2 bulk prefetch of the wall-record rows of the chunk two ahead
  STREAM_PF_RHO = 4,       // L2 bulk prefetch of the density rows above the next chunk
};

// warps per block: 8 makes every population row of a chunk a 1 KB bulk copy (S = 2); the TMA engine
// sustains ~5.9 TB/s with copies of that size against ~5.0 TB/s with 512-byte ones (tools/membench)
#ifndef TXG_STREAM_WARPS
#define TXG_STREAM_WARPS 8
#endif
constexpr int STREAM_WARPS = TXG_STREAM_WARPS;
constexpr int STREAM_THREADS = 32 * STREAM_WARPS;

template <class L, int S>
struct StreamStage {
  static constexpr int PB = STREAM_WARPS * Lanes<S>::NPW;  // positions per chunk (one item per warp)
  double f[S * L::Q][PB];
  uint32_t nbr[L::Q - 1][PB];
  uint32_t mask[PB];
  static constexpr unsigned BYTES = (unsigned)(sizeof(double) * S * L::Q * PB + sizeof(uint32_t) * (L::Q - 1) * PB +
                                               sizeof(uint32_t) * PB);
};

constexpr int STREAM_STAGES = 2;

template <class L, int S>
struct StreamSmem {
  StreamStage<L, S> st[STREAM_STAGES];
  uint64_t full[STREAM_STAGES];   // completes when the bulk copies of the stage have landed
  uint64_t empty[STREAM_STAGES];  // completes when every warp of the block has read the stage
};
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// warp 0: arm the barrier and issue the row copies of chunk c into stage `st`, one row per lane and round
template <class L, int S>
__device__ __forceinline__ void stream_issue(const Grid &g, StreamStage<L, S> &st, uint64_t *bar,
                                             const double *__restrict__ fA, const uint32_t *__restrict__ lmask,
                                             const uint32_t *__restrict__ nbr, long long c, int lane, bool hints,
                                             uint64_t pol) {
  constexpr int Q = L::Q, PB = StreamStage<L, S>::PB, NF = S * Q, NN = Q - 1;
  const long long pos = c * PB;
  if (lane == 0) mbar_expect_tx(bar, StreamStage<L, S>::BYTES);
  __syncwarp();
  for (int row = lane; row < NF + NN + 1; row += 32) {
    void *dst;
    const void *src;
    unsigned bytes;
    if (row < NF) {
      dst = &st.f[row][0];
      src = fA + (long long)row * g.fs + pos;
      bytes = PB * 8;
    } else if (row < NF + NN) {
      dst = &st.nbr[row - NF][0];
      src = nbr + (long long)(row - NF) * g.fs + pos;
      bytes = PB * 4;
    } else {
      dst = &st.mask[0];
      src = lmask + pos;
      bytes = PB * 4;
    }
    if (hints)
      bulk_g2s_hint(dst, src, bytes, bar, pol);
    else
      bulk_g2s(dst, src, bytes, bar);
  }
}

#ifndef TXG_STREAM_MIN_BLOCKS
#define TXG_STREAM_MIN_BLOCKS (16 / TXG_STREAM_WARPS)
#endif
template <class L, int S, bool MRT, int ISO>
__global__ void __launch_bounds__(STREAM_THREADS, TXG_STREAM_MIN_BLOCKS)
    k_collide_stream(Grid g, Phys p, const double *__restrict__ fA, double *__restrict__ fB,
                     const double *__restrict__ rho, const uint32_t *__restrict__ lmask,
                     const uint32_t *__restrict__ nbr, const uint32_t *__restrict__ ffmask,
                     const double *__restrict__ wallrec, long long first, long long count, int opts, int strip) {
  constexpr int Q = L::Q, D = L::D, NPW = Lanes<S>::NPW, PB = StreamStage<L, S>::PB;
  extern __shared__ __align__(128) unsigned char stream_smem_raw[];
  StreamSmem<L, S> &sm = *reinterpret_cast<StreamSmem<L, S> *>(stream_smem_raw);

  const long long c0 = first / PB, c1 = (first + count + PB - 1) / PB;  // chunks [c0, c1)
  const long long nchunks = c1 - c0;
  // strips of `strip` consecutive chunks, dealt round-robin to the blocks
  const long long nstrips = (nchunks + strip - 1) / strip;
  const long long my_strips = ((nstrips - (long long)blockIdx.x) + (long long)gridDim.x - 1) / (long long)gridDim.x;
  if (my_strips <= 0) return;  // block-uniform
  // the last strip may be short; it belongs to block (nstrips-1) % gridDim.x as its last one
  const long long last_len = nchunks - (nstrips - 1) * strip;
  const bool has_last = ((nstrips - 1) % (long long)gridDim.x) == (long long)blockIdx.x;
  const long long mine = my_strips * strip - (has_last ? strip - last_len : 0);
  auto chunk_of = [&](long long k) {
    const long long v = k / strip, w = k - v * strip;
    return c0 + ((long long)blockIdx.x + v * (long long)gridDim.x) * strip + w;
  };
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool hints = (opts & STREAM_HINTS) != 0;
  const uint64_t pol_stream = policy_evict_first();

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STREAM_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], STREAM_WARPS);
    }
    mbar_init_fence();
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int s = 0; s < STREAM_STAGES; ++s)
      if (s < mine) stream_issue<L, S>(g, sm.st[s], &sm.full[s], fA, lmask, nbr, chunk_of(s), lane, hints, pol_stream);
  }

  Item it;
  it.m = lane / NPW;
  it.j = lane - it.m * NPW;
  const bool lane_ok = it.m < S;
  if (!lane_ok) it.m = S - 1;
  const int col = warp * NPW + it.j;  // column of this lane in the stage
  const double *psi_field = rho + (long long)it.m * g.fs;
  double *out = fB + (long long)it.m * Q * g.fs;
  const unsigned fs = (unsigned)g.fs;

  for (long long k = 0; k < mine; ++k) {
    const int s = (int)(k % STREAM_STAGES);
    const unsigned parity = (unsigned)((k / STREAM_STAGES) & 1);
    StreamStage<L, S> &st = sm.st[s];
    it.pos = chunk_of(k) * PB + col;
    it.active = lane_ok && it.pos >= first && it.pos < first + count;
    mbar_wait(&sm.full[s], parity);

    const uint32_t mask = st.mask[col];
    unsigned npos[Q];
    npos[0] = (unsigned)it.pos;
#pragma unroll
    for (int n = 1; n < Q; ++n) npos[n] = st.nbr[n - 1][col];
    double f[Q];
#pragma unroll
    for (int n = 0; n < Q; ++n) f[n] = st.f[it.m * Q + n][col];
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);  // this warp has taken its part of stage s
    if (warp == 0) {
      if (k + STREAM_STAGES < mine) {
        mbar_wait(&sm.empty[s], parity);  // ... and so have the others
        const long long cn = chunk_of(k + STREAM_STAGES);
        stream_issue<L, S>(g, st, &sm.full[s], fA, lmask, nbr, cn, lane, hints, pol_stream);
        if ((opts & STREAM_PF_WALLREC) && wallrec && lane < S * D + D)
          bulk_prefetch_l2(wallrec + (long long)lane * g.fs + cn * PB, PB * 8);
      }
      if constexpr (D == 3) {
        // the density rows the NEXT chunk gathers from the plane above are first touched by this
        // sweep: ask L2 for them one iteration early (their position comes out of the adjacency
        // row of the next chunk, which has normally landed by now)
        if ((opts & STREAM_PF_RHO) && k + 1 < mine && lane < S) {
          const int s1 = (int)((k + 1) % STREAM_STAGES);
          if (mbar_try_wait(&sm.full[s1], (unsigned)(((k + 1) / STREAM_STAGES) & 1))) {
            constexpr int nz = dir_of<L>(0, 0, 1);
            const long long zp = (long long)sm.st[s1].nbr[nz - 1][0];
            long long a = (zp - 32) & ~1ll;
            if (a < 0) a = 0;
            long long b = a + PB + 96;
            if (b > g.fs) b = g.fs;
            bulk_prefetch_l2(rho + (long long)lane * g.fs + a, (unsigned)((b - a) * 8));
          }
        }
      }
    }
    unsigned oe = 0;
    int x = 0, y = 0;
    if constexpr (ISO != 4) {  // wider stencils look their extra neighbours up through P
      const long long lp = it.active ? it.pos : first;  // lanes outside the range: any valid node
      oe = g.list ? __ldg(g.list + lp) : (unsigned)lp;
      xy_of(g, oe, x, y);
    }
    double r = 0.;
#pragma unroll
    for (int n = 0; n < Q; ++n) r += f[n];
    const double psi_m = p.eos ? __ldg(psi_field + it.pos) : r;
    double F[D];
    forces1<L, S, ISO>(g, p, psi_field, ffmask, wallrec, it, oe, x, y, mask, npos, r, psi_m, F);
    double up[D];
    common_velocity1<L, S>(p, it, f, r, F, up);
    collide1<L, MRT>(p, it.m, r, F, up, f);
    if (it.active) {
      const unsigned here = (unsigned)it.pos;
      if (hints) {
        st_hint(out + here, f[0], pol_stream);
        static_for<1, Q>([&](auto n_) {
          constexpr int n = decltype(n_)::value;
          constexpr int on = opp<L>(n);
          const bool bounce = (mask >> n) & 1u;
          const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + npos[n];
          st_hint(out + e, f[n], pol_stream);
        });
      } else {
        out[here] = f[0];
        static_for<1, Q>([&](auto n_) {
          constexpr int n = decltype(n_)::value;
          constexpr int on = opp<L>(n);
          const bool bounce = (mask >> n) & 1u;
          const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + npos[n];
          out[e] = f[n];
        });
      }
    }
  }
}

}  // namespace txg

// stage_kernel.cuh -- K2 of the order-4 step with its streamed inputs staged in shared memory by the tensor copy engine
// (TMA), one double-buffered pipeline PER WARP.  Opt-in (TXG_STAGE=1).
//
// k_step_fused is bound by how many bytes an SM keeps in flight, not by DRAM or the fp64 pipe (profiles/
// r1k_step_fused_ncu_summary.txt: DRAM 56 %, fp64 37 %, issue 42 %, 57 % of the warp time long-scoreboard; without its
// stores it still reads at only 3.2 TB/s, profiles/r1b_ablations.txt): 126 registers leave 16 warps per SM, and a warp
// holds its 19 population loads plus the adjacency row only while it waits for them -- during the ~1300 fp64 issue cycles
// of the collision it has nothing in flight.  Here the rows a warp streams for its NPW positions -- S * Q population rows,
// Q - 1 adjacency rows and the mask row -- are two boxes of two 2-D tensors ([S * Q][fs] doubles, [Q][fs] words), so ONE
// lane fetches the warp's NEXT item with two cp.async.bulk.tensor.2d copies (completion counted on the warp's own
// mbarrier) into the second stage of the warp's double buffer while the warp collides the current item out of the first:
// the copies need no registers and stay in flight for the whole collision.  History (profiles/r2d_stage_results.txt): a
// pipeline per 256-thread block lost 18 % of its time at __syncthreads (two instruction streams per SM where the fused
// kernel has 16); a pipeline per warp with one 1-D bulk copy per row (57 per item) lost 17 % at the copy instruction.
// Blocks are aligned on absolute multiples of LB positions.  What remains on the demand path: the 18 density gathers and
// the wall record (issued as soon as the adjacency row is read from shared memory: one exposed round trip, mostly L2 hits)
// and the scattered push stores.  Same arithmetic in the same order as k_step_fused: results are bit-identical.
// Replaces the same reference procedures as k_step_fused (fused_kernel.cuh).
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint in flow.cu)

#include "fused_kernel.cuh"

namespace txg {

// Compressed adjacency of one item (TXG_STAGE_ADJC; k_build_adjc below): positions run along x, so the neighbour positions of the
// ITEM consecutive nodes of an item along one direction are a base plus small offsets.  Record of item i (positions
// [i * ITEM, (i + 1) * ITEM) of the storage): flags word (bit 0: "escape" -- some direction's offsets do not fit a byte, the lanes read
// the full table nbr_all), Q - 1 bases, ITEM mask words, (Q - 1) x ITEM byte offsets; 432 bytes for D3Q19 against the 1216 of the
// adjacency + mask rows of the item (27 instead of 76 bytes per node).  One contiguous bulk copy per item.
template <class L, int S>
struct AdjcGeom {
  static constexpr int NPW = Lanes<S>::NPW, ITEM = (NPW + 3) & ~3, NB = L::Q - 1;
  static constexpr int OFF_BASE = 4, OFF_MASK = OFF_BASE + 4 * NB, OFF_DELTA = OFF_MASK + 4 * ITEM;
  static constexpr int REC_BYTES = (OFF_DELTA + NB * ITEM + 15) & ~15;
};

template <class L, int S, int NWARPS = 4, bool ADJC = false>
struct StageGeom {
  static constexpr int NW = NWARPS, NT = 32 * NW;         // warps, threads per block
  static constexpr int NPW = Lanes<S>::NPW;               // positions per warp item
  static constexpr int ITEM = (NPW + 3) & ~3;             // staged run per row (inner box: a multiple of 16 bytes for u32 rows too)
  static constexpr int NF = S * L::Q, NA = L::Q;          // population rows; adjacency rows + the mask row
  static constexpr int F_BYTES = NF * ITEM * 8, A_BYTES = ADJC ? AdjcGeom<L, S>::REC_BYTES : NA * ITEM * 4;
  static constexpr int A_OFF = (F_BYTES + 127) & ~127;    // tensor copies land on 128-byte boundaries
  static constexpr int STAGE_BYTES = (A_OFF + A_BYTES + 127) & ~127;
  static constexpr int WARP_BYTES = 2 * STAGE_BYTES;
  static constexpr int SMEM_BYTES = NW * WARP_BYTES;
  // 4 warps: 48 KB per block for S = 2, four blocks per SM and no room left for L1.  The wider blocks (6 warps x 2 blocks,
  // 12 warps x 1 block per SM) leave ~90 KB of L1 to ONE or TWO contiguous runs of positions per SM, so that the density
  // rows a block gathers from (its own rows +- 1 in three planes) can stay in L1 from one round of the block to the next.
  static constexpr int BLOCKS_PER_SM = NW == 4 ? (S == 2 ? 4 : 3) : NW == 6 ? 2 : 1;
};

// One lane of a warp: fetch the item of positions [p0, p0 + ITEM) -- the box {ITEM positions} x {all S * Q population
// rows} of the population tensor and the box {ITEM} x {Q - 1 adjacency rows + mask row} of the adjacency tensor -- into the
// warp's stage at `dst` with TWO descriptor-based tensor copies (cp.async.bulk.tensor.2d, tile mode): the copy engine
// walks the rows itself.  (A first per-warp form issued one cp.async.bulk per row, 57 per item, and lost 17 % of its time
// at the copy instruction: profiles/r2d_stage_results.txt.)  Positions past the end of a row are zero-filled.
template <class L, int S, bool ADJC = false>
__device__ __forceinline__ void stage_issue(unsigned char *dst, uint64_t *bar, const CUtensorMap *tmF, const CUtensorMap *tmA,
                                            long long p0, const unsigned char *__restrict__ adjc = nullptr) {
  using G = StageGeom<L, S, 4, ADJC>;
  // the stage was read by plain loads: order them before the asynchronous writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((unsigned)(G::F_BYTES + G::A_BYTES)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_addr(dst)),
                 "l"(tmF), "r"((int)p0), "r"(0), "r"(smem_addr(bar))
                 : "memory");
    if constexpr (ADJC) {
      // the item's compressed adjacency record: one contiguous bulk copy (p0 is a multiple of ITEM)
      const unsigned char *src = adjc + (p0 / G::ITEM) * (long long)G::A_BYTES;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst + G::A_OFF)),
                   "l"(src), "r"((unsigned)G::A_BYTES), "r"(smem_addr(bar))
                   : "memory");
    } else {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                       smem_addr(dst + G::A_OFF)),
                   "l"(tmA), "r"((int)p0), "r"(0), "r"(smem_addr(bar))
                   : "memory");
    }
  }
}

// one lane: ask L2 for the two boxes of the item at p0 (descriptor-based prefetch, no shared memory, no completion)
__device__ __forceinline__ void stage_prefetch_l2(const CUtensorMap *tmF, const CUtensorMap *tmA, long long p0) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmF), "r"((int)p0), "r"(0) : "memory");
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmA), "r"((int)p0), "r"(0) : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

__device__ __forceinline__ void stage_wait(uint64_t *bar, unsigned parity) {
  unsigned ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_addr(bar)), "r"(parity)
                 : "memory");
    if (!ok && ++spins > (1u << 24)) asm volatile("trap;");  // a copy that never lands fails the launch instead of hanging the device
  }
}

// The collision of ONE staged item: the NPW positions from w0 on, operands in the stage at `sm` (already landed), lane =
// (component m, node slot j); positions outside [lo, hi) are replayed on a valid position of the item and not stored.
template <class L, int S, bool MRT, bool ADJC = false>
__device__ __forceinline__ void stage_item(const Grid &g, const Phys &p, double *__restrict__ fB, const double *__restrict__ rho,
                                           const double *__restrict__ wallrec, long long lo, long long hi, long long w0,
                                           const unsigned char *sm, int m, int j, bool lane_ok,
                                           const uint32_t *__restrict__ nbr_all = nullptr, const unsigned char *sm_next = nullptr,
                                           uint64_t *bar_next = nullptr, unsigned par_next = 0, long long w_next = 0) {
  using G = StageGeom<L, S, 4, ADJC>;
  constexpr int Q = L::Q, D = L::D, ISO = 4, NPW = G::NPW, ITEM = G::ITEM;
  Item it;
  it.m = m;
  it.j = j;
  long long pos = w0 + j;
  it.active = lane_ok && pos >= lo && pos < hi;
  pos = min(max(pos, max(lo, w0)), min(hi, w0 + NPW) - 1);  // replayed lanes: a valid position of this item
  it.pos = pos;
  const int i = (int)(pos - w0);
  double f[Q];
  unsigned npos[Q];
  npos[0] = (unsigned)pos;
  uint32_t mask;
  if constexpr (ADJC) {
    using A = AdjcGeom<L, S>;
    const unsigned char *rec = sm + G::A_OFF;
    const uint32_t *base = reinterpret_cast<const uint32_t *>(rec + A::OFF_BASE);
    const unsigned char *delta = rec + A::OFF_DELTA + i;
    mask = reinterpret_cast<const uint32_t *>(rec + A::OFF_MASK)[i];
    if (*reinterpret_cast<const uint32_t *>(rec) & 1u) {  // (warp-uniform) escape: offsets that do not fit a byte
#pragma unroll
      for (int n = 1; n < Q; ++n) npos[n] = __ldg(nbr_all + (long long)(n - 1) * g.fs + pos);
    } else {
#pragma unroll
      for (int n = 1; n < Q; ++n) npos[n] = base[n - 1] + delta[(n - 1) * ITEM];
    }
  } else {
    const uint32_t *sa = reinterpret_cast<const uint32_t *>(sm + G::A_OFF) + i;
#pragma unroll
    for (int n = 1; n < Q; ++n) npos[n] = sa[(n - 1) * ITEM];
    mask = sa[(Q - 1) * ITEM];
  }
  const double *sf = reinterpret_cast<const double *>(sm) + (size_t)m * Q * ITEM + i;
#pragma unroll
  for (int n = 0; n < Q; ++n) f[n] = sf[n * ITEM];
  const double *psi_field = rho + (long long)it.m * g.fs;
  double r = 0.;
#pragma unroll
  for (int n = 0; n < Q; ++n) r += f[n];
  const double psi_m = p.eos ? __ldg(psi_field + it.pos) : r;
  double F[D];
  forces1_inline<L, S, ISO>(g, p, psi_field, nullptr, wallrec, it, 0u, 0, 0, mask, npos, r, psi_m, F);
  // TXG_STAGE_PG=1: the densities and the wall record of the warp's NEXT item into L1 (prefetch.global.L1 = CCTL.E.PF1) while this
  // item is collided -- the one round trip no stage hides (27 % of the warp time, profiles/r2z_step_stage_ncu_summary.txt).  The next
  // item's adjacency has been in flight since the start of this item.
  if (sm_next) {
    stage_wait(bar_next, par_next);
    const long long pn = w_next + j;  // (slots past the end of the launch hold whatever the copy brought: clamped, never dereferenced)
    uint32_t mask_n;
    const unsigned top = (unsigned)g.fs - 1u;
    if constexpr (ADJC) {
      using A = AdjcGeom<L, S>;
      const unsigned char *rec = sm_next + G::A_OFF;
      const uint32_t *base = reinterpret_cast<const uint32_t *>(rec + A::OFF_BASE);
      const unsigned char *delta = rec + A::OFF_DELTA + j;
      mask_n = reinterpret_cast<const uint32_t *>(rec + A::OFF_MASK)[j];
      if (!(*reinterpret_cast<const uint32_t *>(rec) & 1u)) {
#pragma unroll
        for (int n = 1; n < Q; ++n) prefetch_l1(psi_field + min(base[n - 1] + delta[(n - 1) * ITEM], top));
      }
    } else {
      const uint32_t *sa = reinterpret_cast<const uint32_t *>(sm_next + G::A_OFF) + j;
      mask_n = sa[(Q - 1) * ITEM];
#pragma unroll
      for (int n = 1; n < Q; ++n) prefetch_l1(psi_field + min(sa[(n - 1) * ITEM], top));
    }
    if ((mask_n & MASK_WALLREC) && pn < g.fs) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (p.fluidsolid) prefetch_l1(wallrec + (long long)(it.m * D + d) * g.fs + pn);
        if (p.fluidfluid) prefetch_l1(wallrec + (long long)(S * D + d) * g.fs + pn);
      }
    }
    if (p.eos && pn < g.fs) prefetch_l1(psi_field + pn);
  }
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
#ifndef TXG_ABL_STORE
      const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + npos[n];
      store_population(out + e, f[n]);
#elif TXG_ABL_STORE == 1  // timing ablations only (wrong results; tools/build_variants_r2b.sh): no bounce-back stores
      const unsigned e = (unsigned)n * fs + npos[n];
      if (!bounce) store_population(out + e, f[n]);
      (void)on;
#elif TXG_ABL_STORE == 2  // every store aligned on the node's own position
      const unsigned e = (unsigned)n * fs + here;
      store_population(out + e, f[n]);
      (void)on;
      (void)bounce;
#elif TXG_ABL_STORE == 3  // no stores
      const unsigned e = bounce ? (unsigned)on * fs + here : (unsigned)n * fs + npos[n];
      if (f[n] == 1.2345e300) store_population(out + e, f[n]);
#elif TXG_ABL_STORE == 4  // pushes only where no lane of the half-warp bounces... (bounce lanes write their own slot of row n)
      const unsigned e = bounce ? (unsigned)n * fs + here : (unsigned)n * fs + npos[n];
      store_population(out + e, f[n]);
      (void)on;
#endif
    });
  }
}

// One launch covers the positions [first, first + count); block b owns the absolute positions [(blk0 + b) * LB, + LB) of it
// (LB a multiple of NW * NPW; the host passes blk0 = first / LB); inside a block, item k = positions [k * NPW, + NPW) of the
// block goes to warp k % NW, so that the block's warps walk one contiguous run of positions together.  SHORT blocks (two
// rounds: LB = 2 * NW * NPW = 128 positions for S = 2) measured fastest: the resident blocks then cover one compact,
// advancing window of positions like k_step_fused does.  Longer blocks, and persistent warps drawing their items from a
// ticket counter (one at a time: 11 % of the warp time at the atomic; four at a time: the resident warps spread over
// several planes and the density gathers start missing L2), were slower -- profiles/r2d_stage_results.txt.
template <class L, int S, bool MRT, int NWARPS = 4, bool ADJC = false>
__global__ void __launch_bounds__(StageGeom<L, S, NWARPS, ADJC>::NT, StageGeom<L, S, NWARPS, ADJC>::BLOCKS_PER_SM)
    k_step_stage(Grid g, Phys p, const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmA,
                 double *__restrict__ fB, const double *__restrict__ rho, const double *__restrict__ wallrec, long long first,
                 long long count, long long blk0, int LB, int pf_blocks, const double *__restrict__ fA_rows,
                 const uint32_t *__restrict__ adj_rows, const unsigned char *__restrict__ adjc, const uint32_t *__restrict__ nbr_all, int pg) {
  using G = StageGeom<L, S, NWARPS, ADJC>;
  constexpr int NPW = G::NPW, NW = G::NW;
  extern __shared__ __align__(1024) unsigned char stage_mem[];  // [NW][2][STAGE_BYTES]
  __shared__ __align__(8) uint64_t bars[NW][2];
  const long long last = first + count;               // one past the last position of the launch
  const long long b0 = (blk0 + blockIdx.x) * LB;      // absolute positions of this block: [b0, b0 + LB)
  const long long lo = max(b0, first), hi = min(b0 + LB, last);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pf_blocks > 0: first ask L2 for this warp's items of the block pf_blocks further on -- about a quarter of a wave of
  // resident blocks ahead -- so that the copies of that block's FIRST items (the ones no pipeline has prefetched) hit L2
  if (pf_blocks > 0 && lane == 0) {
    const long long pb = b0 + (long long)pf_blocks * LB;
    for (int k = warp; k * NPW < LB; k += NW) {
      const long long p0 = pb + (long long)k * NPW;
      if (p0 < last) stage_prefetch_l2(&tmF, &tmA, p0);
    }
  }
  // pf_blocks < 0: the same with plain prefetch.global.L2 instructions, one 128-byte line per thread and round (what
  // k_step_fused does for its rows)
  if (pf_blocks < 0) {
    const long long pb = b0 + (long long)(-pf_blocks) * LB;
    if (pb < last) {
      const int fl = LB / 16, al = LB / 32;  // lines per population row / per adjacency row of one block
      const int total = G::NF * fl + G::NA * al;
      for (int t = threadIdx.x; t < total; t += G::NT) {
        if (t < G::NF * fl) {
          const int row = t / fl, seg = t - row * fl;
          prefetch_l2(fA_rows + (long long)row * g.fs + pb + seg * 16);
        } else {
          const int u = t - G::NF * fl, row = u / al, seg = u - row * al;
          prefetch_l2(adj_rows + (long long)row * g.fs + pb + seg * 32);
        }
      }
    }
  }
  if (lo >= hi) return;
  // items of this warp that hold positions of the launch: k = warp (mod NW), k_lo <= k <= k_hi
  int k_lo = (int)((lo - b0) / NPW), k_hi = (int)((hi - 1 - b0) / NPW);
  k_lo += (warp - k_lo % NW + NW) % NW;
  if (k_lo > k_hi) return;
  unsigned char *wmem = stage_mem + (size_t)warp * G::WARP_BYTES;
  uint64_t *wbar = bars[warp];
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&wbar[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&wbar[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  stage_issue<L, S, ADJC>(wmem, &wbar[0], &tmF, &tmA, b0 + (long long)k_lo * NPW, adjc);
  // lane -> (component, node slot) like item_of_lane
  int m = lane / NPW;
  const int j = lane - m * NPW;
  bool lane_ok = true;
  if (m >= S) {
    m = S - 1;
    lane_ok = false;
  }
  int round = 0;
  for (int k = k_lo; k <= k_hi; k += NW, ++round) {
    const int st = round & 1;
    // the other stage was read to the end in the previous round: refill it with this warp's next item
    if (k + NW <= k_hi)
      stage_issue<L, S, ADJC>(wmem + (size_t)(st ^ 1) * G::STAGE_BYTES, &wbar[st ^ 1], &tmF, &tmA, b0 + (long long)(k + NW) * NPW, adjc);
    stage_wait(&wbar[st], (unsigned)((round >> 1) & 1));
    const bool more = pg && k + NW <= k_hi;  // (warp-uniform) prefetch the gathers of the item just requested
    stage_item<L, S, MRT, ADJC>(g, p, fB, rho, wallrec, lo, hi, b0 + (long long)k * NPW, wmem + (size_t)st * G::STAGE_BYTES, m, j, lane_ok, nbr_all,
                                more ? wmem + (size_t)(st ^ 1) * G::STAGE_BYTES : nullptr, &wbar[st ^ 1], (unsigned)(((round + 1) >> 1) & 1),
                                b0 + (long long)(k + NW) * NPW);
  }
}

// Builds the compressed adjacency records (AdjcGeom) of the items [item0, item0 + nitems) from the full table and the mask array: one thread
// per item.  Only positions in [own0, own1) have table entries; the other slots of an item at the edge of the owned range are never read.
template <class L, int S>
__global__ void k_build_adjc(Grid g, const uint32_t *__restrict__ nbr_all, const uint32_t *__restrict__ lmask, unsigned char *__restrict__ adjc,
                             long long item0, long long nitems) {
  using A = AdjcGeom<L, S>;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nitems) return;
  const long long item = item0 + t, p0 = item * A::ITEM;
  unsigned char *rec = adjc + item * (long long)A::REC_BYTES;
  uint32_t *base = reinterpret_cast<uint32_t *>(rec + A::OFF_BASE);
  uint32_t *masks = reinterpret_cast<uint32_t *>(rec + A::OFF_MASK);
  unsigned char *delta = rec + A::OFF_DELTA;
  const long long lo = max(p0, g.own0), hi = min(p0 + A::ITEM, g.own1);
  for (int j = 0; j < A::ITEM; ++j) masks[j] = (p0 + j >= lo && p0 + j < hi) ? lmask[p0 + j] : 0u;
  uint32_t escape = 0;
  for (int n = 0; n < A::NB; ++n) {
    const uint32_t *row = nbr_all + (long long)n * g.fs;
    uint32_t mn = 0xffffffffu, mx = 0;
    for (long long q = lo; q < hi; ++q) {
      const uint32_t v = row[q];
      mn = min(mn, v);
      mx = max(mx, v);
    }
    if (lo >= hi) mn = mx = 0;
    if (mx - mn > 255u) escape = 1;
    base[n] = mn;
    for (int j = 0; j < A::ITEM; ++j) {
      const long long q = p0 + j;
      delta[n * A::ITEM + j] = (q >= lo && q < hi && !escape) ? (unsigned char)(row[q] - mn) : (unsigned char)0;
    }
  }
  *reinterpret_cast<uint32_t *>(rec) = escape;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_step_stage_clc: the same per-warp pipeline, but a block does not end with its LB positions: it asks the hardware block
// scheduler for the index of the next block of the grid that has not started yet (cluster launch control,
// clusterlaunchcontrol.try_cancel, sm_100) and carries on with that block's positions.  The request is asynchronous (the
// 16-byte answer lands in shared memory and completes an mbarrier), it is made one block AHEAD, and the scheduler hands the
// indices out in launch order, so
//   * the resident blocks still cover one compact, advancing window of positions (what the static short blocks are for;
//     persistent blocks with a static stride or a ticket counter in global memory lost exactly that, or paid for the atomic:
//     profiles/r2d_stage_results.txt),
//   * every item of a warp but its very first is prefetched while the previous one is collided: k_step_stage spent 12 % of
//     its warp time waiting for the copies of the first item of each block (profiles/r2z_step_stage_ncu_summary.txt).
// The grid, the block -> positions map and the arithmetic are those of k_step_stage: results are bit-identical.
// One answer slot per parity of the block ordinal; an answer is consumed by all NW warps (`empty` barrier, NW arrivals)
// before warp 0 reuses its slot.  After an answer that says "nothing left" no further request is made (PTX: undefined).
struct ClcShared {
  uint4 answer[2];
  uint64_t full[2], empty[2];
};

__device__ __forceinline__ void clc_request(uint4 *answer, uint64_t *full) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(smem_addr(full)) : "memory");
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(smem_addr(answer)),
               "r"(smem_addr(full))
               : "memory");
}

// all lanes: decode the answer in shared memory; returns the x index of the cancelled (= taken over) block or -1
__device__ __forceinline__ int clc_decode(const uint4 *answer) {
  unsigned ok, x;
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b128 a;\n\tld.shared.b128 a, [%2];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 q, a;\n\tselp.u32 %0, 1, 0, q;\n\t"
      "mov.u32 %1, 0;\n\t@q clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %1, a;\n\t}"
      : "=r"(ok), "=r"(x)
      : "r"(smem_addr(answer))
      : "memory");
  return ok ? (int)x : -1;
}

template <class L, int S, bool MRT, int NWARPS = 4>
__global__ void __launch_bounds__(StageGeom<L, S, NWARPS>::NT, StageGeom<L, S, NWARPS>::BLOCKS_PER_SM)
    k_step_stage_clc(Grid g, Phys p, const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmA,
                     double *__restrict__ fB, const double *__restrict__ rho, const double *__restrict__ wallrec, long long first,
                     long long count, long long blk0, int LB, int pg) {
  using G = StageGeom<L, S, NWARPS>;
  constexpr int NPW = G::NPW, NW = G::NW;
  extern __shared__ __align__(1024) unsigned char stage_mem[];  // [NW][2][STAGE_BYTES]
  __shared__ __align__(8) uint64_t bars[NW][2];
  __shared__ __align__(16) ClcShared clc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long last = first + count;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&clc.full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&clc.empty[s])), "r"(NW));
    }
  }
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bars[warp][0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bars[warp][1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // the only block-wide barrier of the kernel
  if (threadIdx.x == 0) clc_request(&clc.answer[1], &clc.full[1]);  // block ordinal 1 of this resident block
  unsigned char *wmem = stage_mem + (size_t)warp * G::WARP_BYTES;
  uint64_t *wbar = bars[warp];
  int m = lane / NPW;
  const int j = lane - m * NPW;
  bool lane_ok = true;
  if (m >= S) {
    m = S - 1;
    lane_ok = false;
  }
  // issue cursor: block ordinal u of this resident block, its first position ib0, this warp's next item index ik in it
  unsigned u = 0;
  long long ib0 = (blk0 + blockIdx.x) * LB;
  int ik = warp;
  // the next item of this warp (its first position) or -1 when the grid is used up
  auto next_item = [&]() -> long long {
    for (;;) {
      while (ik * NPW < LB) {
        const long long w0 = ib0 + (long long)ik * NPW;
        ik += NW;
        if (w0 + NPW > first && w0 < last) return w0;
      }
      // this block's positions are handed out: take over the next block of the grid
      const unsigned un = u + 1, s = un & 1u, use = (un - 1) >> 1;  // slot s is used for the use-th time
      stage_wait(&clc.full[s], use & 1u);
      const int b = clc_decode(&clc.answer[s]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the answer was read before the slot is written again
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&clc.empty[s])) : "memory");
      if (b < 0) return -1;
      u = un;
      ib0 = (blk0 + b) * LB;
      ik = warp;
      if (warp == 0) {  // ask for the block after that, one block ahead
        const unsigned ur = un + 1, sr = ur & 1u, user = (ur - 1) >> 1;
        if (user > 0) stage_wait(&clc.empty[sr], (user - 1) & 1u);
        if (lane == 0) clc_request(&clc.answer[sr], &clc.full[sr]);
        __syncwarp();
      }
    }
  };
  long long w_next = next_item();
  if (w_next < 0) return;
  stage_issue<L, S>(wmem, &wbar[0], &tmF, &tmA, w_next);
  for (unsigned cnt = 0;; ++cnt) {
    const long long w_cur = w_next;
    const unsigned st = cnt & 1u;
    w_next = next_item();
    // the other stage was read to the end in the previous round: refill it with this warp's next item
    if (w_next >= 0) stage_issue<L, S>(wmem + (size_t)(st ^ 1u) * G::STAGE_BYTES, &wbar[st ^ 1u], &tmF, &tmA, w_next);
    stage_wait(&wbar[st], (cnt >> 1) & 1u);
    stage_item<L, S, MRT>(g, p, fB, rho, wallrec, max(w_cur, first), min(w_cur + NPW, last), w_cur, wmem + (size_t)st * G::STAGE_BYTES, m, j,
                          lane_ok, nullptr, (pg && w_next >= 0) ? wmem + (size_t)(st ^ 1u) * G::STAGE_BYTES : nullptr, &wbar[st ^ 1u],
                          ((cnt + 1) >> 1) & 1u, w_next);
    if (w_next < 0) break;
  }
}

}  // namespace txg

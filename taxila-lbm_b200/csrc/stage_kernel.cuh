// stage_kernel.cuh -- K2 of the order-4 step with its streamed inputs staged in shared memory by bulk copies (TMA),
// one double-buffered pipeline PER WARP.
//
// k_step_fused is bound by how many bytes an SM keeps in flight, not by DRAM or the fp64 pipe (profiles/
// r1k_step_fused_ncu_summary.txt: DRAM 56 %, fp64 37 %, issue 42 %, 57 % of the warp time long-scoreboard): 126 registers
// leave 16 warps per SM, and a warp holds its 19 population loads plus the adjacency row only while it waits for them --
// during the ~1300 fp64 issue cycles of the collision it has nothing in flight.  Here every row a warp streams for its
// NPW positions -- S * Q population rows, Q - 1 adjacency rows and the mask row -- is ONE contiguous, 16-byte aligned run,
// so the warp fetches its NEXT item with S * Q + Q bulk copies (cp.async.bulk.shared.global, completion counted on the
// warp's own mbarrier) into the second stage of its double buffer while it collides the current one out of the first: the
// copies need no registers and stay in flight for the whole collision.  No block-wide barrier: a first version with one
// pipeline per 256-thread block (profiles/r2d_stage_results.txt) lost 18 % of its warp time at __syncthreads and ran
// slower than k_step_fused, because two blocks per SM are two independent instruction streams where the fused kernel has 16.
// Blocks are aligned on absolute multiples of LB positions and a warp's items on multiples of NPW, so every copy is a
// whole 128-byte line (S = 2).  What remains on the demand path: the 18 density gathers and the wall record (issued as
// soon as the adjacency row is read from shared memory: one exposed round trip, mostly L2 hits) and the scattered push
// stores.  Same arithmetic in the same order as k_step_fused: results are bit-identical.
// Replaces the same reference procedures as k_step_fused (fused_kernel.cuh).
#pragma once
#include "fused_kernel.cuh"

namespace txg {

template <class L, int S>
struct StageGeom {
  static constexpr int NT = 128, NW = NT / 32;            // threads, warps per block
  static constexpr int NPW = Lanes<S>::NPW;               // positions per warp item
  // staged run per row: starts on a multiple of 4 entries (16-byte aligned u32 runs) at or below the item's first position
  static constexpr int ITEM = NPW % 4 == 0 ? NPW : ((NPW + 6) & ~3);
  static constexpr int NF = S * L::Q, NA = L::Q - 1;      // population rows, adjacency rows
  static constexpr int F_BYTES = NF * ITEM * 8, A_BYTES = NA * ITEM * 4, M_BYTES = ITEM * 4;
  static constexpr int STAGE_BYTES = F_BYTES + A_BYTES + M_BYTES;  // a multiple of 16
  static constexpr int WARP_BYTES = 2 * STAGE_BYTES;
  static constexpr int SMEM_BYTES = NW * WARP_BYTES;
  static constexpr int BLOCKS_PER_SM = S == 2 ? 4 : 3;  // (shared memory: 48.6 KB per block for S = 2, 58 / 68 KB for S = 1 / 3)
  static_assert(STAGE_BYTES % 16 == 0, "bulk copies want 16-byte multiples");
};

// all lanes of a warp: fetch the rows of positions [p0, p0 + ITEM) into the warp's stage at `dst` (p0 a multiple of 4)
TXG_HD long long stage_start(long long w0) { return w0 & ~3ll; }
template <class L, int S>
__device__ __forceinline__ void stage_issue(unsigned char *dst, uint64_t *bar, long long fs, const double *__restrict__ fA,
                                            const uint32_t *__restrict__ nbr_all, const uint32_t *__restrict__ lmask, long long p0) {
  using G = StageGeom<L, S>;
  const int lane = threadIdx.x & 31;
  // the stage was read by plain loads: order them before the asynchronous writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((unsigned)G::STAGE_BYTES) : "memory");
  __syncwarp();
#pragma unroll
  for (int r0 = 0; r0 < G::NF + G::NA + 1; r0 += 32) {
    const int row = r0 + lane;
    if (row < G::NF + G::NA + 1) {
      const void *src;
      unsigned char *d;
      unsigned bytes;
      if (row < G::NF) {
        src = fA + (long long)row * fs + p0;
        d = dst + row * (G::ITEM * 8);
        bytes = G::ITEM * 8;
      } else if (row < G::NF + G::NA) {
        src = nbr_all + (long long)(row - G::NF) * fs + p0;
        d = dst + G::F_BYTES + (row - G::NF) * (G::ITEM * 4);
        bytes = G::ITEM * 4;
      } else {
        src = lmask + p0;
        d = dst + G::F_BYTES + G::A_BYTES;
        bytes = G::ITEM * 4;
      }
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(d)), "l"(src),
                   "r"(bytes), "r"(smem_addr(bar))
                   : "memory");
    }
  }
}

__device__ __forceinline__ void stage_wait(uint64_t *bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_addr(bar)), "r"(parity)
                 : "memory");
}

// One launch covers the positions [first, first + count); block b owns the absolute positions [(blk0 + b) * LB, + LB) of it
// (LB a multiple of NW * NPW; the host passes blk0 = first / LB); inside a block, item k = positions [k * NPW, + NPW) of the
// block goes to warp k % NW, so that the block's warps walk one contiguous run of positions together.
template <class L, int S, bool MRT>
__global__ void __launch_bounds__(StageGeom<L, S>::NT, StageGeom<L, S>::BLOCKS_PER_SM)
    k_step_stage(Grid g, Phys p, const double *__restrict__ fA, double *__restrict__ fB, const double *__restrict__ rho,
                 const uint32_t *__restrict__ lmask, const uint32_t *__restrict__ nbr_all, const double *__restrict__ wallrec,
                 long long first, long long count, long long blk0, int LB) {
  using G = StageGeom<L, S>;
  constexpr int Q = L::Q, D = L::D, ISO = 4, NPW = G::NPW, NW = G::NW, ITEM = G::ITEM;
  extern __shared__ __align__(128) unsigned char stage_mem[];  // [NW][2][STAGE_BYTES]
  __shared__ __align__(8) uint64_t bars[NW][2];
  const long long last = first + count;               // one past the last position of the launch
  const long long b0 = (blk0 + blockIdx.x) * LB;      // absolute positions of this block: [b0, b0 + LB)
  const long long lo = max(b0, first), hi = min(b0 + LB, last);
  if (lo >= hi) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  stage_issue<L, S>(wmem, &wbar[0], g.fs, fA, nbr_all, lmask, stage_start(b0 + (long long)k_lo * NPW));
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
      stage_issue<L, S>(wmem + (size_t)(st ^ 1) * G::STAGE_BYTES, &wbar[st ^ 1], g.fs, fA, nbr_all, lmask,
                        stage_start(b0 + (long long)(k + NW) * NPW));
    const long long w0 = b0 + (long long)k * NPW;       // first position of the item
    Item it;
    it.m = m;
    it.j = j;
    long long pos = w0 + j;
    it.active = lane_ok && pos >= lo && pos < hi;
    pos = min(max(pos, max(lo, w0)), min(hi, w0 + NPW) - 1);  // replayed lanes: a valid position of this item
    it.pos = pos;
    const unsigned char *sm = wmem + (size_t)st * G::STAGE_BYTES;
    stage_wait(&wbar[st], (unsigned)((round >> 1) & 1));
    const int i = (int)(pos - stage_start(w0));
    double f[Q];
    unsigned npos[Q];
    const uint32_t *sa = reinterpret_cast<const uint32_t *>(sm + G::F_BYTES) + i;
    npos[0] = (unsigned)pos;
#pragma unroll
    for (int n = 1; n < Q; ++n) npos[n] = sa[(n - 1) * ITEM];
    const uint32_t mask = reinterpret_cast<const uint32_t *>(sm + G::F_BYTES + G::A_BYTES)[i];
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
}

}  // namespace txg

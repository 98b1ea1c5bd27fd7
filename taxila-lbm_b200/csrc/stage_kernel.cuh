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

template <class L, int S>
struct StageGeom {
  static constexpr int NT = 128, NW = NT / 32;            // threads, warps per block
  static constexpr int NPW = Lanes<S>::NPW;               // positions per warp item
  static constexpr int ITEM = (NPW + 3) & ~3;             // staged run per row (inner box: a multiple of 16 bytes for u32 rows too)
  static constexpr int NF = S * L::Q, NA = L::Q;          // population rows; adjacency rows + the mask row
  static constexpr int F_BYTES = NF * ITEM * 8, A_BYTES = NA * ITEM * 4;
  static constexpr int A_OFF = (F_BYTES + 127) & ~127;    // tensor copies land on 128-byte boundaries
  static constexpr int STAGE_BYTES = (A_OFF + A_BYTES + 127) & ~127;
  static constexpr int WARP_BYTES = 2 * STAGE_BYTES;
  static constexpr int SMEM_BYTES = NW * WARP_BYTES;
  static constexpr int BLOCKS_PER_SM = S == 2 ? 4 : 3;  // (shared memory: 48 KB per block for S = 2)
};

// One lane of a warp: fetch the item of positions [p0, p0 + ITEM) -- the box {ITEM positions} x {all S * Q population
// rows} of the population tensor and the box {ITEM} x {Q - 1 adjacency rows + mask row} of the adjacency tensor -- into the
// warp's stage at `dst` with TWO descriptor-based tensor copies (cp.async.bulk.tensor.2d, tile mode): the copy engine
// walks the rows itself.  (A first per-warp form issued one cp.async.bulk per row, 57 per item, and lost 17 % of its time
// at the copy instruction: profiles/r2d_stage_results.txt.)  Positions past the end of a row are zero-filled.
template <class L, int S>
__device__ __forceinline__ void stage_issue(unsigned char *dst, uint64_t *bar, const CUtensorMap *tmF, const CUtensorMap *tmA,
                                            long long p0) {
  using G = StageGeom<L, S>;
  // the stage was read by plain loads: order them before the asynchronous writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((unsigned)(G::F_BYTES + G::A_BYTES)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_addr(dst)),
                 "l"(tmF), "r"((int)p0), "r"(0), "r"(smem_addr(bar))
                 : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_addr(dst + G::A_OFF)),
                 "l"(tmA), "r"((int)p0), "r"(0), "r"(smem_addr(bar))
                 : "memory");
  }
}

// one lane: ask L2 for the two boxes of the item at p0 (descriptor-based prefetch, no shared memory, no completion)
__device__ __forceinline__ void stage_prefetch_l2(const CUtensorMap *tmF, const CUtensorMap *tmA, long long p0) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmF), "r"((int)p0), "r"(0) : "memory");
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmA), "r"((int)p0), "r"(0) : "memory");
}

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

// One launch covers the positions [first, first + count); block b owns the absolute positions [(blk0 + b) * LB, + LB) of it
// (LB a multiple of NW * NPW; the host passes blk0 = first / LB); inside a block, item k = positions [k * NPW, + NPW) of the
// block goes to warp k % NW, so that the block's warps walk one contiguous run of positions together.  SHORT blocks (two
// rounds: LB = 2 * NW * NPW = 128 positions for S = 2) measured fastest: the resident blocks then cover one compact,
// advancing window of positions like k_step_fused does.  Longer blocks, and persistent warps drawing their items from a
// ticket counter (one at a time: 11 % of the warp time at the atomic; four at a time: the resident warps spread over
// several planes and the density gathers start missing L2), were slower -- profiles/r2d_stage_results.txt.
template <class L, int S, bool MRT>
__global__ void __launch_bounds__(StageGeom<L, S>::NT, StageGeom<L, S>::BLOCKS_PER_SM)
    k_step_stage(Grid g, Phys p, const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmA,
                 double *__restrict__ fB, const double *__restrict__ rho, const double *__restrict__ wallrec, long long first,
                 long long count, long long blk0, int LB, int pf_blocks, const double *__restrict__ fA_rows,
                 const uint32_t *__restrict__ adj_rows) {
  using G = StageGeom<L, S>;
  constexpr int Q = L::Q, D = L::D, ISO = 4, NPW = G::NPW, NW = G::NW, ITEM = G::ITEM;
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
  stage_issue<L, S>(wmem, &wbar[0], &tmF, &tmA, b0 + (long long)k_lo * NPW);
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
      stage_issue<L, S>(wmem + (size_t)(st ^ 1) * G::STAGE_BYTES, &wbar[st ^ 1], &tmF, &tmA, b0 + (long long)(k + NW) * NPW);
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
    const int i = (int)(pos - w0);
    double f[Q];
    unsigned npos[Q];
    const uint32_t *sa = reinterpret_cast<const uint32_t *>(sm + G::A_OFF) + i;
    npos[0] = (unsigned)pos;
#pragma unroll
    for (int n = 1; n < Q; ++n) npos[n] = sa[(n - 1) * ITEM];
    const uint32_t mask = sa[(Q - 1) * ITEM];
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

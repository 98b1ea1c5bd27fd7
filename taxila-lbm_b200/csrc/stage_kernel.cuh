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

// One launch covers the positions [first, first + count), cut into items of NPW positions on absolute multiples of NPW.
// The grid is persistent (at most BLOCKS_PER_SM blocks per SM); every WARP draws its items from a ticket counter, so that
// at any moment the resident warps work one contiguous, advancing window of positions -- the access pattern of k_step_fused,
// whose L2 / DRAM page locality a static assignment of long runs to blocks loses (profiles/r2d_stage_results.txt: the more
// rounds per block, the slower) -- and every warp always has its NEXT item in flight.  A warp holds three tickets: the
// item it collides, the item whose copies are in flight, and the ticket it has asked for (the atomic's round trip hides
// behind a whole collision).  *ticket must be 0 at launch (the host clears it on the stream).
template <class L, int S, bool MRT>
__global__ void __launch_bounds__(StageGeom<L, S>::NT, StageGeom<L, S>::BLOCKS_PER_SM)
    k_step_stage(Grid g, Phys p, const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmA,
                 double *__restrict__ fB, const double *__restrict__ rho, const double *__restrict__ wallrec, long long first,
                 long long count, unsigned *__restrict__ ticket) {
  using G = StageGeom<L, S>;
  constexpr int Q = L::Q, D = L::D, ISO = 4, NPW = G::NPW, ITEM = G::ITEM;
  extern __shared__ __align__(1024) unsigned char stage_mem[];  // [NW][2][STAGE_BYTES]
  __shared__ __align__(8) uint64_t bars[G::NW][2];
  const long long lo = first, hi = first + count;     // positions of the launch: [lo, hi)
  const long long a0 = lo / NPW;                      // first item (absolute index: item a = positions [a * NPW, + NPW))
  const unsigned nitems = (unsigned)((hi - 1) / NPW - a0 + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *wmem = stage_mem + (size_t)warp * G::WARP_BYTES;
  uint64_t *wbar = bars[warp];
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&wbar[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&wbar[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // tickets: t_cur is collided, t_nxt is in flight, t_req is lane 0's pending atomic (broadcast when it becomes t_nxt)
  unsigned t_cur = 0u, t_nxt = 0u, t_req = 0u;
  if (lane == 0) {
    t_cur = atomicAdd(ticket, 2u);  // two consecutive items to start with
    t_req = atomicAdd(ticket, 1u);
  }
  t_cur = __shfl_sync(0xffffffffu, t_cur, 0);
  t_nxt = t_cur + 1u;
  if (t_cur >= nitems) return;
  stage_issue<L, S>(wmem, &wbar[0], &tmF, &tmA, (a0 + t_cur) * NPW);
  // lane -> (component, node slot) like item_of_lane
  int m = lane / NPW;
  const int j = lane - m * NPW;
  bool lane_ok = true;
  if (m >= S) {
    m = S - 1;
    lane_ok = false;
  }
  for (unsigned round = 0; t_cur < nitems; ++round) {
    const int st = round & 1;
    // the other stage was read to the end in the previous round: refill it with this warp's next item
    if (t_nxt < nitems) stage_issue<L, S>(wmem + (size_t)(st ^ 1) * G::STAGE_BYTES, &wbar[st ^ 1], &tmF, &tmA, (a0 + t_nxt) * NPW);
    const long long w0 = (a0 + t_cur) * NPW;            // first position of the item
    // the tickets move up; lane 0 asks for one more (its value is not needed before the next round)
    t_cur = t_nxt;
    t_nxt = __shfl_sync(0xffffffffu, t_req, 0);
    if (lane == 0) t_req = atomicAdd(ticket, 1u);
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

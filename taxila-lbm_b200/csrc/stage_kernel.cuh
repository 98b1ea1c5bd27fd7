// stage_kernel.cuh -- K2 of the order-4 step with its streamed inputs staged in shared memory by bulk copies (TMA).
//
// k_step_fused is bound by how many bytes an SM keeps in flight, not by DRAM or the fp64 pipe (profiles/
// r1k_step_fused_ncu_summary.txt: DRAM 56 %, fp64 37 %, issue 42 %, 57 % of the warp time long-scoreboard): 126 registers
// leave 16 warps per SM, and a warp holds its 19 population loads plus the adjacency row only while it waits for them --
// during the ~1300 fp64 issue cycles of the collision it has nothing in flight.  Here a block of NT threads owns LB
// consecutive positions and walks them in chunks of CH = (NT / 32) * NPW positions.  Every row the chunk streams --
// S * Q population rows, Q - 1 adjacency rows and the mask row -- is ONE contiguous, 16-byte aligned run of CH entries, so
// one elected warp fetches chunk k + 1 with S * Q + Q bulk copies (cp.async.bulk.shared.global, completion counted on an
// mbarrier) into the second stage of a double buffer while all warps collide chunk k out of the first: the copies need no
// registers and stay in flight for the whole collision (two blocks per SM: ~97 KB in flight per SM against ~30 KB).
// Blocks are aligned on absolute multiples of LB positions, so every copy is a whole number of 128-byte lines.
// What remains on the demand path: the 18 density gathers and the wall record (issued as soon as the adjacency row is
// read from shared memory: one exposed round trip, mostly L2 hits) and the scattered push stores.
// Same arithmetic in the same order as k_step_fused: results are bit-identical.
// Replaces the same reference procedures as k_step_fused (fused_kernel.cuh).
#pragma once
#include "fused_kernel.cuh"

namespace txg {

template <class L, int S>
struct StageGeom {
  static constexpr int NT = S == 1 ? 128 : 256;           // threads per block
  static constexpr int NPW = Lanes<S>::NPW;
  static constexpr int CH = (NT / 32) * NPW;              // positions per chunk (a multiple of 4: 16-byte aligned u32 runs)
  static constexpr int NF = S * L::Q, NA = L::Q - 1;      // population rows, adjacency rows
  static constexpr int F_BYTES = NF * CH * 8, A_BYTES = NA * CH * 4, M_BYTES = CH * 4;
  static constexpr int STAGE_BYTES = F_BYTES + A_BYTES + M_BYTES;  // a multiple of 16
  static constexpr int SMEM_BYTES = 2 * STAGE_BYTES;
  static_assert(CH % 4 == 0 && STAGE_BYTES % 16 == 0, "bulk copies want 16-byte multiples");
};

// warp 0: fetch the rows of positions [p0, p0 + CH) into the stage at `dst`
template <class L, int S>
__device__ __forceinline__ void stage_issue(unsigned char *dst, uint64_t *bar, const Grid &g, const double *__restrict__ fA,
                                            const uint32_t *__restrict__ nbr_all, const uint32_t *__restrict__ lmask, long long p0) {
  using G = StageGeom<L, S>;
  const int lane = threadIdx.x & 31;
  if (lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((unsigned)G::STAGE_BYTES) : "memory");
  __syncwarp();
  for (int row = lane; row < G::NF + G::NA + 1; row += 32) {
    const void *src;
    unsigned char *d;
    unsigned bytes;
    if (row < G::NF) {
      src = fA + (long long)row * g.fs + p0;
      d = dst + (size_t)row * G::CH * 8;
      bytes = G::CH * 8;
    } else if (row < G::NF + G::NA) {
      src = nbr_all + (long long)(row - G::NF) * g.fs + p0;
      d = dst + G::F_BYTES + (size_t)(row - G::NF) * G::CH * 4;
      bytes = G::CH * 4;
    } else {
      src = lmask + p0;
      d = dst + G::F_BYTES + G::A_BYTES;
      bytes = G::CH * 4;
    }
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(d)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
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
// (LB a multiple of CH; the host passes blk0 = first / LB).
template <class L, int S, bool MRT>
__global__ void __launch_bounds__(StageGeom<L, S>::NT, 2)
    k_step_stage(Grid g, Phys p, const double *__restrict__ fA, double *__restrict__ fB, const double *__restrict__ rho,
                 const uint32_t *__restrict__ lmask, const uint32_t *__restrict__ nbr_all, const double *__restrict__ wallrec,
                 long long first, long long count, long long blk0, int LB) {
  using G = StageGeom<L, S>;
  constexpr int Q = L::Q, D = L::D, ISO = 4, NPW = G::NPW, CH = G::CH;
  extern __shared__ __align__(128) unsigned char stage_mem[];  // [2][STAGE_BYTES]
  __shared__ __align__(8) uint64_t bars[2];
  const long long last = first + count;               // one past the last position of the launch
  const long long b0 = (blk0 + blockIdx.x) * LB;      // absolute positions of this block: [b0, b1)
  const long long lo = max(b0, first), hi = min(b0 + LB, last);
  if (lo >= hi) return;
  // chunks of the block that hold positions of the launch
  const int k0 = (int)((lo - b0) / CH), k1 = (int)((hi - 1 - b0) / CH);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bars[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bars[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) stage_issue<L, S>(stage_mem, &bars[0], g, fA, nbr_all, lmask, b0 + (long long)k0 * CH);
  // lane -> (component, node slot) like item_of_lane
  int m = lane / NPW;
  const int j = lane - m * NPW;
  bool lane_ok = true;
  if (m >= S) {
    m = S - 1;
    lane_ok = false;
  }
  for (int k = k0; k <= k1; ++k) {
    const int st = (k - k0) & 1;
    // the other stage was read to the end in the previous round (barrier below): refill it with the next chunk
    if (warp == 0 && k < k1)
      stage_issue<L, S>(stage_mem + (size_t)(st ^ 1) * G::STAGE_BYTES, &bars[st ^ 1], g, fA, nbr_all, lmask, b0 + (long long)(k + 1) * CH);
    const long long c0 = b0 + (long long)k * CH;       // first position of the chunk
    // this warp's positions [w0, w0 + NPW) clipped to the launch; a warp wholly outside skips the arithmetic
    const long long w0 = c0 + (long long)warp * NPW;
    const bool warp_on = w0 < hi && w0 + NPW > lo;
    Item it;
    it.m = m;
    it.j = j;
    long long pos = w0 + j;
    it.active = lane_ok && pos >= lo && pos < hi;
    pos = min(max(pos, max(lo, w0)), min(hi, w0 + NPW) - 1);  // replayed lanes: a valid position of this warp
    it.pos = pos;
    const unsigned char *sm = stage_mem + (size_t)st * G::STAGE_BYTES;
    stage_wait(&bars[st], (unsigned)(((k - k0) >> 1) & 1));
    double f[Q];
    unsigned npos[Q];
    uint32_t mask = 0u;
    if (warp_on) {
      const int i = (int)(pos - c0);
      const double *sf = reinterpret_cast<const double *>(sm) + (size_t)m * Q * CH + i;
#pragma unroll
      for (int n = 0; n < Q; ++n) f[n] = sf[n * CH];
      const uint32_t *sa = reinterpret_cast<const uint32_t *>(sm + G::F_BYTES) + i;
      npos[0] = (unsigned)pos;
#pragma unroll
      for (int n = 1; n < Q; ++n) npos[n] = sa[(n - 1) * CH];
      mask = reinterpret_cast<const uint32_t *>(sm + G::F_BYTES + G::A_BYTES)[i];
    }
    __syncthreads();  // every thread has its operands in registers: the stage may be refilled in the next round
    if (!warp_on) continue;
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

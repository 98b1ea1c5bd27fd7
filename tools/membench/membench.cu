// membench.cu -- attainable HBM bandwidth for the access structure of the collide kernel:
// R row-streams read + R row-streams written, rows `fs` doubles apart, positions consecutive.
//   A: plain loads/stores, one lane per (position, half of the rows), high occupancy
//   B: like A but capped at 16 warps/SM through dynamic shared memory (the collide kernel's occupancy)
//   C: persistent blocks, bulk async copies (TMA) into shared memory two chunks ahead, plain stores
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int Q = 19, S = 2, R = S * Q;

__global__ void __launch_bounds__(128) k_copy(const double *__restrict__ a, double *__restrict__ b, long long fs, long long n, int shift, int lshift = 0) {
  extern __shared__ unsigned char dummy[];
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, m = lane >> 4, j = lane & 15;
  const long long pos = w * 16 + j;
  if (pos >= n) return;
  double f[Q];
  const long long lp = pos + lshift < n ? pos + lshift : pos;
#pragma unroll
  for (int q = 0; q < Q; ++q) f[q] = __ldg(a + (long long)(m * Q + q) * fs + lp);
  const long long op = pos + shift < n ? pos + shift : pos;
#pragma unroll
  for (int q = 0; q < Q; ++q) b[(long long)(m * Q + q) * fs + op] = f[q] + 1.0;
}

// stores shifted by `shift` but re-aligned through shared memory: the block's PB values of each row
// target [base+shift, base+shift+PB); lanes map to aligned target addresses, so every sector is
// written by one request except the two at the ends of the block's range
template <int PB>
__global__ void __launch_bounds__(128) k_copy_staged(const double *__restrict__ a, double *__restrict__ b, long long fs, long long n, int shift) {
  __shared__ double st[R][PB + 8];
  const long long base = (long long)blockIdx.x * PB;
  if (base + PB > n) return;
  for (int idx = threadIdx.x; idx < R * PB; idx += 128) {
    const int row = idx / PB, j = idx - row * PB;
    st[row][j] = __ldg(a + (long long)row * fs + base + j);
  }
  __syncthreads();
  // targets t in [base+shift, base+shift+PB): aligned walk starting at (base+shift) & ~15
  const long long t0 = (base + shift) & ~15ll;
  constexpr int SPAN = PB + 16;
  for (int idx = threadIdx.x; idx < R * SPAN; idx += 128) {
    const int row = idx / SPAN, k = idx - row * SPAN;
    const long long t = t0 + k;
    const long long j = t - (base + shift);
    if (j >= 0 && j < PB && t < n) b[(long long)row * fs + t] = st[row][j] + 1.0;
  }
}

// push-like copy: stores shifted per direction by the neighbour offsets of a 512^2 porous plane, and
// optionally one fp64 reduction (RED.ADD.F64) per pushed value into rho[target] (the "density by atomics" idea)
__global__ void __launch_bounds__(128) k_push(const double *__restrict__ a, double *__restrict__ b, double *__restrict__ rho, long long fs, long long n, int atomics) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, m = lane >> 4, j = lane & 15;
  const long long pos = w * 16 + j;
  if (pos >= n) return;
  const int row = 231, plane = 117965;
  const int off[Q] = {0, 1, row, -1, -row, plane, -plane, row + 1, row - 1, -row - 1, -row + 1, plane + 1, plane - 1, -plane - 1, -plane + 1, plane + row, plane - row, -plane - row, -plane + row};
  double f[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) f[q] = __ldg(a + (long long)(m * Q + q) * fs + pos);
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    long long t = pos + off[q];
    if (t < 0 || t >= n) t = pos;
    b[(long long)(m * Q + q) * fs + t] = f[q] + 1.0;
    if (atomics) atomicAdd(rho + (long long)m * fs + t, f[q]);
  }
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t *bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int PB, int NST>
struct Sm { double f[NST][R][PB]; uint64_t full[NST]; uint64_t empty[NST]; };

template <int PB, int NST>
__global__ void __launch_bounds__(128) k_tma(const double *__restrict__ a, double *__restrict__ b, long long fs, long long n, int pfd = 0) {
  extern __shared__ __align__(128) unsigned char raw[];
  Sm<PB, NST> &sm = *reinterpret_cast<Sm<PB, NST> *>(raw);
  const long long nchunks = n / PB;
  const long long mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sm.full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sm.empty[s])), "r"(4));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, long long k) {
    const long long pos = ((long long)blockIdx.x + k * gridDim.x) * PB;
    if (pfd > 0 && k + pfd < mine) {  // L2 bulk prefetch of the rows of the chunk pfd iterations ahead
      const long long pp = ((long long)blockIdx.x + (k + pfd) * gridDim.x) * PB;
      for (int row = lane; row < R; row += 32)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a + (long long)row * fs + pp), "r"((unsigned)(PB * 8)) : "memory");
    }
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.full[s])), "r"((unsigned)(R * PB * 8)) : "memory");
    __syncwarp();
    for (int row = lane; row < R; row += 32)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&sm.f[s][row][0])),
                   "l"(a + (long long)row * fs + pos), "r"((unsigned)(PB * 8)), "r"(smem_u32(&sm.full[s])) : "memory");
  };
  if (warp == 0)
    for (int s = 0; s < NST; ++s) if (s < mine) issue(s, s);
  constexpr int PPW = PB / 4;  // positions per warp
  for (long long k = 0; k < mine; ++k) {
    const int s = (int)(k % NST);
    const unsigned parity = (unsigned)((k / NST) & 1);
    while (!try_wait(&sm.full[s], parity)) {}
    const long long base = ((long long)blockIdx.x + k * gridDim.x) * PB;
    // each warp: PPW positions x 2 components; lane -> (m, j) with j stepping by 16
    for (int j0 = 0; j0 < PPW; j0 += 16) {
      const int m = lane >> 4, j = warp * PPW + j0 + (lane & 15);
      double f[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) f[q] = sm.f[s][m * Q + q][j];
      if (j0 + 16 >= PPW) {
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.empty[s])) : "memory");
        if (warp == 0 && k + NST < mine) {
          while (!try_wait(&sm.empty[s], parity)) {}
          issue(s, k + NST);
        }
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) b[(long long)(m * Q + q) * fs + base + j] = f[q] + 1.0;
    }
  }
}

// C2: the TMA-fed copy with the other ingredients of the collide kernel switched on one by one:
//   work  > 0 : a dependent fp64 FMA chain of `work` steps per value group (register-only compute)
//   push  = 1 : stores go to the neighbour offsets of a 512^2 porous plane (scattered like the push)
//   gath  = 1 : 18 gathers per lane from a density array at the same offsets
template <int PB, int NST, int NW>
__global__ void __launch_bounds__(32 * NW) k_tma2(const double *__restrict__ a, double *__restrict__ b, const double *__restrict__ rho,
                                                  long long fs, long long n, int work, int push, int gath) {
  extern __shared__ __align__(128) unsigned char raw[];
  Sm<PB, NST> &sm = *reinterpret_cast<Sm<PB, NST> *>(raw);
  const long long nchunks = n / PB;
  const long long mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sm.full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sm.empty[s])), "r"(NW));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, long long k) {
    const long long pos = ((long long)blockIdx.x + k * gridDim.x) * PB;
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.full[s])), "r"((unsigned)(R * PB * 8)) : "memory");
    __syncwarp();
    for (int row = lane; row < R; row += 32)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&sm.f[s][row][0])),
                   "l"(a + (long long)row * fs + pos), "r"((unsigned)(PB * 8)), "r"(smem_u32(&sm.full[s])) : "memory");
  };
  if (warp == 0)
    for (int s = 0; s < NST; ++s) if (s < mine) issue(s, s);
  static_assert(PB == 16 * NW, "one item per warp");
  const int rowo = 231, plane = 117965;
  const int off[Q] = {0, 1, rowo, -1, -rowo, plane, -plane, rowo + 1, rowo - 1, -rowo - 1, -rowo + 1, plane + 1, plane - 1, -plane - 1, -plane + 1, plane + rowo, plane - rowo, -plane - rowo, -plane + rowo};
  for (long long k = 0; k < mine; ++k) {
    const int s = (int)(k % NST);
    const unsigned parity = (unsigned)((k / NST) & 1);
    while (!try_wait(&sm.full[s], parity)) {}
    const long long base = ((long long)blockIdx.x + k * gridDim.x) * PB;
    const int m = lane >> 4, j = warp * 16 + (lane & 15);
    const long long pos = base + j;
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) f[q] = sm.f[s][m * Q + q][j];
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.empty[s])) : "memory");
    if (warp == 0 && k + NST < mine) {
      while (!try_wait(&sm.empty[s], parity)) {}
      issue(s, k + NST);
    }
    double g = 0.;
    if (gath) {
#pragma unroll
      for (int q = 1; q < Q; ++q) {
        // gath = 2: only the 8 row centres (the x+-1 values would come from warp shuffles)
        if (gath == 2 && !(q == 2 || q == 4 || q == 5 || q == 6 || q >= 15)) continue;
        long long t = pos + off[q];
        if (t < 0 || t >= n) t = pos;
        g += __ldg(rho + (long long)m * fs + t);
      }
    }
    double acc = g;
    for (int it = 0; it < work; ++it) {
#pragma unroll
      for (int q = 0; q < Q; ++q) acc = fma(acc, 1.0000001, f[q]);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      long long t = pos;
      if (push) { t = pos + off[q]; if (t < 0 || t >= n) t = pos; }
      b[(long long)(m * Q + q) * fs + t] = f[q] + acc;
    }
  }
}

int main(int argc, char **argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 60000000ll / 256 * 256;
  const long long fs = n + 128;
  double *a, *b;
  CK(cudaMalloc(&a, sizeof(double) * R * fs));
  CK(cudaMalloc(&b, sizeof(double) * R * fs));
  CK(cudaMemset(a, 0, sizeof(double) * R * fs));
  CK(cudaMemset(b, 0, sizeof(double) * R * fs));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double gb = 2.0 * R * n * 8 / 1e9;
  auto timeit = [&](const char *name, auto launch) {
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%-48s %8.3f ms  %7.1f GB/s\n", name, ms, gb / (ms * 1e-3));
  };
  const unsigned blocks = (unsigned)((n / 16 + 3) / 4);
  timeit("A plain copy, full occupancy", [&] { k_copy<<<blocks, 128>>>(a, b, fs, n, 0); });
  timeit("A plain copy, stores shifted by 1", [&] { k_copy<<<blocks, 128>>>(a, b, fs, n, 1); });
  timeit("A plain copy, loads shifted by 1", [&] { k_copy<<<blocks, 128>>>(a, b, fs, n, 0, 1); });
  timeit("A plain copy, stores shifted by 4", [&] { k_copy<<<blocks, 128>>>(a, b, fs, n, 4); });
  timeit("A plain copy, stores shifted by 16", [&] { k_copy<<<blocks, 128>>>(a, b, fs, n, 16); });
  timeit("A plain copy, stores shifted by 2", [&] { k_copy<<<blocks, 128>>>(a, b, fs, n, 2); });
  timeit("D staged PB=64, shift 0", [&] { k_copy_staged<64><<<(unsigned)(n / 64), 128>>>(a, b, fs, n, 0); });
  timeit("D staged PB=64, shift 1", [&] { k_copy_staged<64><<<(unsigned)(n / 64), 128>>>(a, b, fs, n, 1); });
  timeit("D staged PB=128, shift 1", [&] { k_copy_staged<128><<<(unsigned)(n / 128), 128>>>(a, b, fs, n, 1); });
  {
    double *rho; CK(cudaMalloc(&rho, sizeof(double) * S * fs)); CK(cudaMemset(rho, 0, sizeof(double) * S * fs));
    timeit("E push-like copy (neighbour offsets), no atomics", [&] { k_push<<<blocks, 128>>>(a, b, rho, fs, n, 0); });
    timeit("E push-like copy + 19 RED.ADD.F64 per lane", [&] { k_push<<<blocks, 128>>>(a, b, rho, fs, n, 1); });
    cudaFree(rho);
  }
  for (int kb : {48}) {  // dynamic smem per block -> blocks per SM ~ 227/kb
    CK(cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
    char nm[64]; snprintf(nm, sizeof nm, "B plain copy, %d KB smem/block (%d blk/SM)", kb, 227 / (kb + 1));
    timeit(nm, [&] { k_copy<<<blocks, 128, kb * 1024>>>(a, b, fs, n, 0); });
  }
  {
    constexpr int PB = 64, NST = 2;
    const int smem = sizeof(Sm<PB, NST>) + 128;
    CK(cudaFuncSetAttribute(k_tma<PB, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tma<PB, NST>, 128, smem));
    char nm[64]; snprintf(nm, sizeof nm, "C TMA PB=64 2 stages (%d blk/SM, %d B smem)", nb, smem);
    timeit(nm, [&] { k_tma<PB, NST><<<148 * nb, 128, smem>>>(a, b, fs, n); });
    for (int pfd : {2, 4, 8, 16}) {
      snprintf(nm, sizeof nm, "C TMA PB=64 2 stages + L2 bulk prefetch %d ahead", pfd);
      timeit(nm, [&] { k_tma<PB, NST><<<148 * nb, 128, smem>>>(a, b, fs, n, pfd); });
    }
    for (int g : {2, 4}) {
      snprintf(nm, sizeof nm, "C TMA PB=64 2 stages, only %d blk/SM, pf 8", g);
      timeit(nm, [&] { k_tma<PB, NST><<<148 * g, 128, smem>>>(a, b, fs, n, 8); });
      snprintf(nm, sizeof nm, "C TMA PB=64 2 stages, only %d blk/SM, no pf", g);
      timeit(nm, [&] { k_tma<PB, NST><<<148 * g, 128, smem>>>(a, b, fs, n, 0); });
    }
  }
  {
    constexpr int PB = 64, NST = 3;
    const int smem = sizeof(Sm<PB, NST>) + 128;
    CK(cudaFuncSetAttribute(k_tma<PB, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tma<PB, NST>, 128, smem));
    char nm[64]; snprintf(nm, sizeof nm, "C TMA PB=64 3 stages (%d blk/SM, %d B smem)", nb, smem);
    timeit(nm, [&] { k_tma<PB, NST><<<148 * nb, 128, smem>>>(a, b, fs, n); });
  }
  {
    constexpr int PB = 128, NST = 2;
    const int smem = sizeof(Sm<PB, NST>) + 128;
    CK(cudaFuncSetAttribute(k_tma<PB, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tma<PB, NST>, 128, smem));
    char nm[64]; snprintf(nm, sizeof nm, "C TMA PB=128 2 stages (%d blk/SM, %d B smem)", nb, smem);
    timeit(nm, [&] { k_tma<PB, NST><<<148 * nb, 128, smem>>>(a, b, fs, n); });
  }
  {
    constexpr int PB = 256, NST = 2;
    const int smem = sizeof(Sm<PB, NST>) + 128;
    CK(cudaFuncSetAttribute(k_tma<PB, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tma<PB, NST>, 128, smem));
    char nm[64]; snprintf(nm, sizeof nm, "C TMA PB=256 2 stages (%d blk/SM, %d B smem)", nb, smem);
    timeit(nm, [&] { k_tma<PB, NST><<<148 * nb, 128, smem>>>(a, b, fs, n); });
  }
  {
    double *rho; CK(cudaMalloc(&rho, sizeof(double) * S * fs)); CK(cudaMemset(rho, 0, sizeof(double) * S * fs));
    auto runcfg = [&](auto kern, int nw, int smem, const char *tag) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int nb; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * nw, smem));
      for (int cfg = 0; cfg < 9; ++cfg) {
        const int work = (cfg == 1 || cfg >= 4) ? 30 : 0, push = (cfg == 2 || cfg >= 5), gath = cfg == 7 ? 2 : (cfg == 8 ? 0 : (cfg == 3 || cfg >= 6));
        if (cfg == 8) continue;
        char nm[96]; snprintf(nm, sizeof nm, "C2 %s (%d blk/SM) work=%d push=%d gath=%d", tag, nb, work, push, gath);
        timeit(nm, [&] { kern<<<148 * nb, 32 * nw, smem>>>(a, b, rho, fs, n, work, push, gath); });
      }
    };
    runcfg(k_tma2<128, 2, 8>, 8, (int)sizeof(Sm<128, 2>) + 128, "PB=128 8 warps");
    runcfg(k_tma2<64, 2, 4>, 4, (int)sizeof(Sm<64, 2>) + 128, "PB=64 4 warps");
    runcfg(k_tma2<64, 3, 4>, 4, (int)sizeof(Sm<64, 3>) + 128, "PB=64 4 warps 3 stages");
    cudaFree(rho);
  }
  return 0;
}

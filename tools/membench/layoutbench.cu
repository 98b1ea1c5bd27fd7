// layoutbench.cu -- does the row layout of the population arrays matter at the collide kernel's occupancy?
// One lane per (position, component); 19 population loads, 8 adjacency loads, 18 dependent density
// gathers, `work` rounds of fp64 FMAs, 19 pushed stores; 16 warps/SM (dynamic smem cap).
//   layout 0: f[row][fs]                (76 row streams, 483 MB apart: ~120 distinct 2 MB pages per item)
//   layout 1: f[pos/64][row][64]        (all rows of 64 positions in one 19 KB block)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o layoutbench layoutbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int Q = 19, S = 2, R = S * Q, NC = 8;
template <int LAY, int ROWS>
__device__ __forceinline__ long long idx(int row, long long pos, long long fs) {
  if (LAY == 0) return (long long)row * fs + pos;
  return ((pos >> 6) * ROWS + row) * 64 + (pos & 63);
}
__global__ void k_init_nbr(uint32_t *nb, long long fs, long long n, int lay) {
  const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const int row = 231, plane = 117965;
  const int off[NC] = {row, -row, plane, -plane, plane + row, plane - row, -plane - row, -plane + row};
  for (int k = 0; k < NC; ++k) {
    long long t = pos + off[k];
    if (t < 1 || t >= n - 1) t = pos;
    nb[lay ? idx<1, NC>(k, pos, fs) : idx<0, NC>(k, pos, fs)] = (uint32_t)t;
  }
}
// variant bits: 1 = gather addresses by arithmetic (no dependence on the adjacency loads);
//               2 = gather only the 8 centres + own node, x+-1 values by warp shuffles
//               4 = the fp64 work runs as 4 independent chains (throughput- instead of latency-bound)
template <int LAY, int VAR>
__global__ void __launch_bounds__(128) k_lbm(const double *__restrict__ a, double *__restrict__ b, const double *__restrict__ rho,
                                             const uint32_t *__restrict__ nb, long long fs, long long n, int work, int gath, int push) {
  extern __shared__ unsigned char dummy[];
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, m = lane >> 4, j = lane & 15;
  const long long pos = w * 16 + j;
  if (pos < 1 || pos >= n - 1) return;
  unsigned cen[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) cen[k] = __ldg(nb + idx<LAY, NC>(k, pos, fs));
  double f[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) f[q] = __ldg(a + idx<LAY, R>(m * Q + q, pos, fs));
  // neighbour positions: own row +-1, centres, centres +-1 (10 of the 18 derive from centres)
  unsigned t[Q];
  t[0] = (unsigned)pos; t[1] = t[0] + 1; t[2] = t[0] - 1;
#pragma unroll
  for (int k = 0; k < NC; ++k) t[3 + k] = cen[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) { t[11 + 2 * k] = cen[k] + 1; t[12 + 2 * k] = cen[k] - 1; }
  double g = 0.;
  if (gath) {
    if (VAR & 1) {
      const int row = 231, plane = 117965;
      const int off[Q] = {0, 1, -1, row, -row, plane, -plane, plane + row, plane - row, -plane - row, -plane + row,
                          row + 1, row - 1, -row + 1, -row - 1, plane + 1, plane - 1, -plane + 1, -plane - 1};
#pragma unroll
      for (int q = 1; q < Q; ++q) {
        long long tt = pos + off[q];
        if (tt < 1 || tt >= n - 1) tt = pos;
        g += __ldg(rho + (long long)m * fs + tt);
      }
    } else if (VAR & 2) {
      double c[NC + 1];
      c[0] = __ldg(rho + (long long)m * fs + t[0]);
#pragma unroll
      for (int k = 0; k < NC; ++k) c[k + 1] = __ldg(rho + (long long)m * fs + cen[k]);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        g += c[k] + __shfl_up_sync(0xffffffffu, c[k], 1) + __shfl_down_sync(0xffffffffu, c[k], 1);
      }
#pragma unroll
      for (int k = 5; k < NC + 1; ++k) g += c[k];
    } else {
#pragma unroll
      for (int q = 1; q < Q; ++q) g += __ldg(rho + (long long)m * fs + t[q]);
    }
  }
  double acc = g;
  if (VAR & 4) {
    double a0 = g, a1 = g + 1., a2 = g + 2., a3 = g + 3.;
    for (int it = 0; it < work; ++it) {
#pragma unroll
      for (int q = 0; q + 3 < Q; q += 4) {
        a0 = fma(a0, 1.0000001, f[q]); a1 = fma(a1, 1.0000001, f[q + 1]);
        a2 = fma(a2, 1.0000001, f[q + 2]); a3 = fma(a3, 1.0000001, f[q + 3]);
      }
      a0 = fma(a0, 1.0000001, f[16]); a1 = fma(a1, 1.0000001, f[17]); a2 = fma(a2, 1.0000001, f[18]);
    }
    acc = (a0 + a1) + (a2 + a3);
  } else {
    for (int it = 0; it < work; ++it) {
#pragma unroll
      for (int q = 0; q < Q; ++q) acc = fma(acc, 1.0000001, f[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) b[idx<LAY, R>(m * Q + q, push ? (long long)t[q] : pos, fs)] = f[q] + acc;
}
// two-kernel split: A = adjacency + gathers -> 3 force components per lane; B = populations + forces, work, push
__global__ void __launch_bounds__(128) k_forceA(const double *__restrict__ rho, const uint32_t *__restrict__ nb, double *__restrict__ F,
                                                long long fs, long long n) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, m = lane >> 4, j = lane & 15;
  const long long pos = w * 16 + j;
  if (pos < 1 || pos >= n - 1) return;
  unsigned cen[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) cen[k] = __ldg(nb + idx<0, NC>(k, pos, fs));
  double g0 = 0., g1 = 0., g2 = 0.;
  g0 += __ldg(rho + (long long)m * fs + pos + 1) - __ldg(rho + (long long)m * fs + pos - 1);
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const double v = __ldg(rho + (long long)m * fs + cen[k]);
    g1 += (k & 1) ? -v : v;
    g2 += (k & 2) ? -v : v;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g0 += __ldg(rho + (long long)m * fs + cen[k] + 1) - __ldg(rho + (long long)m * fs + cen[k] - 1);
  }
  F[(long long)(m * 3 + 0) * fs + pos] = g0;
  F[(long long)(m * 3 + 1) * fs + pos] = g1;
  F[(long long)(m * 3 + 2) * fs + pos] = g2;
}
template <int VAR>
__global__ void __launch_bounds__(128) k_collB(const double *__restrict__ a, double *__restrict__ b, const double *__restrict__ F,
                                               const uint32_t *__restrict__ nb, long long fs, long long n, int work) {
  extern __shared__ unsigned char dummy[];
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, m = lane >> 4, j = lane & 15;
  const long long pos = w * 16 + j;
  if (pos < 1 || pos >= n - 1) return;
  unsigned cen[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) cen[k] = __ldg(nb + idx<0, NC>(k, pos, fs));
  double f[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) f[q] = __ldg(a + idx<0, R>(m * Q + q, pos, fs));
  const double g = __ldg(F + (long long)(m * 3 + 0) * fs + pos) + __ldg(F + (long long)(m * 3 + 1) * fs + pos) + __ldg(F + (long long)(m * 3 + 2) * fs + pos);
  unsigned t[Q];
  t[0] = (unsigned)pos; t[1] = t[0] + 1; t[2] = t[0] - 1;
#pragma unroll
  for (int k = 0; k < NC; ++k) t[3 + k] = cen[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) { t[11 + 2 * k] = cen[k] + 1; t[12 + 2 * k] = cen[k] - 1; }
  double acc = g;
  if (VAR & 4) {
    double a0 = g, a1 = g + 1., a2 = g + 2., a3 = g + 3.;
    for (int it = 0; it < work; ++it) {
#pragma unroll
      for (int q = 0; q + 3 < Q; q += 4) {
        a0 = fma(a0, 1.0000001, f[q]); a1 = fma(a1, 1.0000001, f[q + 1]);
        a2 = fma(a2, 1.0000001, f[q + 2]); a3 = fma(a3, 1.0000001, f[q + 3]);
      }
      a0 = fma(a0, 1.0000001, f[16]); a1 = fma(a1, 1.0000001, f[17]); a2 = fma(a2, 1.0000001, f[18]);
    }
    acc = (a0 + a1) + (a2 + a3);
  } else {
    for (int it = 0; it < work; ++it) {
#pragma unroll
      for (int q = 0; q < Q; ++q) acc = fma(acc, 1.0000001, f[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) b[idx<0, R>(m * Q + q, (long long)t[q], fs)] = f[q] + acc;
}
int main(int argc, char **argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 60000000ll / 256 * 256;
  const long long fs = n + 128;
  double *a, *b, *rho; uint32_t *nb;
  CK(cudaMalloc(&a, sizeof(double) * R * fs)); CK(cudaMalloc(&b, sizeof(double) * R * fs));
  CK(cudaMalloc(&rho, sizeof(double) * S * fs)); CK(cudaMalloc(&nb, sizeof(uint32_t) * NC * fs));
  CK(cudaMemset(a, 0, sizeof(double) * R * fs)); CK(cudaMemset(b, 0, sizeof(double) * R * fs)); CK(cudaMemset(rho, 0, sizeof(double) * S * fs));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double gb = (2.0 * R * 8 + NC * 4 + 16) * n / 1e9;
  const unsigned blocks = (unsigned)((n / 16 + 3) / 4);
  double *F; CK(cudaMalloc(&F, sizeof(double) * 6 * fs)); CK(cudaMemset(F, 0, sizeof(double) * 6 * fs));
  k_init_nbr<<<(unsigned)((n + 255) / 256), 256>>>(nb, fs, n, 0);
  CK(cudaDeviceSynchronize());
  auto timeit = [&](const char *name, auto launch) {
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 4; ++i) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 4;
    printf("%-72s %8.3f ms  %7.1f GB/s\n", name, ms, gb / (ms * 1e-3));
  };
  const int SM = 48 * 1024;
#define VARIANT(V, label) { CK(cudaFuncSetAttribute(k_lbm<0, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); \
    for (int work : {0, 30}) for (int gath : {0, 1}) { char nm[128]; snprintf(nm, sizeof nm, "16 w/SM %-34s work=%2d gath=%d push=1", label, work, gath); \
      timeit(nm, [&] { k_lbm<0, V><<<blocks, 128, SM>>>(a, b, rho, nb, fs, n, work, gath, 1); }); } }
  VARIANT(0, "dependent gathers, 1 chain")
  VARIANT(4, "dependent gathers, 4 chains")
  VARIANT(5, "arithmetic gather addresses, 4 chains")
  VARIANT(6, "centre gathers + shuffles, 4 chains")
  CK(cudaFuncSetAttribute(k_collB<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  CK(cudaFuncSetAttribute(k_collB<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  timeit("split A: adjacency + 18 gathers -> F (full occupancy)", [&] { k_forceA<<<blocks, 128>>>(rho, nb, F, fs, n); });
  timeit("split B: f + F, work=30 (4 chains), push, 16 w/SM", [&] { k_collB<4><<<blocks, 128, SM>>>(a, b, F, nb, fs, n, 30); });
  timeit("split B: f + F, work=30 (1 chain), push, 16 w/SM", [&] { k_collB<0><<<blocks, 128, SM>>>(a, b, F, nb, fs, n, 30); });
  timeit("split B: f + F, work=0, push, 16 w/SM", [&] { k_collB<4><<<blocks, 128, SM>>>(a, b, F, nb, fs, n, 0); });
  return 0;
}

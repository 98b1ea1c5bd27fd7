// setup_kernels.cuh -- non-templated helper kernels (included by flow.cu only)
#pragma once
#include <cstdint>

namespace txg {

// walls(rg..) doubles -> u8 classes (include/taxila_gpu.h TXG_CLASS_*); exact, value by value
__global__ void k_classify(const double *__restrict__ walls, uint8_t *__restrict__ cls, long long n,
                           int *__restrict__ bad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double w = walls[i];
  uint8_t c;
  if (w == 0.)
    c = 0;
  else if (w >= 1. && w <= 100. && w == (double)(int)w)
    c = (uint8_t)(int)w;
  else if (w == 900.)
    c = 250;
  else if (w == 901.)
    c = 251;
  else if (w == 902.)
    c = 252;
  else if (w == 800.)
    c = 253;
  else if (w == 999.)
    c = 255;
  else if (w > 0.)
    c = 254;
  else {
    c = 254;  // negative / NaN: the reference treats "not .eq. 0" as non-fluid
    atomicAdd(bad, 1);
  }
  cls[i] = c;
}

// fluid-node list, pass 1: fluid nodes (nbmask bit 31 clear) per 256-slot chunk of each plane
// (and the number of fluid nodes that carry a wall record, bit 30, accumulated into *nrec)
__global__ void k_count_fluid(const uint32_t *__restrict__ nbmask, long long plane, int bpp, unsigned *__restrict__ cnt,
                              int *__restrict__ nrec) {
  const long long z = blockIdx.x / bpp;
  const long long r = (long long)(blockIdx.x % bpp) * 256 + threadIdx.x;
  const uint32_t mask = r < plane ? nbmask[z * plane + r] : 0x80000000u;
  const bool fluid = !(mask >> 31);
  const int n = __syncthreads_count(fluid);
  const int w = __syncthreads_count(fluid && (mask & 0x40000000u));
  if (threadIdx.x == 0) {
    cnt[blockIdx.x] = (unsigned)n;
    if (w) atomicAdd(nrec, w);
  }
}

// pass 2: off[chunk] = list position of the chunk's first fluid node; rank inside the chunk by ballot
__global__ void k_fill_fluid(const uint32_t *__restrict__ nbmask, long long plane, int bpp,
                             const unsigned *__restrict__ off, uint32_t *__restrict__ list) {
  __shared__ unsigned warp_cnt[8];
  const long long z = blockIdx.x / bpp;
  const long long r = (long long)(blockIdx.x % bpp) * 256 + threadIdx.x;
  const bool fluid = r < plane && !(nbmask[z * plane + r] >> 31);
  const unsigned b = __ballot_sync(0xffffffffu, fluid);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[w] = __popc(b);
  __syncthreads();
  unsigned base = off[blockIdx.x];
  for (int k = 0; k < w; ++k) base += warp_cnt[k];
  if (fluid) list[base + __popc(b & ((1u << lane) - 1u))] = (uint32_t)(z * plane + r);
}

// host AoS (ghosted, dof = K*S with index k*S+m) -> device SoA [(m*K+k)][z][y][x] and back.
// The staging buffer holds the nzl ghosted (in x,y) z-planes of owned planes [zl0, zl0+nzl); only owned
// nodes are touched.
__global__ void k_import_aos(const double *__restrict__ src, double *__restrict__ dst, int NX, int NY, int gw,
                             int S, int K, long long dst_stride, long long dst_plane0, int zl0, int nzl) {
  // one thread per (zl, y, x, c)
  const int dof = S * K;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nzl * NY * NX * dof;
  if (idx >= total) return;
  int c = (int)(idx % dof);
  long long node = idx / dof;
  int x = (int)(node % NX);
  int y = (int)((node / NX) % NY);
  int zz = (int)(node / ((long long)NX * NY));
  const int gnx = NX + 2 * gw, gny = NY + 2 * gw;
  long long s = (((long long)zz * gny + (y + gw)) * gnx + (x + gw)) * dof + c;
  int m = c % S, k = c / S;
  dst[(long long)(m * K + k) * dst_stride + (dst_plane0 + zl0 + zz) * ((long long)NX * NY) + (long long)y * NX + x] = src[s];
}

__global__ void k_export_aos(double *__restrict__ dst, const double *__restrict__ src, int NX, int NY, int gw, int S,
                             int K, long long src_stride, long long src_plane0, int zl0, int nzl) {
  const int dof = S * K;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nzl * NY * NX * dof;
  if (idx >= total) return;
  int c = (int)(idx % dof);
  long long node = idx / dof;
  int x = (int)(node % NX);
  int y = (int)((node / NX) % NY);
  int zz = (int)(node / ((long long)NX * NY));
  const int gnx = NX + 2 * gw, gny = NY + 2 * gw;
  long long d = (((long long)zz * gny + (y + gw)) * gnx + (x + gw)) * dof + c;
  int m = c % S, k = c / S;
  dst[d] = src[(long long)(m * K + k) * src_stride + (src_plane0 + zl0 + zz) * ((long long)NX * NY) + (long long)y * NX + x];
}

// max |(old - cur)/cur| over a range, then old = cur  (DistributionCalcDeltaNorm,
// lbm_distribution_function.F90:809-833).  NaN-propagating like VecNorm is not attempted: a NaN
// ratio (0/0 on an untouched entry) is skipped, which is what fluid-only storage implies.
__global__ void k_delta_norm(const double *__restrict__ cur, double *__restrict__ old, long long n,
                             unsigned long long *__restrict__ out_bits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.;
  if (i < n) {
    const double c = cur[i], o = old[i];
    if (c != 0.) v = fabs((o - c) / c);
    old[i] = c;
  }
  // warp max then one atomic per warp; doubles >= 0 order like their bit patterns
  for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
  if ((threadIdx.x & 31) == 0 && v > 0.) atomicMax(out_bits, (unsigned long long)__double_as_longlong(v));
}

}  // namespace txg

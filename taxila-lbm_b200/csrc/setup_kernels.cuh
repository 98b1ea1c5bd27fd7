// setup_kernels.cuh -- non-templated helper kernels (included by flow.cu only)
#pragma once
#include <cstdint>

#include "bitrows.cuh"

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

// fluid-node list over the EXTENDED slab (owned planes + Rz ghost planes each side), from the classes.
// pass 1: fluid nodes per 256-slot chunk of each extended plane
__device__ __forceinline__ bool ext_fluid(const Grid &g, const uint8_t *__restrict__ cls, long long zz, long long r) {
  const int y = (int)(r / g.NX), x = (int)(r - (long long)y * g.NX);
  return cls[((long long)zz * g.cny + (y + g.R)) * g.cnx + (x + g.R)] == 0;
}
__global__ void k_count_fluid(Grid g, const uint8_t *__restrict__ cls, int bpp, unsigned *__restrict__ cnt) {
  const long long zz = blockIdx.x / bpp;
  const long long r = (long long)(blockIdx.x % bpp) * 256 + threadIdx.x;
  const bool fluid = r < g.plane && ext_fluid(g, cls, zz, r);
  const int n = __syncthreads_count(fluid);
  if (threadIdx.x == 0) cnt[blockIdx.x] = (unsigned)n;
}

// pass 2: off[chunk] = position of the chunk's first fluid node; rank inside the chunk by ballot.
// Writes P[oe] for every node (fluid or not) and list[P[oe]] = oe for the fluid ones.
__global__ void k_fill_fluid(Grid g, const uint8_t *__restrict__ cls, int bpp, const unsigned *__restrict__ off,
                             uint32_t *__restrict__ P, uint32_t *__restrict__ list) {
  __shared__ unsigned warp_cnt[8];
  const long long zz = blockIdx.x / bpp;
  const long long r = (long long)(blockIdx.x % bpp) * 256 + threadIdx.x;
  const bool in = r < g.plane;
  const bool fluid = in && ext_fluid(g, cls, zz, r);
  const unsigned b = __ballot_sync(0xffffffffu, fluid);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[w] = __popc(b);
  __syncthreads();
  unsigned base = off[blockIdx.x];
  for (int k = 0; k < w; ++k) base += warp_cnt[k];
  const unsigned pos = base + __popc(b & ((1u << lane) - 1u));
  if (in) {
    const long long oe = zz * g.plane + r;
    P[oe] = pos;
    if (fluid) list[pos] = (uint32_t)oe;
  }
}

// per-position copy of the dense masks of the owned nodes; counts the wall records
__global__ void k_gather_mask(Grid g, const uint32_t *__restrict__ nbmask, const uint32_t *__restrict__ list,
                              uint32_t *__restrict__ lmask, int *__restrict__ nrec) {
  const long long pos = g.own0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool rec = false;
  if (pos < g.own1) {
    const long long oe = list ? (long long)list[pos] : pos;
    const uint32_t m = nbmask[oe - (long long)g.Rz * g.plane];
    lmask[pos] = m;
    rec = (m & 0x40000000u) != 0;
  }
  const int w = __syncthreads_count(rec);
  if (threadIdx.x == 0 && w) atomicAdd(nrec, w);
}

// host AoS (ghosted, dof = K*S with index k*S+m) <-> device SoA block (m*K+k).  The staging buffer
// holds the nzl ghosted (in x,y) z-planes of owned planes [zl0, zl0+nzl); only owned nodes are touched.
// dense = 1: the device array is [S*K][nnodes] over the owned nodes; dense = 0: position-indexed
// [S*K][fs] holding fluid nodes only (solid nodes are skipped on import and exported as 0).
__global__ void k_import_aos(Grid g, const double *__restrict__ src, double *__restrict__ dst, int gw, int S, int K,
                             int dense, int zl0, int nzl) {
  const int dof = S * K;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nzl * g.plane * dof;
  if (idx >= total) return;
  int c = (int)(idx % dof);
  long long node = idx / dof;
  int x = (int)(node % g.NX);
  int y = (int)((node / g.NX) % g.NY);
  int zz = (int)(node / g.plane);
  const int gnx = g.NX + 2 * gw, gny = g.NY + 2 * gw;
  const double v = src[(((long long)zz * gny + (y + gw)) * gnx + (x + gw)) * dof + c];
  int m = c % S, k = c / S;
  const long long o = (long long)(zl0 + zz) * g.plane + (long long)y * g.NX + x;
  if (dense) {
    dst[(long long)(m * K + k) * g.nnodes + o] = v;
  } else {
    const long long oe = o + (long long)g.Rz * g.plane;
    if (!g.P)
      dst[(long long)(m * K + k) * g.fs + oe] = v;
    else if (g.P[oe + 1] != g.P[oe])
      dst[(long long)(m * K + k) * g.fs + g.P[oe]] = v;
  }
}

__global__ void k_export_aos(Grid g, double *__restrict__ dst, const double *__restrict__ src, int gw, int S, int K,
                             int dense, int zl0, int nzl) {
  const int dof = S * K;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nzl * g.plane * dof;
  if (idx >= total) return;
  int c = (int)(idx % dof);
  long long node = idx / dof;
  int x = (int)(node % g.NX);
  int y = (int)((node / g.NX) % g.NY);
  int zz = (int)(node / g.plane);
  const int gnx = g.NX + 2 * gw, gny = g.NY + 2 * gw;
  int m = c % S, k = c / S;
  const long long o = (long long)(zl0 + zz) * g.plane + (long long)y * g.NX + x;
  double v = 0.;
  if (dense) {
    v = src[(long long)(m * K + k) * g.nnodes + o];
  } else {
    const long long oe = o + (long long)g.Rz * g.plane;
    if (!g.P)
      v = src[(long long)(m * K + k) * g.fs + oe];
    else if (g.P[oe + 1] != g.P[oe])
      v = src[(long long)(m * K + k) * g.fs + g.P[oe]];
  }
  dst[(((long long)zz * gny + (y + gw)) * gnx + (x + gw)) * dof + c] = v;
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

// ------------------------------------------------------------------ set-up of the bit rows (once per walls upload)
// one thread per (extended row incl. the two y ghost rows, word): bits and edge flags from the class array (its ghost
// rows / columns hold the wrapped classes of a periodic box and wall codes otherwise), the start from the position map
__global__ void k_build_bitrows(Grid g, const uint8_t *__restrict__ cls, int NW, BitrowEntry *__restrict__ rows,
                                uint32_t *__restrict__ rowend) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rpp = g.NY + 2, nzE = g.NZl + 2 * g.Rz;
  if (t >= (long long)nzE * rpp * NW) return;
  const int w = (int)(t % NW);
  const long long row = t / NW;
  const int zz = (int)(row / rpp), yg = (int)(row - (long long)zz * rpp) - 1;  // yg = -1 .. NY
  // the class array has R >= 1 ghost rows: row yg sits at yg + R
  const uint8_t *crow = cls + ((long long)zz * g.cny + (yg + g.R)) * g.cnx + g.R;
  auto fluid = [&](int x) -> bool { return x >= -1 && x <= g.NX && crow[x] == 0; };
  uint32_t bits = 0u;
  for (int b = 0; b < 32; ++b) {
    const int x = 32 * w + b;
    if (x < g.NX && fluid(x)) bits |= 1u << b;
  }
  // the real row behind a y ghost row (periodic: the wrapped row; otherwise the row is solid and its start is unused)
  int yr = yg;
  if (yg < 0) yr = g.pery ? g.NY - 1 : 0;
  if (yg >= g.NY) yr = g.pery ? 0 : g.NY - 1;
  const long long oe0 = ((long long)zz * g.NY + yr) * g.NX;
  const int x0 = min(32 * w, g.NX);
  const uint32_t start = g.P ? g.P[oe0 + x0] : (uint32_t)(oe0 + x0);
  uint32_t se = start & BITROW_POSMASK;
  if (fluid(32 * w - 1)) se |= 1u << 30;
  if (32 * w + 32 <= g.NX && fluid(32 * w + 32)) se |= 1u << 31;
  // the flag of x = NX (ghost column) when the row does not end on a word boundary: the bit after the last node.
  // No prefix count ever includes it (b <= (NX - 1) & 31 for every node of the row).
  if (32 * w < g.NX && g.NX < 32 * w + 32 && fluid(g.NX)) bits |= 1u << (g.NX & 31);
  rows[t] = BitrowEntry{se, bits};
  if (w == 0) rowend[row] = g.P ? g.P[oe0 + g.NX] : (uint32_t)(oe0 + g.NX);
}

// xrow[pos] = x | rowid << 11 for every stored position (owned and ghost planes)
__global__ void k_build_xrow(Grid g, long long pos0, long long nstore, uint32_t *__restrict__ xrow) {
  const long long pos = pos0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= nstore) return;
  const unsigned oe = g.list ? g.list[pos] : (unsigned)pos;
  const unsigned plane = (unsigned)g.plane;
  const unsigned zz = oe / plane, r = oe - zz * plane;
  const unsigned y = r / (unsigned)g.NX, x = r - y * (unsigned)g.NX;
  xrow[pos] = x | ((zz * (unsigned)(g.NY + 2) + y + 1u) << BITROW_XBITS);
}

// the solid nodes of the dense diagnostics arrays: rhot = 0, prs = null_pressure, velt = 0 (what k_export writes there)
__global__ void k_export_fill_solid(Grid g, const uint32_t *__restrict__ nbmask, int D, double *__restrict__ rhot, double *__restrict__ prs,
                                    double *__restrict__ velt, double null_pressure) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= g.nnodes || !(nbmask[o] >> 31)) return;
  if (rhot) rhot[o] = 0.;
  if (prs) prs[o] = null_pressure;
  if (velt)
    for (int d = 0; d < D; ++d) velt[o * D + d] = 0.;
}

}  // namespace txg

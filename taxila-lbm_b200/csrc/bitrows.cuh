// bitrows.cuh -- adjacency of the fluid-only storage without a per-node table.
//
// Positions ascend in (z, y, x), so the position of any node follows from the start of its x-row and the number
// of fluid nodes before it in that row.  The row table holds, per extended row (zz, y) -- y = -1 .. NY: the two
// extra rows are the y ghost rows of the class array, i.e. the wrapped rows when y is periodic and all-solid
// otherwise -- and per 32-node word w of the row, one 8-byte entry
//     bits  : bit b = node x = 32 w + b is fluid                                  (nodes beyond NX: 0)
//     se    : position of the first fluid node at or after x = 32 w  (28 bits: Q * fs < 2^32 bounds every position)
//             | bit 30: node x = 32 w - 1 is fluid   | bit 31: node x = 32 w + 32 is fluid
//             (x = -1 and x = NX are the x ghost columns of the class array: the wrapped nodes when x is periodic)
// so that a lane at x reads ONE entry per neighbour row and gets the fluid flags of x-1, x, x+1 and the positions
// P(x-1), P(x), P(x+1) of that row (P of a solid node = position of the next fluid node, like kernels.cuh).  The
// 18 neighbour positions and the solid mask of a D3Q19 node cost 9 such entries -- 72 bytes that hit L1, shared by
// all lanes of the row -- instead of 18 x 4 + 4 bytes per node streamed from HBM (nbr_all + lmask: 4.6 GB per step
// at 512^3).  Only the periodic wrap in x needs a second look-up (rowend / word 0 of the row), on the two face columns.
//
// xrow[pos] = x | rowid << 11 (rowid = zz * (NY + 2) + y + 1) tells a lane where it is: 4 bytes per node, coalesced.
//
// Checked on the device: k_step_band must give the bits of the table-driven kernel on every box of
// tests/test_zgpu_step_forms.py (periodic and closed faces, rows of several words, 2-D, one to three components).
#pragma once
#include <cstdint>

#ifndef TXG_HD
#define TXG_HD __host__ __device__ __forceinline__
#endif

namespace txg {

constexpr int BITROW_XBITS = 11;                 // NX <= 2048
constexpr uint32_t BITROW_POSMASK = 0x0fffffffu;  // 28-bit positions
constexpr uint32_t BITROW_MAX_ROWS = 1u << (32 - BITROW_XBITS);

struct BitrowEntry {
  uint32_t se, bits;
};

TXG_HD int txg_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
TXG_HD uint32_t txg_funnel_r(uint32_t lo, uint32_t hi, int s) {  // bits [s, s+32) of hi:lo, 0 <= s < 32
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return s == 0 ? lo : (lo >> s) | (hi << (32 - s));
#endif
}

// The three nodes x-1, x, x+1 of one row, from the row's entry of word x >> 5 (b = x & 31):
//   win bit 0/1/2 = node x-1 / x / x+1 is fluid;  pm, p0, pp = P(x-1), P(x), P(x+1).
// For a fluid node these are the nodes' positions.  On a periodic x face the flag of the wrapped node is right and
// its position is not (see bitrow_wrap_*).
struct RowTriple {
  uint32_t win, pm, p0, pp;
};
TXG_HD RowTriple bitrow_triple(BitrowEntry e, int b) {
  const uint32_t lo = (e.bits << 1) | ((e.se >> 30) & 1u);  // bit j = node 32 w + j - 1
  const uint32_t hi = (e.bits >> 31) | ((e.se >> 31) << 1);
  RowTriple t;
  t.win = txg_funnel_r(lo, hi, b) & 7u;
  t.p0 = (e.se & BITROW_POSMASK) + (uint32_t)txg_popc(e.bits & ((1u << b) - 1u));
  t.pm = t.p0 - (t.win & 1u);
  t.pp = t.p0 + ((t.win >> 1) & 1u);
  return t;
}
// position of a single node of the row (any word): the entry of word x >> 5
TXG_HD uint32_t bitrow_pos(BitrowEntry e, int b) { return (e.se & BITROW_POSMASK) + (uint32_t)txg_popc(e.bits & ((1u << b) - 1u)); }
TXG_HD bool bitrow_fluid(BitrowEntry e, int b) { return (e.bits >> b) & 1u; }

}  // namespace txg

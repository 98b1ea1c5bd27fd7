// lag_schedule.h -- host-only (plain C++, no CUDA): the block schedule of the one-pass step (k_step_fused_lag).
// Included by flow.cu and compiled on its own by tests/test_lag_schedule.py, which checks coverage and the
// dependency claim below against the real adjacency of random porous boxes.
//
// Why.  The two-kernel step reads the populations twice (k_moments: 8SQ bytes per fluid node, then the fused
// forces + collide + push: 16SQ), which caps it at 641/945 = 68 % of the step's HBM roofline.
// The density of a node is complete once every node that pushes into it has collided -- its 3 x 3 (y, z)
// neighbourhood of rows -- so the sum can run inside the same launch, a little behind the collision front,
// while the freshly pushed populations are still in L2.  "A little" must be small against L2: in the storage
// order (z, y, x) the front is a whole 512^2 plane (36 MB of pushed populations + as much streamed input)
// and the window did not survive (profiles/r1b_ablations.txt, r1h).  This schedule therefore walks the slab
// in y-bands: band b = rows [b*BR, (b+1)*BR), and inside a band plane by plane.  The rows of one band in one
// plane are one contiguous range of positions, because positions ascend in (z, y, x).
//
// The launch is a 1-D grid of nrows * grid_x blocks: block index = row * grid_x + x (schedule row, block inside the
// row); blocks are dispatched in index order, so a row's blocks follow all blocks of earlier rows.  Row
// r = b * rows_per_band + k of band b holds
//   C(b, k)   if k < NZl: collide + push the fluid nodes of rows [b*BR, (b+1)*BR) of owned plane k, PB positions
//             per block, blocks x = 0 .. nC-1; the last thing a C block does is fence + done[r] += 1;
//   M(b, zm)  with zm = k - 1 - lag, if 1 <= zm <= NZl-2: sum the new populations of the rows of plane zm whose
//             sources complete with band b -- row y belongs to band max(band(y-1), band(y), band(y+1)), wrapped
//             when y is periodic; at most two runs of consecutive rows -- MB positions per block, blocks
//             x = nC .. nC+nM-1.  Planes 0 and NZl-1 receive populations through the z halo and are summed by
//             k_moments after it.  An M block first waits until done[r'] == nC(r') for the <= 9 rows
//             r' = b' * rows_per_band + zm + dz, b' in depbands[b], dz in {-1, 0, 1}.
// Dependency claim: every node that pushes into a position of an M block (its fluid lattice neighbours and,
// for bounce-back, itself) belongs to a C block of one of those rows, and every C block of those rows precedes
// the M block in linear block order (earlier row, or same row and smaller x).  With the hardware dispatching
// blocks in linear order a waiting M block therefore only waits for blocks that are resident or finished.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace txg {

struct LagRow {
  uint32_t cfirst, ccount;    // C part: positions [cfirst, cfirst + ccount)
  uint32_t m0first, m0count;  // M part, first run of rows
  uint32_t m1first, m1count;  // M part, second run (the wrapped row 0 of the last band)
};
constexpr int LAG_MAX_BANDS = 16;

struct LagSchedule {
  std::vector<LagRow> rows;
  int nbands = 0, rows_per_band = 0, lag = 0, PB = 0, MB = 0;
  int depbands[LAG_MAX_BANDS][3];  // bands whose C rows an M block of band b waits for (-1 = none)
  uint32_t grid_x = 0;             // max over rows of nC + nM
  long long c_blocks = 0, m_blocks = 0;
  bool ok = false;  // false: box not eligible, rows empty
};

inline uint32_t lag_blocks(uint32_t count, int per_block) { return (count + (uint32_t)per_block - 1) / (uint32_t)per_block; }

// row_off: [(NZl + 2 Rz) * NY + 1] position of the first fluid node at or after the start of each extended row
//          (zz, y), zz = z + Rz; the last entry is the number of stored nodes.  (= P[(zz*NY + y)*NX].)
// PB / MB: positions per C block / per M block; BR: rows per band (>= 2; raised until the bands fit
// LAG_MAX_BANDS and the rows fit max_rows); lag: extra planes between the collision of plane z + 1 and the sum of plane z.
inline LagSchedule build_lag_schedule(int NY, int NZl, int Rz, int pery, const uint32_t *row_off, int PB, int MB, int BR,
                                      int lag, int max_rows) {
  LagSchedule s;
  if (Rz < 1 || NZl < 4 || NY < 4 || BR < 2 || PB < 1 || MB < PB || lag < 0 || lag > NZl) return s;
  BR = std::min(BR, NY);
  const int rpb = NZl + lag;
  while ((NY + BR - 1) / BR > LAG_MAX_BANDS || (long long)((NY + BR - 1) / BR) * rpb > max_rows) {
    if (BR >= NY) return s;
    BR = std::min(NY, BR * 2);
  }
  const int NB = (NY + BR - 1) / BR;
  auto band = [&](int y) { return y / BR; };
  auto roff = [&](int z, int y) { return row_off[(size_t)(z + Rz) * NY + y]; };  // y = NY: start of the next plane
  auto wrap = [&](int y) { return (y < 0 || y >= NY) ? (pery ? (y + NY) % NY : -1) : y; };
  // the band whose collision completes the sources of row y
  std::vector<int> mb((size_t)NY);
  for (int y = 0; y < NY; ++y) {
    int m = band(y);
    for (int dy = -1; dy <= 1; dy += 2)
      if (wrap(y + dy) >= 0) m = std::max(m, band(wrap(y + dy)));
    mb[(size_t)y] = m;
  }
  // per band: the runs of rows it sums (at most two) and the bands it waits for (at most three)
  struct Runs {
    int y0[2], y1[2], n = 0;
  };
  std::vector<Runs> runs((size_t)NB);
  for (int b = 0; b < NB; ++b) {
    for (int k = 0; k < 3; ++k) s.depbands[b][k] = -1;
    int nd = 0;
    for (int y = 0; y < NY;) {
      if (mb[(size_t)y] != b) {
        ++y;
        continue;
      }
      int y1 = y;
      while (y1 < NY && mb[(size_t)y1] == b) ++y1;
      Runs &r = runs[(size_t)b];
      if (r.n == 2) return s;
      r.y0[r.n] = y;
      r.y1[r.n] = y1;
      ++r.n;
      for (int q = y; q < y1; ++q)
        for (int dy = -1; dy <= 1; ++dy) {
          const int yy = wrap(q + dy);
          if (yy < 0) continue;
          const int bb = band(yy);
          if (bb == s.depbands[b][0] || bb == s.depbands[b][1] || bb == s.depbands[b][2]) continue;
          if (nd == 3) return s;
          s.depbands[b][nd++] = bb;
        }
      y = y1;
    }
  }
  s.rows.assign((size_t)NB * rpb, LagRow{0, 0, 0, 0, 0, 0});
  for (int b = 0; b < NB; ++b) {
    const int y0 = b * BR, y1 = std::min(NY, y0 + BR);
    for (int k = 0; k < rpb; ++k) {
      LagRow &r = s.rows[(size_t)b * rpb + k];
      if (k < NZl) {
        r.cfirst = roff(k, y0);
        r.ccount = roff(k, y1) - roff(k, y0);
      }
      const int zm = k - 1 - lag;
      if (zm >= 1 && zm <= NZl - 2) {
        const Runs &ru = runs[(size_t)b];
        if (ru.n > 0) {
          r.m0first = roff(zm, ru.y0[0]);
          r.m0count = roff(zm, ru.y1[0]) - r.m0first;
        }
        if (ru.n > 1) {
          r.m1first = roff(zm, ru.y0[1]);
          r.m1count = roff(zm, ru.y1[1]) - r.m1first;
        }
      }
      const uint32_t nc = lag_blocks(r.ccount, PB), nm = lag_blocks(r.m0count, MB) + lag_blocks(r.m1count, MB);
      s.c_blocks += nc;
      s.m_blocks += nm;
      s.grid_x = std::max(s.grid_x, nc + nm);
    }
  }
  s.nbands = NB;
  s.rows_per_band = rpb;
  s.lag = lag;
  s.PB = PB;
  s.MB = MB;
  s.ok = s.grid_x > 0;
  return s;
}

}  // namespace txg

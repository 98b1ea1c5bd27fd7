// C entry point around taxila-lbm_b200/csrc/lag_schedule.h (host-only code of the CUDA library) so that
// tests/test_lag_schedule.py can check the very schedule builder flow.cu uses, without a GPU.
#include "../../taxila-lbm_b200/csrc/lag_schedule.h"

// returns the number of schedule rows (-1: box not eligible); rows as 6 u32 each; meta = {nbands, rows_per_band,
// grid_x, c_blocks, m_blocks}; depbands as [LAG_MAX_BANDS][3]
extern "C" long long lag_build(int NY, int NZl, int Rz, int pery, const uint32_t *row_off, int PB, int MB, int BR, int lag,
                               int max_rows, uint32_t *rows, long long cap, long long *meta, int32_t *depbands) {
  const txg::LagSchedule s = txg::build_lag_schedule(NY, NZl, Rz, pery, row_off, PB, MB, BR, lag, max_rows);
  if (!s.ok) return -1;
  for (long long i = 0; i < (long long)s.rows.size() && i < cap; ++i) {
    const txg::LagRow &r = s.rows[(size_t)i];
    const uint32_t v[6] = {r.cfirst, r.ccount, r.m0first, r.m0count, r.m1first, r.m1count};
    for (int k = 0; k < 6; ++k) rows[i * 6 + k] = v[k];
  }
  meta[0] = s.nbands;
  meta[1] = s.rows_per_band;
  meta[2] = s.grid_x;
  meta[3] = s.c_blocks;
  meta[4] = s.m_blocks;
  for (int b = 0; b < txg::LAG_MAX_BANDS; ++b)
    for (int k = 0; k < 3; ++k) depbands[b * 3 + k] = b < s.nbands ? s.depbands[b][k] : -1;
  return (long long)s.rows.size();
}

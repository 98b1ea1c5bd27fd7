// C entry point around taxila-lbm_b200/csrc/specular_table.h (host-only code of the CUDA library) so that
// tests/test_specular_table.py can run the very table builder txg_set_walls uses, without a GPU.
#include <cstring>

#include "../../taxila-lbm_b200/csrc/specular_table.h"

extern "C" long long spec_build(int Q, int D, const int *c, int NX, int NY, int NZl, int R, int Rz, const int *per,
                                const uint8_t *cls, const uint32_t *P, long long fs, uint32_t *dst, uint32_t *src,
                                long long cap, long long *parked) {
  txg::LatticeTab lt;
  std::memset(&lt, 0, sizeof lt);
  lt.Q = Q;
  lt.D = D;
  for (int n = 0; n < Q; ++n)
    for (int d = 0; d < 3; ++d) lt.c[n][d] = c[n * 3 + d];
  txg::SpecularTable t;
  txg::build_specular_table(lt, NX, NY, NZl, R, Rz, per, cls, P, fs, t);
  *parked = t.parked;
  const long long n = (long long)t.dst.size();
  for (long long i = 0; i < n && i < cap; ++i) {
    dst[i] = t.dst[i];
    src[i] = t.src[i];
  }
  return n;
}

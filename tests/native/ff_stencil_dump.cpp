// Prints the Shan-Chen stencil DATA the device code is compiled with (csrc/ff_stencil.cuh entries + the isotropy weights
// of csrc/lattice.cuh) as JSON, for tests/test_stencil_independent.py.  Host build of the device headers: g++ with the
// CUDA function qualifiers defined away.
#define __host__
#define __device__
#define __forceinline__ inline
#include <cstdio>

#include "../../taxila-lbm_b200/csrc/lattice.cuh"

template <class L>
static void dump(const char *name, const int *orders, int norders, bool last) {
  using FF = typename L::FF;
  printf("\"%s\": {\"D\": %d, \"entries\": [", name, L::D);
  for (int e = 0; e < FF::E; ++e) {
    printf("%s{\"off\": [%d, %d, %d], \"L\": %d, \"gate\": %d, \"alts\": [", e ? ", " : "", FF::off[e][0], FF::off[e][1], FF::off[e][2],
           (int)FF::L[e], (int)FF::gate[e]);
    for (int a = 0; a < FF::nalt[e]; ++a) {
      printf("%s[", a ? ", " : "");
      for (int j = 0; j < FF::altlen[e][a]; ++j)
        printf("%s[%d, %d, %d]", j ? ", " : "", FF::los[e][a][j][0], FF::los[e][a][j][1], FF::los[e][a][j][2]);
      printf("]");
    }
    printf("]}");
  }
  printf("], \"ffw\": {");
  for (int k = 0; k < norders; ++k) {
    printf("%s\"%d\": [", k ? ", " : "", orders[k]);
    for (int l = 0; l <= 10; ++l) printf("%s%.17g", l ? ", " : "", L::ffw(orders[k], l));
    printf("]");
  }
  printf("}}%s\n", last ? "" : ",");
}

int main() {
  const int o3[] = {4, 8}, o2[] = {4, 8, 10};
  printf("{\n");
  dump<txg::D3Q19>("D3Q19", o3, 2, false);
  dump<txg::D2Q9>("D2Q9", o2, 3, true);
  printf("}\n");
  return 0;
}

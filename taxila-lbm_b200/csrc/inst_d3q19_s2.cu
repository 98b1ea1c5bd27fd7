// inst_d3q19_s2.cu -- kernel instantiations for D3Q19, 2 component(s)
#include "flow.h"
namespace txg {
bool kernel_set_d3q19_s2(bool mrt, int iso, KernelSet *out) {
  if (iso == 4) {
    *out = mrt ? make_kernel_set<D3Q19, 2, true, 4>("d3q19_s2_mrt_iso4")
               : make_kernel_set<D3Q19, 2, false, 4>("d3q19_s2_srt_iso4");
    return true;
  }
  if (iso == 8) {
    *out = mrt ? make_kernel_set<D3Q19, 2, true, 8>("d3q19_s2_mrt_iso8")
               : make_kernel_set<D3Q19, 2, false, 8>("d3q19_s2_srt_iso8");
    return true;
  }
  return false;
}
}  // namespace txg

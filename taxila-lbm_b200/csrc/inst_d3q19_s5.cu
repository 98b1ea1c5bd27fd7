// inst_d3q19_s5.cu -- kernel instantiations for D3Q19, 5 component(s)
#include "flow.h"
namespace txg {
bool kernel_set_d3q19_s5(bool mrt, int iso, KernelSet *out) {
  if (iso == 4) {
    *out = mrt ? make_kernel_set<D3Q19, 5, true, 4>("d3q19_s5_mrt_iso4")
               : make_kernel_set<D3Q19, 5, false, 4>("d3q19_s5_srt_iso4");
    return true;
  }
  if (iso == 8) {
    *out = mrt ? make_kernel_set<D3Q19, 5, true, 8>("d3q19_s5_mrt_iso8")
               : make_kernel_set<D3Q19, 5, false, 8>("d3q19_s5_srt_iso8");
    return true;
  }
  return false;
}
}  // namespace txg

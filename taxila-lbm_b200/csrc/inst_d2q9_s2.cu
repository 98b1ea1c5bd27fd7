// inst_d2q9_s2.cu -- kernel instantiations for D2Q9, 2 component(s)
#include "flow.h"
namespace txg {
bool kernel_set_d2q9_s2(bool mrt, int iso, KernelSet *out) {
  if (iso == 4) {
    *out = mrt ? make_kernel_set<D2Q9, 2, true, 4>("d2q9_s2_mrt_iso4")
               : make_kernel_set<D2Q9, 2, false, 4>("d2q9_s2_srt_iso4");
    return true;
  }
  if (iso == 8) {
    *out = mrt ? make_kernel_set<D2Q9, 2, true, 8>("d2q9_s2_mrt_iso8")
               : make_kernel_set<D2Q9, 2, false, 8>("d2q9_s2_srt_iso8");
    return true;
  }
  if (iso == 10) {
    *out = mrt ? make_kernel_set<D2Q9, 2, true, 10>("d2q9_s2_mrt_iso10")
               : make_kernel_set<D2Q9, 2, false, 10>("d2q9_s2_srt_iso10");
    return true;
  }
  return false;
}
}  // namespace txg

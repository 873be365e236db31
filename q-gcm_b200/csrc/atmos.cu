// Atmosphere mixed layer (aml + amladf, src/amlsubs.F:47-563) and the coupled forcing
// xforc (src/xfosubs.F:52-858).
#include "qgcm_internal.h"

namespace qg {

void launch_xforc(qgcm_model *m) {
  if (m->ocean_only) {
    // ocean_only builds execute only the oceanic Ekman tail (src/xfosubs.F:568-709)
    launch_xforc_ocean_ekman(m);
    return;
  }
  throw std::runtime_error("qgcm_xforc: coupled forcing kernels are not built yet");
}

void launch_aml(qgcm_model *m) {
  (void)m;
  throw std::runtime_error("qgcm_aml: atmospheric mixed layer kernels are not built yet");
}

}  // namespace qg

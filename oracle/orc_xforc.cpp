// TEST INFRASTRUCTURE ONLY -- CPU oracle (parity unpinned, see orc_model.h).
// Placeholder translation unit for the coupled forcing of src/xfosubs.F:140-563,
// :711-853 (bicubic/bilinear regridding, stress, fluxes).
#include "orc_model.h"
namespace orc {}

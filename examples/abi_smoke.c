/* The C ABI seen from plain C (C99): include/qgcm_b200.h compiles without C++, the shared
 * library links from C, and the error convention works (non-zero status + qgcm_last_error,
 * the counterpart of the reference's print-and-stop, src/nc_subs.F:84-112).  Needs no GPU:
 * qgcm_create rejects a configuration with the wrong ABI version before it touches CUDA.
 *
 *   gcc -std=c99 -pedantic -Wall -Iinclude examples/abi_smoke.c \
 *       -Lq-gcm_b200/csrc -lqgcm_b200 -Wl,-rpath,$PWD/q-gcm_b200/csrc -o /tmp/abi_smoke && /tmp/abi_smoke
 */
#include <stdio.h>
#include <string.h>

#include "qgcm_b200.h"

int main(void) {
  qgcm_config cfg;
  qgcm_model *m = NULL;
  qgcm_scalars s;
  qgcm_valids_report v;
  qgcm_monitor_ocean mo;
  qgcm_monitor_atmos ma;
  int32_t jp0 = -1, nown = -1;
  memset(&cfg, 0, sizeof cfg);
  printf("abi %d, sizeof: config %lu, scalars %lu, valids %lu, monitor_ocean %lu, monitor_atmos %lu\n", qgcm_abi_version(),
         (unsigned long)sizeof cfg, (unsigned long)sizeof s, (unsigned long)sizeof v, (unsigned long)sizeof mo, (unsigned long)sizeof ma);
  if (qgcm_abi_version() != QGCM_ABI_VERSION) return 1;
  cfg.abi_version = QGCM_ABI_VERSION + 1;        /* a caller built against another header */
  cfg.struct_bytes = (int32_t)sizeof cfg;
  if (qgcm_create(&cfg, &m) == 0) return 2;
  printf("rejected as expected: %s\n", qgcm_last_error());
  if (strstr(qgcm_last_error(), "ABI") == NULL) return 3;
  /* pure host arithmetic: the rows of a 4801-row grid owned by rank 3 of 8 */
  if (qgcm_slab_bounds(4801, 8, 3, &jp0, &nown) != 0 || nown < 600 || nown > 601) return 4;
  printf("rank 3 of 8 owns p rows [%d, %d)\n", (int)jp0, (int)(jp0 + nown));
  return 0;
}

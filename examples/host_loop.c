/* A compiled host driving the C ABI the way the Fortran main program would:
 * create the device model from a qgcm_config, upload the state the host owns, let the device
 * recompute what depends on it (src/q-gcm.F:711-976), run the time loop (src/q-gcm.F:1220-1408),
 * accumulate the running sums at their cadence, and download only what an output step needs.
 *
 *   host_loop <dir> <nt_last>
 * <dir>/config.bin      the qgcm_config, byte for byte
 * <dir>/fields.txt      one "name count" line per input field, data in <dir>/<name>.f64
 * writes <dir>/out_<name>.f64 for po, qo, sst, po_avg (and pa, ast when there is an atmosphere)
 * and prints a few monitor values.  tests/test_gpu_host_loop.py builds and runs it on the GPU box.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qgcm_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    if ((call) != 0) {                                                           \
      fprintf(stderr, "libqgcm_b200 error in %s: %s\n", #call, qgcm_last_error()); \
      return 1; /* the reference's convention is print + stop, src/nc_subs.F:84-112 */ \
    }                                                                            \
  } while (0)

static double *read_f64(const char *dir, const char *name, long n) {
  char path[1024];
  FILE *f;
  double *buf = (double *)malloc(sizeof(double) * (size_t)n);
  snprintf(path, sizeof path, "%s/%s.f64", dir, name);
  f = fopen(path, "rb");
  if (!f || !buf || fread(buf, sizeof(double), (size_t)n, f) != (size_t)n) {
    fprintf(stderr, "cannot read %s\n", path);
    exit(2);
  }
  fclose(f);
  return buf;
}

static int write_field(qgcm_model *m, const char *dir, const char *name) {
  char path[1024];
  int64_t n = 0;
  double *buf;
  FILE *f;
  CHECK(qgcm_field_size(m, name, &n));
  buf = (double *)malloc(sizeof(double) * (size_t)n);
  CHECK(qgcm_get_field(m, name, buf, n));
  snprintf(path, sizeof path, "%s/out_%s.f64", dir, name);
  f = fopen(path, "wb");
  if (!f || fwrite(buf, sizeof(double), (size_t)n, f) != (size_t)n) return 1;
  fclose(f);
  free(buf);
  return 0;
}

int main(int argc, char **argv) {
  const char *dir;
  long nt_last, nt;
  char path[1024], name[64];
  qgcm_config cfg;
  qgcm_model *m = NULL;
  qgcm_monitor_ocean mon;
  qgcm_valids_report val;
  int32_t nsumat, nsumoc, nsumk247;
  FILE *f;
  long count;
  int ocean_only;
  if (argc < 3) return 2;
  dir = argv[1];
  nt_last = atol(argv[2]);
  snprintf(path, sizeof path, "%s/config.bin", dir);
  f = fopen(path, "rb");
  if (!f || fread(&cfg, sizeof cfg, 1, f) != 1) { fprintf(stderr, "cannot read %s\n", path); return 2; }
  fclose(f);
  ocean_only = (cfg.flags & QGCM_OCEAN_ONLY) != 0;
  cfg.flags |= QGCM_OCNC_AVG_K247;               /* -Docnc_avg_k247: po accumulated after every ocean step */

  CHECK(qgcm_create(&cfg, &m));                  /* after homsol, src/q-gcm.F:976 */
  snprintf(path, sizeof path, "%s/fields.txt", dir);
  f = fopen(path, "r");
  if (!f) return 2;
  while (fscanf(f, "%63s %ld", name, &count) == 2) {
    double *buf = read_f64(dir, name, count);
    CHECK(qgcm_set_field(m, name, buf, (int64_t)count));
    free(buf);
  }
  fclose(f);
  /* the device recomputes what depends on the uploaded state: src/q-gcm.F:711-749, :870, :976 */
  CHECK(qgcm_constr(m));
  CHECK(qgcm_qcomp_ocean(m));
  if (!ocean_only) CHECK(qgcm_qcomp_atmos(m));
  CHECK(qgcm_xforc(m));
  CHECK(qgcm_homsol(m));
  CHECK(qgcm_tavini(m));                          /* src/q-gcm.F:1194 */

  for (nt = 1; nt <= nt_last; ++nt) {
    CHECK(qgcm_run(m, nt, nt));                   /* one pass of the loop body, src/q-gcm.F:1220-1408 */
    if (nt % cfg.nstr == 0) CHECK(qgcm_tavocn(m));  /* stands for mod(ntdone,ntavoc).eq.nmidoc, :1480 */
    if (nt == nt_last) {                          /* valids + monnc_comp on the device, :1278, :1442 */
      CHECK(qgcm_valids(m, &val));
      CHECK(qgcm_monnc_ocean(m, &mon));
      if (!val.solnok) { fprintf(stderr, "valids: solution invalid\n"); return 3; }
    }
  }
  CHECK(qgcm_tav_counts(m, &nsumat, &nsumoc, &nsumk247));
  printf("nsumoc %d nsum_ocavg %d kealoc1 %.15e utauoc %.15e cnmloc %.15e\n", (int)nsumoc, (int)nsumk247, mon.kealoc[0], mon.utauoc,
         mon.cnmloc);
  if (write_field(m, dir, "po") || write_field(m, dir, "qo") || write_field(m, dir, "sst") || write_field(m, dir, "po_avg") ||
      write_field(m, dir, "pocav"))
    return 1;
  if (!ocean_only && (write_field(m, dir, "pa") || write_field(m, dir, "ast"))) return 1;
  CHECK(qgcm_destroy(m));
  return 0;
}

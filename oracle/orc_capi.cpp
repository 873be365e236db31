// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (liborc.so), loaded by
// tests/ and by bench.py's cpu_baseline / --impl reference legs only.
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "orc_model.h"

using orc::Model;

static thread_local std::string g_err;

#define ORC_TRY(body)                 \
  try {                               \
    body;                             \
    return 0;                         \
  } catch (const std::exception &e) { \
    g_err = e.what();                 \
    return 1;                         \
  }

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

int orc_create(const qgcm_config *cfg, Model **out) { ORC_TRY(*out = new Model(*cfg)); }
int orc_destroy(Model *m) {
  delete m;
  return 0;
}

int orc_field_size(Model *m, const char *name, int64_t *n) {
  orc::vec *v = m->field(name);
  if (!v) {
    g_err = std::string("orc: unknown field ") + name;
    return 1;
  }
  *n = (int64_t)v->size();
  return 0;
}
int orc_set_field(Model *m, const char *name, const double *host, int64_t n) {
  orc::vec *v = m->field(name);
  if (!v || (int64_t)v->size() != n) {
    g_err = std::string("orc: bad field/size ") + name;
    return 1;
  }
  std::memcpy(v->data(), host, sizeof(double) * (size_t)n);
  return 0;
}
int orc_get_field(Model *m, const char *name, double *host, int64_t n) {
  orc::vec *v = m->field(name);
  if (!v || (int64_t)v->size() != n) {
    g_err = std::string("orc: bad field/size ") + name;
    return 1;
  }
  std::memcpy(host, v->data(), sizeof(double) * (size_t)n);
  return 0;
}
int orc_set_scalars(Model *m, const qgcm_scalars *s) {
  m->s = *s;
  return 0;
}
int orc_get_scalars(Model *m, qgcm_scalars *s) {
  *s = m->s;
  return 0;
}

int orc_constr(Model *m) { ORC_TRY(m->constr()); }
int orc_homsol(Model *m) { ORC_TRY(m->homsol()); }
int orc_qcomp_ocean(Model *m) { ORC_TRY(m->qcomp_ocean()); }
int orc_qcomp_atmos(Model *m) { ORC_TRY(m->qcomp_atmos()); }
int orc_helmholtz(Model *m, int which, double *wrk, const double *b) {
  ORC_TRY(if (which == 0) { if (m->cyclic) m->hscyoc(wrk, b); else m->hsbxoc(wrk, b); } else m->hscyat(wrk, b));
}
int orc_xforc(Model *m) { ORC_TRY(m->xforc()); }
int orc_oml(Model *m) { ORC_TRY(m->oml()); }
int orc_qgostep(Model *m) { ORC_TRY(m->qgostep()); }
int orc_ocinvq(Model *m) { ORC_TRY(m->ocinvq()); }
int orc_ocqbdy(Model *m) { ORC_TRY(m->ocqbdy(m->qo.data(), m->po.data())); }
int orc_aml(Model *m) { ORC_TRY(m->aml()); }
int orc_qgastep(Model *m) { ORC_TRY(m->qgastep()); }
int orc_atinvq(Model *m) { ORC_TRY(m->atinvq()); }
int orc_atqzbd(Model *m) { ORC_TRY(m->atqzbd(m->qa.data(), m->pa.data())); }
int orc_tlavg_ocean(Model *m) { ORC_TRY(m->tlavg_ocean()); }
int orc_tlavg_atmos(Model *m) { ORC_TRY(m->tlavg_atmos()); }
int orc_ocean_step(Model *m) {
  ORC_TRY(m->oml(); m->qgostep(); m->ocinvq(); m->ocqbdy(m->qo.data(), m->po.data()));
}
int orc_atmos_step(Model *m) {
  ORC_TRY(m->aml(); m->qgastep(); m->atinvq(); m->atqzbd(m->qa.data(), m->pa.data()));
}
int orc_run(Model *m, int64_t a, int64_t b) { ORC_TRY(m->run(a, b)); }
int orc_valids(Model *m, qgcm_valids_report *r) { ORC_TRY(m->valids(r)); }
int orc_tavini(Model *m) { ORC_TRY(m->tavini(3)); }
int orc_tavatm(Model *m) { ORC_TRY(m->tavatm()); }
int orc_tavocn(Model *m) { ORC_TRY(m->tavocn()); }
int orc_avg_ocn_k247(Model *m) { ORC_TRY(m->avg_ocn_k247()); }
int orc_qocdiag(Model *m, int32_t nsko, double *out, int64_t n) {
  ORC_TRY({
    int mw = m->nxpo % nsko;
    const int64_t iw = std::min(mw, 1) + (m->nxpo - mw) / nsko;
    mw = m->nypo % nsko;
    const int64_t jw = std::min(mw, 1) + (m->nypo - mw) / nsko;
    if (n != 5 * iw * jw * m->nlo) throw std::runtime_error("orc_qocdiag: element count mismatch");
    m->qocdiag(nsko, out);
  });
}
int orc_monnc_ocean(Model *m, qgcm_monitor_ocean *r) { ORC_TRY(m->monnc_ocean(r)); }
int orc_monnc_atmos(Model *m, qgcm_monitor_atmos *r) { ORC_TRY(m->monnc_atmos(r)); }
int orc_tav_counts(Model *m, int32_t *nsumat, int32_t *nsumoc, int32_t *nsum_ocavg) {
  *nsumat = m->nsumat;
  *nsumoc = m->nsumoc;
  *nsum_ocavg = m->nsum_ocavg;
  return 0;
}

// transform primitives, for pinning against scipy.fft (tests/test_oracle_fft.py)
int orc_rfftf(int n, double *r) {
  ORC_TRY(orc::FftPlan p; p.init(n); orc::vec s(2 * (size_t)n + 4); orc::rfftf(p, r, s.data()));
}
int orc_rfftb(int n, double *r) {
  ORC_TRY(orc::FftPlan p; p.init(n); orc::vec s(2 * (size_t)n + 4); orc::rfftb(p, r, s.data()));
}
// x holds n-1 data points followed by one scratch element (n doubles in total)
int orc_dsint(int n, double *x) {
  ORC_TRY(orc::FftPlan p; p.init(n); orc::vec s(2 * (size_t)n + 4); orc::dsint(p, x, s.data()));
}
double orc_xintp(const double *v, int nxp, int nyp) { return Model::xintp(v, nxp, nyp); }
double orc_xintt(const double *v, int nxt, int nyt) { return Model::xintt(v, nxt, nyt); }
}

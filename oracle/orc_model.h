// TEST INFRASTRUCTURE ONLY -- CPU oracle for the Q-GCM hot path.  Never linked into or imported
// by the product.  Pinned: the reference ships no golden vectors and no Fortran compiler exists
// here, so the reference's own sources are translated statement by statement to C++
// (oracle/f2cpp.py -> oracle/_ref) and this restatement is compared with that translation after
// every procedure call on box, channel and coupled decks (tests/test_reference_pin.py: 1e-17 ..
// 1e-14; see oracle/README.md for what the translation does and does not cover).
//
// Loop-for-loop C++ restatement of the reference's Fortran step routines, same loop
// order and expression association, compiled with -ffp-contract=off.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/qgcm_b200.h"
#include "orc_fft.h"

namespace orc {

typedef std::vector<double> vec;

struct Model {
  qgcm_config c;
  bool ocean_only, atmos_only, cyclic, sb_hflux, nb_hflux, tau_udiff;
  // grid (src/parameters_data.F:78-88)
  int nxto, nyto, nxpo, nypo, nlo;
  int nxta, nyta, nxpa, nypa, nla;
  int ndxr, nx1, ny1, nstr;
  int nxtaor, nytaor, nxpaor, nypaor;
  // derived constants (src/q-gcm.F:377-452)
  double fnot, beta;
  double dxo, dyo, hdxom1, dxom2, xlo, ylo, rdxof0, rrcpoc, tdto, dto, ocnorm;
  double dxa, dya, hdxam1, dxam2, xla, yla, rdxaf0, rrcpat, tdta, dta, atnorm, raoro;
  vec ypo, yporel, yto, ytorel, ypa, yparel, yta, ytarel;
  // Helmholtz (src/q-gcm.F:929-973)
  double aoc, aat;
  vec bd2oc, bd2at;
  FftPlan planoc, planat;
  // ocean state (src/ocstate_data.F:39-55, src/intrfac_data.F:39-48)
  vec po, pom, qo, qom, wekpo, wekto, entoc, ddynoc;
  vec sst, sstm, sstbar, tauxo, tauyo, fnetoc;
  // ocean homogeneous solutions (src/ochomog_data.F:44-55)
  vec ochom, pch1oc, pch2oc, pbhoc;
  // atmosphere state (src/atstate_data.F:37-40)
  vec pa, pam, qa, qam, wekpa, wekta, entat, ddynat, dtopat, xc1ast;
  vec ast, astm, astbar, hmixa, hmixam, tauxa, tauya, fnetat, uekat, vekat;
  vec pch1at, pch2at, pbhat;
  // xforc module storage (src/xfosubs.F:43-46)
  vec stbbb, stbus, stbun, stbvs, stbvn;  // bicubic weights (bcuini)
  bool bcu_ready = false;
  // running sums of src/timavge.F:46-85 (allocated by tavini)
  vec txatav, tyatav, wtatav, fmatav, astav, patav, qatav, uufa, tufa, utufa, vvfa, tvfa, vtvfa;
  vec txocav, tyocav, wpocav, wtocav, fmocav, sstav, pocav, qocav, uufo, tufo, utufo, vvfo, tvfo, vtvfo;
  vec po_avg;
  int nsumat = 0, nsumoc = 0, nsum_ocavg = 0;
  qgcm_scalars s;
  // work arrays of the step routines.  The reference declares them as automatic arrays
  // (e.g. dqdt(nxpo,nypo,nlo), src/qgosubs.F:65): stack storage that is neither zeroed nor
  // re-mapped from call to call.  A fresh std::vector per call would add a serial zero-fill and
  // a page fault per 4 KB of every temporary to each CPU step, so they persist here instead.
  vec wk_del2p, wk_dqdt, wk_d4p, wk_wrk, wk_rhs, wk_xfo, wk_del2t;
  vec wk_u1ator, wk_v1ator, wk_tauxaor, wk_tauyaor, wk_wektaor, wk_asto;      // xforc, src/xfosubs.F:100-118
  // ORC_POISON=1 fills them with NaN at every use: no routine may rely on what a previous
  // call left behind (tests/test_golden_fingerprints.py runs the fingerprints that way too)
  static vec &work(vec &v, size_t n) {
    if (v.size() != n) v.resize(n);
    static const bool poison = std::getenv("ORC_POISON") != nullptr;
    if (poison) std::fill(v.begin(), v.end(), std::nan(""));
    return v;
  }

  explicit Model(const qgcm_config &cfg);
  vec *field(const std::string &name);

  // src/intsubs.f
  static double xintt(const double *v, int nxt, int nyt);
  static double xintp(const double *v, int nxp, int nyp);
  // src/vorsubs.F
  void qcomp(double *q, const double *p, const double *aaa, const double *yprel, double dxm2,
             int nxp, int nyp, int nl, const double *ddyn, int kbot) const;
  void merqcy(double *q, const double *p, const double *aaa, const double *yprel, double dxm2,
              int nxp, int nyp, int nl, const double *ddyn, int kbot) const;
  void ocqbdy(double *q, const double *p);
  void atqzbd(double *q, const double *p);
  // src/qgosubs.F
  void qgostep();
  void ocadif(double *dqdt, const double *d2p, double ah2ock, double ah4ock, double bcfaco,
              const double *p, const double *q, double adfaco, int k);
  // src/ocisubs.F
  void ocinvq();
  void hsbxoc(double *wrk, const double *boc);
  void hscyoc(double *wrk, const double *boc);
  // src/omlsubs.F
  void oml();
  void omladf(double *rhs, const double *po1);
  // src/xfosubs.F
  void xforc();
  void xforc_ocean_ekman();
  // src/conhoms.F
  void constr();
  void homsol();
  // src/q-gcm.F:1328-1407
  void tlavg_ocean();
  void tlavg_atmos();
  // src/qgasubs.F, src/atisubs.F, src/amlsubs.F
  void qgastep();
  void atadif(double *dqdt, const double *d2p, double ah4atk, double bcfaat, const double *p,
              const double *q, double adfaca, int k);
  void atinvq();
  void hscyat(double *wrk, const double *bat);
  void aml();
  void amladf(double *rhsat, double *rhshm, const double *pa1);
  // src/q-gcm.F:719-749
  void qcomp_ocean();
  void qcomp_atmos();
  // src/valsubs.F
  void valids(qgcm_valids_report *rep);
  // src/timavge.F
  void tavini(int which = 3);
  void tavatm();
  void tavocn();
  void avg_ocn_k247();
  // src/qocdiag.F:303-683
  void qocdiag(int nsko, double *out);
  // src/monitor_diag.F:480-840
  void monnc_ocean(qgcm_monitor_ocean *rep);
  void couroc(qgcm_monitor_ocean *rep);
  void monnc_atmos(qgcm_monitor_atmos *rep);
  void courat(qgcm_monitor_atmos *rep);
  void run(int64_t nt_first, int64_t nt_last);
};

}  // namespace orc

// TEST INFRASTRUCTURE ONLY -- CPU oracle (pinned against the translated reference, see orc_model.h).
// Ocean routines: restatement of src/qgosubs.F, src/ocisubs.F, src/omlsubs.F,
// src/vorsubs.F (qcomp, merqcy, ocqbdy), src/intsubs.f, src/conhoms.F (ocean parts),
// src/xfosubs.F:568-709, src/q-gcm.F:1328-1366.  All indices below are 1-based
// through the accessor macros so the loops read like the Fortran they follow.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "orc_model.h"

namespace orc {

static const double PI = 3.14159265358979324;
static const double TWOPI = 6.28318530717958648;

#define IX2(i, j, nx) ((size_t)((i)-1) + (size_t)(nx) * (size_t)((j)-1))
#define IX3(i, j, k, nx, ny) ((size_t)((i)-1) + (size_t)(nx) * ((size_t)((j)-1) + (size_t)(ny) * (size_t)((k)-1)))
#define MAT(a, i, j, ld) (a)[((i)-1) + (size_t)(ld) * ((j)-1)]

static inline double sign(double a, double b) { return b >= 0.0 ? std::fabs(a) : -std::fabs(a); }

Model::Model(const qgcm_config &cfg) : c(cfg) {
  if (cfg.abi_version != QGCM_ABI_VERSION || cfg.struct_bytes != (int)sizeof(qgcm_config))
    throw std::runtime_error("orc::Model: qgcm_config ABI mismatch");
  std::memset(&s, 0, sizeof(s));
  ocean_only = cfg.flags & QGCM_OCEAN_ONLY;
  atmos_only = cfg.flags & QGCM_ATMOS_ONLY;
  cyclic = cfg.flags & QGCM_CYCLIC_OCEAN;
  sb_hflux = cfg.flags & QGCM_SB_HFLUX;
  nb_hflux = cfg.flags & QGCM_NB_HFLUX;
  tau_udiff = cfg.flags & QGCM_TAU_UDIFF;
  nxto = cfg.nxto; nyto = cfg.nyto; nlo = cfg.nlo;
  nxta = cfg.nxta; nyta = cfg.nyta; nla = cfg.nla;
  nxpo = nxto + 1; nypo = nyto + 1; nxpa = nxta + 1; nypa = nyta + 1;
  ndxr = cfg.ndxr; nx1 = cfg.nx1; ny1 = cfg.ny1; nstr = cfg.nstr;
  nxtaor = nxta * ndxr; nytaor = nyta * ndxr; nxpaor = nxtaor + 1; nypaor = nytaor + 1;
  fnot = cfg.fnot; beta = cfg.beta;
  // src/q-gcm.F:377-441
  dxo = cfg.dxo; dta = cfg.dta;
  dxa = ndxr * dxo;
  dto = nstr * dta;
  dya = dxa; hdxam1 = 0.5 / dxa; dxam2 = 1.0 / (dxa * dxa);
  xla = nxta * dxa; yla = nyta * dya;
  ypa.resize(nypa); yparel.resize(nypa); yta.resize(nyta); ytarel.resize(nyta);
  for (int j = 1; j <= nypa; ++j) { ypa[j - 1] = (j - 1) * dya; yparel[j - 1] = ypa[j - 1] - 0.5 * yla; }
  for (int j = 1; j <= nyta; ++j) { yta[j - 1] = ypa[j - 1] + 0.5 * dya; ytarel[j - 1] = yta[j - 1] - 0.5 * yla; }
  dyo = dxo; hdxom1 = 0.5 / dxo; dxom2 = 1.0 / (dxo * dxo);
  xlo = nxto * dxo; ylo = nyto * dyo;
  ypo.resize(nypo); yporel.resize(nypo); yto.resize(nyto); ytorel.resize(nyto);
  for (int j = 1; j <= nypo; ++j) { ypo[j - 1] = (ny1 - 1) * dya + (j - 1) * dyo; yporel[j - 1] = ypo[j - 1] - 0.5 * yla; }
  for (int j = 1; j <= nyto; ++j) { yto[j - 1] = ypo[j - 1] + 0.5 * dyo; ytorel[j - 1] = yto[j - 1] - 0.5 * yla; }
  rdxaf0 = 1.0 / (dxa * fnot);
  rdxof0 = 1.0 / (dxo * fnot);
  rrcpat = 1.0 / (cfg.rhoat * cfg.cpat);
  rrcpoc = 1.0 / (cfg.rhooc * cfg.cpoc);
  raoro = cfg.rhoat / cfg.rhooc;
  tdto = 2.0 * dto;
  tdta = 2.0 * dta;
  atnorm = 1.0 / ((double)nxta * nyta);
  ocnorm = 1.0 / ((double)nxto * nyto);

  const size_t np = (size_t)nxpo * nypo, nt = (size_t)nxto * nyto;
  if (!atmos_only) {
    po.assign(np * nlo, 0.0); pom = po; qo = po; qom = po;
    wekpo.assign(np, 0.0); entoc = wekpo; ddynoc = wekpo; tauxo = wekpo; tauyo = wekpo;
    wekto.assign(nt, 0.0); sst = wekto; sstm = wekto; fnetoc = wekto;
    sstbar.assign(nyto, 0.0);
    if (cyclic) {
      pch1oc.assign((size_t)nypo * (nlo - 1), 0.0); pch2oc = pch1oc; pbhoc.assign(nypo, 0.0);
    } else {
      ochom.assign(np * (nlo - 1), 0.0);
    }
    // src/q-gcm.F:929-953
    aoc = 1.0 / (dyo * dyo);
    bd2oc.assign(nxto, 0.0);
    if (cyclic) {
      for (int i = 2; i <= nxto / 2; ++i) {
        int i1 = 2 * i - 1;
        bd2oc[i1 - 2] = -2.0 * aoc + 2.0 * dxom2 * (std::cos((i - 1) * TWOPI / nxto) - 1.0);
        bd2oc[i1 - 1] = bd2oc[i1 - 2];
      }
      bd2oc[0] = -2.0 * aoc;
      bd2oc[nxto - 1] = -2.0 * aoc - 4.0 * dxom2;
    } else {
      for (int i = 2; i <= nxto; ++i)
        bd2oc[i - 2] = -2.0 * aoc + 2.0 * dxom2 * (std::cos((i - 1) * PI / nxto) - 1.0);
      bd2oc[nxto - 1] = 0.0;
    }
    planoc.init(nxto);
  }
  if (!ocean_only) {
    const size_t npa = (size_t)nxpa * nypa, nta = (size_t)nxta * nyta;
    pa.assign(npa * nla, 0.0); pam = pa; qa = pa; qam = pa;
    wekpa.assign(npa, 0.0); entat = wekpa; ddynat = wekpa; dtopat = wekpa; tauxa = wekpa; tauya = wekpa;
    wekta.assign(nta, 0.0); ast = wekta; astm = wekta; hmixa = wekta; hmixam = wekta;
    fnetat = wekta; xc1ast = wekta;
    astbar.assign(nyta, 0.0);
    uekat.assign((size_t)nxpa * nyta, 0.0); vekat.assign((size_t)nxta * nypa, 0.0);
    pch1at.assign((size_t)nypa * (nla - 1), 0.0); pch2at = pch1at; pbhat.assign(nypa, 0.0);
    // src/q-gcm.F:955-973
    aat = 1.0 / (dya * dya);
    bd2at.assign(nxta, 0.0);
    for (int i = 2; i <= nxta / 2; ++i) {
      int i1 = 2 * i - 1;
      bd2at[i1 - 2] = -2.0 * aat + 2.0 * dxam2 * (std::cos((i - 1) * TWOPI / nxta) - 1.0);
      bd2at[i1 - 1] = bd2at[i1 - 2];
    }
    bd2at[0] = -2.0 * aat;
    bd2at[nxta - 1] = -2.0 * aat - 4.0 * dxam2;
    planat.init(nxta);
  }
}

vec *Model::field(const std::string &n) {
#define F(x) if (n == #x) return &x;
  F(po) F(pom) F(qo) F(qom) F(wekpo) F(wekto) F(entoc) F(ddynoc) F(sst) F(sstm) F(sstbar)
  F(tauxo) F(tauyo) F(fnetoc) F(ochom) F(pch1oc) F(pch2oc) F(pbhoc)
  F(pa) F(pam) F(qa) F(qam) F(wekpa) F(wekta) F(entat) F(ddynat) F(dtopat) F(xc1ast)
  F(ast) F(astm) F(astbar) F(hmixa) F(hmixam) F(tauxa) F(tauya) F(fnetat) F(uekat) F(vekat)
  F(pch1at) F(pch2at) F(pbhat)
  F(txatav) F(tyatav) F(wtatav) F(fmatav) F(astav) F(patav) F(qatav) F(uufa) F(tufa) F(utufa) F(vvfa) F(tvfa) F(vtvfa)
  F(txocav) F(tyocav) F(wpocav) F(wtocav) F(fmocav) F(sstav) F(pocav) F(qocav) F(uufo) F(tufo) F(utufo) F(vvfo) F(tvfo)
  F(vtvfo) F(po_avg)
#undef F
  return nullptr;
}

// ---------------------------------------------------------------- src/intsubs.f:40-74
double Model::xintt(const double *valt, int nxt, int nyt) {
  double sumt = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sumt)
  for (int j = 1; j <= nyt; ++j) {
    double sumi = 0.0;
    for (int i = 1; i <= nxt; ++i) sumi = sumi + valt[IX2(i, j, nxt)];
    sumt = sumt + sumi;
  }
  return sumt;
}

// ---------------------------------------------------------------- src/intsubs.f:78-133
double Model::xintp(const double *valp, int nxp, int nyp) {
  double sump = 0.0;
  double xxs = 0.5 * valp[IX2(1, 1, nxp)];
  double xxn = 0.5 * valp[IX2(1, nyp, nxp)];
#pragma omp parallel for schedule(static) reduction(+ : sump)
  for (int j = 2; j <= nyp - 1; ++j) {
    double sumi = 0.5 * valp[IX2(1, j, nxp)];
    for (int i = 2; i <= nxp - 1; ++i) sumi = sumi + valp[IX2(i, j, nxp)];
    sumi = sumi + 0.5 * valp[IX2(nxp, j, nxp)];
    sump = sump + sumi;
  }
  for (int i = 2; i <= nxp - 1; ++i) {
    xxs = xxs + valp[IX2(i, 1, nxp)];
    xxn = xxn + valp[IX2(i, nyp, nxp)];
  }
  xxs = xxs + 0.5 * valp[IX2(nxp, 1, nxp)];
  xxn = xxn + 0.5 * valp[IX2(nxp, nyp, nxp)];
  return sump + 0.5 * (xxs + xxn);
}

// ---------------------------------------------------------------- src/vorsubs.F:49-135
void Model::qcomp(double *q, const double *p, const double *aaa, const double *yprel, double dxm2,
                  int nxp, int nyp, int nl, const double *ddyn, int kbot) const {
  const double dx2fac = dxm2 / fnot;
#define Q(i, j, k) q[IX3(i, j, k, nxp, nyp)]
#define P(i, j, k) p[IX3(i, j, k, nxp, nyp)]
#define A(i, j) MAT(aaa, i, j, nl)
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= nyp - 1; ++j) {
    const double betay = beta * yprel[j - 1];
    for (int i = 2; i <= nxp - 1; ++i)
      Q(i, j, 1) = dx2fac * (P(i, j - 1, 1) + P(i - 1, j, 1) + P(i + 1, j, 1) + P(i, j + 1, 1) - 4.0 * P(i, j, 1)) + betay -
                   fnot * (A(1, 1) * P(i, j, 1) + A(1, 2) * P(i, j, 2));
    for (int k = 2; k <= nl - 1; ++k)
      for (int i = 2; i <= nxp - 1; ++i)
        Q(i, j, k) = dx2fac * (P(i, j - 1, k) + P(i - 1, j, k) + P(i + 1, j, k) + P(i, j + 1, k) - 4.0 * P(i, j, k)) + betay -
                     fnot * (A(k, k - 1) * P(i, j, k - 1) + A(k, k) * P(i, j, k) + A(k, k + 1) * P(i, j, k + 1));
    for (int i = 2; i <= nxp - 1; ++i)
      Q(i, j, nl) = dx2fac * (P(i, j - 1, nl) + P(i - 1, j, nl) + P(i + 1, j, nl) + P(i, j + 1, nl) - 4.0 * P(i, j, nl)) + betay -
                    fnot * (A(nl, nl - 1) * P(i, j, nl - 1) + A(nl, nl) * P(i, j, nl));
    for (int i = 2; i <= nxp - 1; ++i) Q(i, j, kbot) = Q(i, j, kbot) + ddyn[IX2(i, j, nxp)];
  }
}

// ---------------------------------------------------------------- src/vorsubs.F:142-236
void Model::merqcy(double *q, const double *p, const double *aaa, const double *yprel, double dxm2,
                   int nxp, int nyp, int nl, const double *ddyn, int kbot) const {
  const double dx2fac = dxm2 / fnot;
  for (int j = 2; j <= nyp - 1; ++j) {
    const double betay = beta * yprel[j - 1];
    Q(1, j, 1) = dx2fac * (P(1, j - 1, 1) + P(nxp - 1, j, 1) + P(2, j, 1) + P(1, j + 1, 1) - 4.0 * P(1, j, 1)) + betay -
                 fnot * (A(1, 1) * P(1, j, 1) + A(1, 2) * P(1, j, 2));
    Q(nxp, j, 1) = Q(1, j, 1);
    for (int k = 2; k <= nl - 1; ++k) {
      Q(1, j, k) = dx2fac * (P(1, j - 1, k) + P(nxp - 1, j, k) + P(2, j, k) + P(1, j + 1, k) - 4.0 * P(1, j, k)) + betay -
                   fnot * (A(k, k - 1) * P(1, j, k - 1) + A(k, k) * P(1, j, k) + A(k, k + 1) * P(1, j, k + 1));
      Q(nxp, j, k) = Q(1, j, k);
    }
    Q(1, j, nl) = dx2fac * (P(1, j - 1, nl) + P(nxp - 1, j, nl) + P(2, j, nl) + P(1, j + 1, nl) - 4.0 * P(1, j, nl)) + betay -
                  fnot * (A(nl, nl - 1) * P(1, j, nl - 1) + A(nl, nl) * P(1, j, nl));
    Q(nxp, j, nl) = Q(1, j, nl);
    Q(1, j, kbot) = Q(1, j, kbot) + ddyn[IX2(1, j, nxp)];
    Q(nxp, j, kbot) = Q(1, j, kbot);
  }
#undef Q
#undef P
#undef A
}

// ---------------------------------------------------------------- src/vorsubs.F:245-388
void Model::ocqbdy(double *q, const double *p) {
  const double bcfaco = c.bccooc * dxom2 / (0.5 * c.bccooc + 1.0) / fnot;
  const double betays = beta * yporel[0];
  const double betayn = beta * yporel[nypo - 1];
#define Q(i, j, k) q[IX3(i, j, k, nxpo, nypo)]
#define P(i, j, k) p[IX3(i, j, k, nxpo, nypo)]
#define A(i, j) MAT(c.amatoc, i, j, nlo)
  double f0Am, f0Ac, f0Ap;
  f0Ac = fnot * A(1, 1);
  f0Ap = fnot * A(1, 2);
  for (int i = 1; i <= nxpo; ++i) {
    Q(i, 1, 1) = bcfaco * (P(i, 2, 1) - P(i, 1, 1)) - (f0Ac * P(i, 1, 1) + f0Ap * P(i, 1, 2)) + betays;
    Q(i, nypo, 1) = bcfaco * (P(i, nypo - 1, 1) - P(i, nypo, 1)) - (f0Ac * P(i, nypo, 1) + f0Ap * P(i, nypo, 2)) + betayn;
  }
  for (int k = 2; k <= nlo - 1; ++k) {
    f0Am = fnot * A(k, k - 1); f0Ac = fnot * A(k, k); f0Ap = fnot * A(k, k + 1);
    for (int i = 1; i <= nxpo; ++i) {
      Q(i, 1, k) = bcfaco * (P(i, 2, k) - P(i, 1, k)) - (f0Am * P(i, 1, k - 1) + f0Ac * P(i, 1, k) + f0Ap * P(i, 1, k + 1)) + betays;
      Q(i, nypo, k) = bcfaco * (P(i, nypo - 1, k) - P(i, nypo, k)) -
                      (f0Am * P(i, nypo, k - 1) + f0Ac * P(i, nypo, k) + f0Ap * P(i, nypo, k + 1)) + betayn;
    }
  }
  f0Am = fnot * A(nlo, nlo - 1);
  f0Ac = fnot * A(nlo, nlo);
  for (int i = 1; i <= nxpo; ++i) {
    Q(i, 1, nlo) = bcfaco * (P(i, 2, nlo) - P(i, 1, nlo)) - (f0Am * P(i, 1, nlo - 1) + f0Ac * P(i, 1, nlo)) + betays +
                   ddynoc[IX2(i, 1, nxpo)];
    Q(i, nypo, nlo) = bcfaco * (P(i, nypo - 1, nlo) - P(i, nypo, nlo)) -
                      (f0Am * P(i, nypo, nlo - 1) + f0Ac * P(i, nypo, nlo)) + betayn + ddynoc[IX2(i, nypo, nxpo)];
  }
  if (!cyclic) {
    f0Ac = fnot * A(1, 1);
    f0Ap = fnot * A(1, 2);
    for (int j = 2; j <= nypo - 1; ++j) {
      const double betay = beta * yporel[j - 1];
      Q(1, j, 1) = bcfaco * (P(2, j, 1) - P(1, j, 1)) - (f0Ac * P(1, j, 1) + f0Ap * P(1, j, 2)) + betay;
      Q(nxpo, j, 1) = bcfaco * (P(nxpo - 1, j, 1) - P(nxpo, j, 1)) - (f0Ac * P(nxpo, j, 1) + f0Ap * P(nxpo, j, 2)) + betay;
    }
    for (int k = 2; k <= nlo - 1; ++k) {
      f0Am = fnot * A(k, k - 1); f0Ac = fnot * A(k, k); f0Ap = fnot * A(k, k + 1);
      for (int j = 2; j <= nypo - 1; ++j) {
        const double betay = beta * yporel[j - 1];
        Q(1, j, k) = bcfaco * (P(2, j, k) - P(1, j, k)) - (f0Am * P(1, j, k - 1) + f0Ac * P(1, j, k) + f0Ap * P(1, j, k + 1)) + betay;
        Q(nxpo, j, k) = bcfaco * (P(nxpo - 1, j, k) - P(nxpo, j, k)) -
                        (f0Am * P(nxpo, j, k - 1) + f0Ac * P(nxpo, j, k) + f0Ap * P(nxpo, j, k + 1)) + betay;
      }
    }
    f0Am = fnot * A(nlo, nlo - 1);
    f0Ac = fnot * A(nlo, nlo);
    for (int j = 2; j <= nypo - 1; ++j) {
      const double betay = beta * yporel[j - 1];
      Q(1, j, nlo) = bcfaco * (P(2, j, nlo) - P(1, j, nlo)) - (f0Am * P(1, j, nlo - 1) + f0Ac * P(1, j, nlo)) + betay +
                     ddynoc[IX2(1, j, nxpo)];
      Q(nxpo, j, nlo) = bcfaco * (P(nxpo - 1, j, nlo) - P(nxpo, j, nlo)) -
                        (f0Am * P(nxpo, j, nlo - 1) + f0Ac * P(nxpo, j, nlo)) + betay + ddynoc[IX2(nxpo, j, nxpo)];
    }
  }
#undef Q
#undef P
#undef A
}

// ---------------------------------------------------------------- src/q-gcm.F:719-732
void Model::qcomp_ocean() {
  qcomp(qo.data(), po.data(), c.amatoc, yporel.data(), dxom2, nxpo, nypo, nlo, ddynoc.data(), nlo);
  qcomp(qom.data(), pom.data(), c.amatoc, yporel.data(), dxom2, nxpo, nypo, nlo, ddynoc.data(), nlo);
  ocqbdy(qo.data(), po.data());
  ocqbdy(qom.data(), pom.data());
  if (cyclic) {
    merqcy(qo.data(), po.data(), c.amatoc, yporel.data(), dxom2, nxpo, nypo, nlo, ddynoc.data(), nlo);
    merqcy(qom.data(), pom.data(), c.amatoc, yporel.data(), dxom2, nxpo, nypo, nlo, ddynoc.data(), nlo);
  }
}

// ---------------------------------------------------------------- src/qgosubs.F:45-221
void Model::qgostep() {
  const size_t np = (size_t)nxpo * nypo;
  vec &del2p = work(wk_del2p, np), &dqdt = work(wk_dqdt, np * nlo);
  const double adfaco = 1.0 / (12.0 * dxo * dyo * fnot);
  const double bcfaco = c.bccooc * dxom2 / (0.5 * c.bccooc + 1.0);
  double fohfac[QGCM_NLMAX];
  for (int k = 1; k <= nlo; ++k) fohfac[k - 1] = fnot / c.hoc[k - 1];
  const double bdrfac = 0.5 * sign(1.0, fnot) * c.delek / c.hoc[nlo - 1];
#define POM(i, j, k) pom[IX3(i, j, k, nxpo, nypo)]
#define D2(i, j) del2p[IX2(i, j, nxpo)]
  for (int k = 1; k <= nlo; ++k) {
    for (int i = 1; i <= nxpo; ++i) {
      D2(i, 1) = bcfaco * (POM(i, 2, k) - POM(i, 1, k));
      D2(i, nypo) = bcfaco * (POM(i, nypo - 1, k) - POM(i, nypo, k));
    }
#pragma omp parallel for schedule(static)
    for (int j = 2; j <= nypo - 1; ++j) {
      if (cyclic)
        D2(1, j) = (POM(1, j - 1, k) + POM(nxpo - 1, j, k) + POM(2, j, k) + POM(1, j + 1, k) - 4.0 * POM(1, j, k)) * dxom2;
      else
        D2(1, j) = bcfaco * (POM(2, j, k) - POM(1, j, k));
      for (int i = 2; i <= nxpo - 1; ++i)
        D2(i, j) = (POM(i, j - 1, k) + POM(i - 1, j, k) + POM(i + 1, j, k) + POM(i, j + 1, k) - 4.0 * POM(i, j, k)) * dxom2;
      if (cyclic)
        D2(nxpo, j) = D2(1, j);
      else
        D2(nxpo, j) = bcfaco * (POM(nxpo - 1, j, k) - POM(nxpo, j, k));
    }
    ocadif(&dqdt[np * (k - 1)], del2p.data(), c.ah2oc[k - 1], c.ah4oc[k - 1], bcfaco, &po[np * (k - 1)],
           &qo[np * (k - 1)], adfaco, k);
  }
  if (cyclic) {
    double bdsums = 0.0, bdsumn = 0.0;
    for (int i = 1; i <= nxpo - 1; ++i) {
      bdsums = bdsums + (POM(i, 2, nlo) - POM(i, 1, nlo));
      bdsumn = bdsumn + (POM(i, nypo, nlo) - POM(i, nypo - 1, nlo));
    }
    s.bdrins = 0.5 * sign(1.0, fnot) * c.delek * bdsums;
    s.bdrinn = 0.5 * sign(1.0, fnot) * c.delek * bdsumn;
  }
#define DQ(i, j, k) dqdt[IX3(i, j, k, nxpo, nypo)]
#define QO(i, j, k) qo[IX3(i, j, k, nxpo, nypo)]
#define QOM(i, j, k) qom[IX3(i, j, k, nxpo, nypo)]
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= nypo - 1; ++j) {
    double qdot[QGCM_NLMAX];
    for (int i = 1; i <= nxpo; ++i) {
      qdot[0] = DQ(i, j, 1) + fohfac[0] * (wekpo[IX2(i, j, nxpo)] - entoc[IX2(i, j, nxpo)]);
      qdot[1] = DQ(i, j, 2) + fohfac[1] * entoc[IX2(i, j, nxpo)];
      for (int k = 3; k <= nlo; ++k) qdot[k - 1] = DQ(i, j, k);
      qdot[nlo - 1] = qdot[nlo - 1] - bdrfac * D2(i, j);
      for (int k = 1; k <= nlo; ++k) {
        const double qold = QO(i, j, k);
        QO(i, j, k) = QOM(i, j, k) + tdto * qdot[k - 1];
        QOM(i, j, k) = qold;
      }
    }
  }
  for (int k = 1; k <= nlo; ++k)
    for (int i = 1; i <= nxpo; ++i) {
      QOM(i, 1, k) = QO(i, 1, k);
      QOM(i, nypo, k) = QO(i, nypo, k);
    }
#undef POM
#undef D2
#undef DQ
#undef QO
#undef QOM
}

// ---------------------------------------------------------------- src/qgosubs.F:231-446
void Model::ocadif(double *dqdt, const double *d2p, double ah2ock, double ah4ock, double bcfaco,
                   const double *p, const double *q, double adfaco, int k) {
  const size_t np = (size_t)nxpo * nypo;
  vec &d4p = work(wk_d4p, np);
  const double ah2fac = ah2ock / fnot;
  const double ah4fac = ah4ock / fnot;
#define D2(i, j) d2p[IX2(i, j, nxpo)]
#define D4(i, j) d4p[IX2(i, j, nxpo)]
#define P(i, j) p[IX2(i, j, nxpo)]
#define Q(i, j) q[IX2(i, j, nxpo)]
#define DQ(i, j) dqdt[IX2(i, j, nxpo)]
  if (cyclic) {
    double aj5sms = 0.5 * Q(1, 1) * (P(2, 2) - P(nxpo - 1, 2));
    double aj9sms = 0.5 * Q(1, 2) * (P(2, 2) - P(nxpo - 1, 2));
    for (int i = 2; i <= nxpo - 1; ++i) {
      aj5sms = aj5sms + Q(i, 1) * (P(i + 1, 2) - P(i - 1, 2));
      aj9sms = aj9sms + Q(i, 2) * (P(i + 1, 2) - P(i - 1, 2));
    }
    aj5sms = aj5sms + 0.5 * Q(nxpo, 1) * (P(2, 2) - P(nxpo - 1, 2));
    aj9sms = aj9sms + 0.5 * Q(nxpo, 2) * (P(2, 2) - P(nxpo - 1, 2));
    double ajis = fnot * adfaco * (aj5sms + 2.0 * aj9sms);
    s.ajisoc[k - 1] = dxo * dyo * ajis;
  }
  for (int i = 1; i <= nxpo; ++i) {
    D4(i, 1) = bcfaco * (D2(i, 2) - D2(i, 1));
    D4(i, nypo) = bcfaco * (D2(i, nypo - 1) - D2(i, nypo));
  }
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= nypo - 1; ++j) {
    if (cyclic)
      D4(1, j) = (D2(1, j - 1) + D2(nxpo - 1, j) + D2(2, j) + D2(1, j + 1) - 4.0 * D2(1, j)) * dxom2;
    else
      D4(1, j) = bcfaco * (D2(2, j) - D2(1, j));
    for (int i = 2; i <= nxpo - 1; ++i)
      D4(i, j) = dxom2 * (D2(i, j - 1) + D2(i - 1, j) + D2(i + 1, j) + D2(i, j + 1) - 4.0 * D2(i, j));
    if (cyclic)
      D4(nxpo, j) = D4(1, j);
    else
      D4(nxpo, j) = bcfaco * (D2(nxpo - 1, j) - D2(nxpo, j));
  }
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= nypo - 1; ++j) {
    if (cyclic) {
      const double d6p = dxom2 * (D4(1, j - 1) + D4(nxpo - 1, j) + D4(2, j) + D4(1, j + 1) - 4.0 * D4(1, j));
      const double diffus = ah2fac * D4(1, j) - ah4fac * d6p;
      DQ(1, j) = adfaco * ((Q(2, j) - Q(nxpo - 1, j)) * (P(1, j + 1) - P(1, j - 1)) +
                           (Q(1, j - 1) - Q(1, j + 1)) * (P(2, j) - P(nxpo - 1, j)) +
                           Q(2, j) * (P(2, j + 1) - P(2, j - 1)) -
                           Q(nxpo - 1, j) * (P(nxpo - 1, j + 1) - P(nxpo - 1, j - 1)) -
                           Q(1, j + 1) * (P(2, j + 1) - P(nxpo - 1, j + 1)) +
                           Q(1, j - 1) * (P(2, j - 1) - P(nxpo - 1, j - 1)) +
                           P(1, j + 1) * (Q(2, j + 1) - Q(nxpo - 1, j + 1)) -
                           P(1, j - 1) * (Q(2, j - 1) - Q(nxpo - 1, j - 1)) -
                           P(2, j) * (Q(2, j + 1) - Q(2, j - 1)) +
                           P(nxpo - 1, j) * (Q(nxpo - 1, j + 1) - Q(nxpo - 1, j - 1))) +
                 diffus;
    } else {
      DQ(1, j) = 0.0;
    }
    for (int i = 2; i <= nxpo - 1; ++i) {
      const double d6p = dxom2 * (D4(i, j - 1) + D4(i - 1, j) + D4(i + 1, j) + D4(i, j + 1) - 4.0 * D4(i, j));
      const double diffus = ah2fac * D4(i, j) - ah4fac * d6p;
      DQ(i, j) = adfaco * ((Q(i + 1, j) - Q(i - 1, j)) * (P(i, j + 1) - P(i, j - 1)) +
                           (Q(i, j - 1) - Q(i, j + 1)) * (P(i + 1, j) - P(i - 1, j)) +
                           Q(i + 1, j) * (P(i + 1, j + 1) - P(i + 1, j - 1)) -
                           Q(i - 1, j) * (P(i - 1, j + 1) - P(i - 1, j - 1)) -
                           Q(i, j + 1) * (P(i + 1, j + 1) - P(i - 1, j + 1)) +
                           Q(i, j - 1) * (P(i + 1, j - 1) - P(i - 1, j - 1)) +
                           P(i, j + 1) * (Q(i + 1, j + 1) - Q(i - 1, j + 1)) -
                           P(i, j - 1) * (Q(i + 1, j - 1) - Q(i - 1, j - 1)) -
                           P(i + 1, j) * (Q(i + 1, j + 1) - Q(i + 1, j - 1)) +
                           P(i - 1, j) * (Q(i - 1, j + 1) - Q(i - 1, j - 1))) +
                 diffus;
    }
    if (cyclic)
      DQ(nxpo, j) = DQ(1, j);
    else
      DQ(nxpo, j) = 0.0;
  }
  if (cyclic) {
    double aj5smn = -0.5 * Q(1, nypo) * (P(2, nypo - 1) - P(nxpo - 1, nypo - 1));
    double aj9smn = -0.5 * Q(1, nypo - 1) * (P(2, nypo - 1) - P(nxpo - 1, nypo - 1));
    for (int i = 2; i <= nxpo - 1; ++i) {
      aj5smn = aj5smn - Q(i, nypo) * (P(i + 1, nypo - 1) - P(i - 1, nypo - 1));
      aj9smn = aj9smn - Q(i, nypo - 1) * (P(i + 1, nypo - 1) - P(i - 1, nypo - 1));
    }
    aj5smn = aj5smn - 0.5 * Q(nxpo, nypo) * (P(2, nypo - 1) - P(nxpo - 1, nypo - 1));
    aj9smn = aj9smn - 0.5 * Q(nxpo, nypo - 1) * (P(2, nypo - 1) - P(nxpo - 1, nypo - 1));
    double ajin = fnot * adfaco * (aj5smn + 2.0 * aj9smn);
    s.ajinoc[k - 1] = dxo * dyo * ajin;
    double ah3sms = 0.0, ah3smn = 0.0, ah5sms = 0.0, ah5smn = 0.0;
    for (int i = 1; i <= nxpo - 1; ++i) {
      ah3sms = ah3sms + (D2(i, 2) - D2(i, 1));
      ah3smn = ah3smn + (D2(i, nypo) - D2(i, nypo - 1));
      ah5sms = ah5sms + (D4(i, 2) - D4(i, 1));
      ah5smn = ah5smn + (D4(i, nypo) - D4(i, nypo - 1));
    }
    s.ap3soc[k - 1] = ah2ock * ah3sms;
    s.ap3noc[k - 1] = ah2ock * ah3smn;
    s.ap5soc[k - 1] = ah4ock * ah5sms;
    s.ap5noc[k - 1] = ah4ock * ah5smn;
  }
#undef D2
#undef D4
#undef P
#undef Q
#undef DQ
}

// Thomas solve shared by hsbxoc/hscyoc/hscyat (src/ocisubs.F:470-488, :575-593)
static inline void thomas_col(double *wrk, int nxp, int nyp, int i, double bi, double a, double ftnorm,
                              double *gam, double *uvec) {
#define W(i, j) wrk[IX2(i, j, nxp)]
  double betinv = 1.0 / bi;
  uvec[2] = W(i, 2) * betinv;
  for (int j = 3; j <= nyp - 1; ++j) {
    gam[j] = a * betinv;
    betinv = 1.0 / (bi - a * gam[j]);
    uvec[j] = (W(i, j) - a * uvec[j - 1]) * betinv;
  }
  for (int j = nyp - 2; j >= 2; --j) uvec[j] = uvec[j] - gam[j + 1] * uvec[j + 1];
  for (int j = 2; j <= nyp - 1; ++j) W(i, j) = ftnorm * uvec[j];
#undef W
}

// ---------------------------------------------------------------- src/ocisubs.F:415-512
void Model::hsbxoc(double *wrk, const double *boc) {
  const double ftnorm = 0.5 / nxto;
#pragma omp parallel
  {
    vec scratch(2 * (size_t)nxto + 4), gam(nypo + 1), uvec(nypo + 1);
#pragma omp for schedule(static)
    for (int j = 2; j <= nypo - 1; ++j) dsint(planoc, &wrk[IX2(2, j, nxpo)], scratch.data());
#pragma omp for schedule(static)
    for (int i = 2; i <= nxpo - 1; ++i) thomas_col(wrk, nxpo, nypo, i, boc[i - 2], aoc, ftnorm, gam.data(), uvec.data());
#pragma omp for schedule(static)
    for (int j = 2; j <= nypo - 1; ++j) {
      dsint(planoc, &wrk[IX2(2, j, nxpo)], scratch.data());
      wrk[IX2(1, j, nxpo)] = 0.0;
      wrk[IX2(nxpo, j, nxpo)] = 0.0;
    }
  }
  for (int i = 1; i <= nxpo; ++i) {
    wrk[IX2(i, 1, nxpo)] = 0.0;
    wrk[IX2(i, nypo, nxpo)] = 0.0;
  }
}

// ---------------------------------------------------------------- src/ocisubs.F:521-618
void Model::hscyoc(double *wrk, const double *boc) {
  const double ftnorm = 1.0 / nxto;
#pragma omp parallel
  {
    vec scratch(2 * (size_t)nxto + 4), gam(nypo + 1), uvec(nypo + 1);
#pragma omp for schedule(static)
    for (int j = 2; j <= nypo - 1; ++j) rfftf(planoc, &wrk[IX2(1, j, nxpo)], scratch.data());
#pragma omp for schedule(static)
    for (int i = 1; i <= nxto; ++i) thomas_col(wrk, nxpo, nypo, i, boc[i - 1], aoc, ftnorm, gam.data(), uvec.data());
#pragma omp for schedule(static)
    for (int j = 2; j <= nypo - 1; ++j) {
      rfftb(planoc, &wrk[IX2(1, j, nxpo)], scratch.data());
      wrk[IX2(nxpo, j, nxpo)] = wrk[IX2(1, j, nxpo)];
    }
  }
  for (int i = 1; i <= nxpo; ++i) {
    wrk[IX2(i, 1, nxpo)] = 0.0;
    wrk[IX2(i, nypo, nxpo)] = 0.0;
  }
}

// 2x2.. (nlo-1)x(nlo-1) dense solve with partial pivoting + one sweep of iterative
// refinement: what DGETRS + DGERFS do at src/ocisubs.F:359-370 (LAPACK is not vendored
// in the reference and no version is pinned; this restates the published algorithm).
static void lu_solve_refine(const double *a, int n, const double *rhs, double *x) {
  double lu[QGCM_NLMAX * QGCM_NLMAX];
  int piv[QGCM_NLMAX];
  for (int i = 0; i < n * n; ++i) lu[i] = a[i];
  for (int kk = 0; kk < n; ++kk) {
    int p = kk;
    for (int i = kk + 1; i < n; ++i)
      if (std::fabs(lu[i + n * kk]) > std::fabs(lu[p + n * kk])) p = i;
    piv[kk] = p;
    if (p != kk)
      for (int j = 0; j < n; ++j) std::swap(lu[kk + n * j], lu[p + n * j]);
    for (int i = kk + 1; i < n; ++i) {
      lu[i + n * kk] /= lu[kk + n * kk];
      for (int j = kk + 1; j < n; ++j) lu[i + n * j] -= lu[i + n * kk] * lu[kk + n * j];
    }
  }
  auto solve = [&](double *b) {
    for (int kk = 0; kk < n; ++kk) std::swap(b[kk], b[piv[kk]]);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) b[i] -= lu[i + n * j] * b[j];
    for (int i = n - 1; i >= 0; --i) {
      for (int j = i + 1; j < n; ++j) b[i] -= lu[i + n * j] * b[j];
      b[i] /= lu[i + n * i];
    }
  };
  for (int i = 0; i < n; ++i) x[i] = rhs[i];
  solve(x);
  double r[QGCM_NLMAX];
  for (int i = 0; i < n; ++i) {
    double acc = rhs[i];
    for (int j = 0; j < n; ++j) acc -= a[i + n * j] * x[j];
    r[i] = acc;
  }
  solve(r);
  for (int i = 0; i < n; ++i) x[i] += r[i];
}

// ---------------------------------------------------------------- src/ocisubs.F:64-407
void Model::ocinvq() {
  const double ecrito = 1.0e-13;
  const size_t np = (size_t)nxpo * nypo;
  vec &wrk = work(wk_wrk, np * nlo);
  double xinhom[QGCM_NLMAX];
#define WRK(i, j, m) wrk[IX3(i, j, m, nxpo, nypo)]
#define QO(i, j, k) qo[IX3(i, j, k, nxpo, nypo)]
#define PO(i, j, k) po[IX3(i, j, k, nxpo, nypo)]
#define POM(i, j, k) pom[IX3(i, j, k, nxpo, nypo)]
#define CTL2M(k, m) MAT(c.ctl2moc, k, m, nlo)
#define CTM2L(m, k) MAT(c.ctm2loc, m, k, nlo)
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= nypo - 1; ++j) {
    const double betay = beta * yporel[j - 1];
    double ql[QGCM_NLMAX];
    for (int i = 1; i <= nxpo; ++i) {
      for (int k = 1; k <= nlo; ++k) ql[k - 1] = QO(i, j, k) - betay;
      ql[nlo - 1] = ql[nlo - 1] - ddynoc[IX2(i, j, nxpo)];
      for (int m = 1; m <= nlo; ++m) {
        double qm = 0.0;
        for (int k = 1; k <= nlo; ++k) qm = qm + CTL2M(k, m) * ql[k - 1];
        WRK(i, j, m) = fnot * qm;
      }
    }
  }
  vec boc(nxto);
  for (int m = 1; m <= nlo; ++m) {
    for (int i = 1; i <= nxto; ++i) boc[i - 1] = bd2oc[i - 1] - c.rdm2oc[m - 1];
    if (cyclic)
      hscyoc(&wrk[np * (m - 1)], boc.data());
    else
      hsbxoc(&wrk[np * (m - 1)], boc.data());
    xinhom[m - 1] = xintp(&wrk[np * (m - 1)], nxpo, nypo);
    xinhom[m - 1] = xinhom[m - 1] * dxo * dyo;
    s.xinhom_oc[m - 1] = xinhom[m - 1];
  }
  if (cyclic) {
    double rhss[QGCM_NLMAX], rhsn[QGCM_NLMAX], ocsnew[QGCM_NLMAX], ocnnew[QGCM_NLMAX];
    double clhss[QGCM_NLMAX], clhsn[QGCM_NLMAX], c1[QGCM_NLMAX], c2[QGCM_NLMAX], c3;
    double aipmod[QGCM_NLMAX], aiplay[QGCM_NLMAX];
    const double entfac = 0.5 * dyo * fnot * fnot;
    const double *hoc = c.hoc;
    rhss[0] = (entfac / hoc[0]) * s.enisoc[0] + (fnot / hoc[0]) * s.txisoc + s.ajisoc[0] - s.ap3soc[0] + s.ap5soc[0];
    rhsn[0] = (entfac / hoc[0]) * s.eninoc[0] - (fnot / hoc[0]) * s.txinoc + s.ajinoc[0] + s.ap3noc[0] - s.ap5noc[0];
    for (int k = 2; k <= nlo - 1; ++k) {
      rhss[k - 1] = (entfac / hoc[k - 1]) * (s.enisoc[k - 1] - s.enisoc[k - 2]) + s.ajisoc[k - 1] - s.ap3soc[k - 1] + s.ap5soc[k - 1];
      rhsn[k - 1] = (entfac / hoc[k - 1]) * (s.eninoc[k - 1] - s.eninoc[k - 2]) + s.ajinoc[k - 1] + s.ap3noc[k - 1] - s.ap5noc[k - 1];
    }
    rhss[nlo - 1] = -(entfac / hoc[nlo - 1]) * s.enisoc[nlo - 2] + s.ajisoc[nlo - 1] - s.ap3soc[nlo - 1] + s.ap5soc[nlo - 1] +
                    (fnot / hoc[nlo - 1]) * s.bdrins;
    rhsn[nlo - 1] = -(entfac / hoc[nlo - 1]) * s.eninoc[nlo - 2] + s.ajinoc[nlo - 1] + s.ap3noc[nlo - 1] - s.ap5noc[nlo - 1] -
                    (fnot / hoc[nlo - 1]) * s.bdrinn;
    for (int k = 1; k <= nlo; ++k) {
      ocsnew[k - 1] = s.ocncsp[k - 1] + tdto * rhss[k - 1];
      ocnnew[k - 1] = s.ocncnp[k - 1] + tdto * rhsn[k - 1];
      s.ocncsp[k - 1] = s.ocncs[k - 1];
      s.ocncnp[k - 1] = s.ocncn[k - 1];
      s.ocncs[k - 1] = ocsnew[k - 1];
      s.ocncn[k - 1] = ocnnew[k - 1];
    }
    for (int m = 1; m <= nlo; ++m) {
      double ayis = 0.5 * WRK(1, 2, m);
      double ayin = -0.5 * WRK(1, nypo - 1, m);
      for (int i = 2; i <= nxpo - 1; ++i) {
        ayis = ayis + WRK(i, 2, m);
        ayin = ayin - WRK(i, nypo - 1, m);
      }
      ayis = ayis + 0.5 * WRK(nxpo, 2, m);
      ayin = ayin - 0.5 * WRK(nxpo, nypo - 1, m);
      ayis = ayis * (dxo / dyo);
      ayin = ayin * (dxo / dyo);
      clhss[m - 1] = 0.0;
      clhsn[m - 1] = 0.0;
      for (int k = 1; k <= nlo; ++k) {
        clhss[m - 1] = clhss[m - 1] + CTL2M(k, m) * ocsnew[k - 1];
        clhsn[m - 1] = clhsn[m - 1] + CTL2M(k, m) * ocnnew[k - 1];
      }
      clhss[m - 1] = clhss[m - 1] + ayis;
      clhsn[m - 1] = clhsn[m - 1] - ayin;
    }
    c3 = clhss[0] * s.hbsioc;
    for (int m = 1; m <= nlo - 1; ++m) {
      c1[m - 1] = s.hc2noc[m - 1] * clhss[m] - s.hc2soc[m - 1] * clhsn[m];
      c2[m - 1] = s.hc1soc[m - 1] * clhsn[m] - s.hc1noc[m - 1] * clhss[m];
    }
    aipmod[0] = xinhom[0] + c3 * s.aipbho;
    for (int m = 2; m <= nlo; ++m) aipmod[m - 1] = xinhom[m - 1] + (c1[m - 2] + c2[m - 2]) * s.aipcho[m - 2];
    for (int k = 1; k <= nlo; ++k) {
      double pl = 0.0;
      for (int m = 1; m <= nlo; ++m) pl = pl + CTM2L(m, k) * aipmod[m - 1];
      aiplay[k - 1] = pl;
    }
    for (int k = 1; k <= nlo - 1; ++k) {
      const double est1 = aiplay[k] - aiplay[k - 1];
      const double est2 = s.dpiocp[k - 1] - tdto * c.gpoc[k - 1] * s.xon[k - 1];
      const double edif = est1 - est2;
      const double esum = std::fabs(est1) + std::fabs(est2);
      s.ermaso[k - 1] = edif;
      if (esum > (ecrito * xlo * ylo * tdto * c.gpoc[k - 1]))
        s.emfroc[k - 1] = 2.0 * edif / esum;
      else
        s.emfroc[k - 1] = 0.0;
      s.dpiocp[k - 1] = s.dpioc[k - 1];
      s.dpioc[k - 1] = aiplay[k] - aiplay[k - 1];
    }
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= nypo; ++j) {
      double homcor[QGCM_NLMAX], pm[QGCM_NLMAX];
      homcor[0] = c3 * pbhoc[j - 1];
      for (int m = 2; m <= nlo; ++m)
        homcor[m - 1] = c1[m - 2] * pch1oc[IX2(j, m - 1, nypo)] + c2[m - 2] * pch2oc[IX2(j, m - 1, nypo)];
      for (int i = 1; i <= nxpo; ++i) {
        for (int m = 1; m <= nlo; ++m) pm[m - 1] = WRK(i, j, m) + homcor[m - 1];
        for (int k = 1; k <= nlo; ++k) {
          POM(i, j, k) = PO(i, j, k);
          double pl = 0.0;
          for (int m = 1; m <= nlo; ++m) pl = pl + CTM2L(m, k) * pm[m - 1];
          PO(i, j, k) = pl;
        }
      }
    }
  } else {
    double aient[QGCM_NLMAX], rhs[QGCM_NLMAX], hclco[QGCM_NLMAX];
    aient[0] = s.xon[0];
    for (int k = 2; k <= nlo - 1; ++k) aient[k - 1] = 0.0;
    for (int k = 1; k <= nlo - 1; ++k) {
      const double aitmp = s.dpioc[k - 1];
      s.dpioc[k - 1] = s.dpiocp[k - 1] - tdto * c.gpoc[k - 1] * aient[k - 1];
      s.dpiocp[k - 1] = aitmp;
      double rhsum = 0.0;
      for (int m = 1; m <= nlo; ++m) rhsum = rhsum + MAT(s.cdiffo, m, k, nlo) * xinhom[m - 1];
      rhs[k - 1] = s.dpioc[k - 1] - rhsum;
    }
    lu_solve_refine(s.cdhoc, nlo - 1, rhs, hclco);
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= nypo; ++j) {
      double pm[QGCM_NLMAX];
      for (int i = 1; i <= nxpo; ++i) {
        pm[0] = WRK(i, j, 1);
        for (int m = 2; m <= nlo; ++m) pm[m - 1] = WRK(i, j, m) + hclco[m - 2] * ochom[IX3(i, j, m - 1, nxpo, nypo)];
        for (int k = 1; k <= nlo; ++k) {
          POM(i, j, k) = PO(i, j, k);
          double pl = 0.0;
          for (int m = 1; m <= nlo; ++m) pl = pl + CTM2L(m, k) * pm[m - 1];
          PO(i, j, k) = pl;
        }
      }
    }
  }
#undef WRK
#undef QO
#undef PO
#undef POM
}

// ---------------------------------------------------------------- src/omlsubs.F:47-236
void Model::oml() {
  const size_t nt = (size_t)nxto * nyto;
  vec &rhs = work(wk_rhs, nt), &xfo = work(wk_xfo, nt);
  const double hmoinv = 1.0 / c.hmoc;
  const double dtoinv = 1.0 / (c.toc[0] - c.toc[1]);
  const double entfac = c.hmoc * dtoinv / tdto;
  omladf(rhs.data(), po.data());
  double xfosum = 0.0, cfrasm = 0.0, centsm = 0.0;
  const double toc1 = c.toc[0];
#define T2(a, i, j) a[IX2(i, j, nxto)]
#pragma omp parallel for schedule(static) reduction(+ : cfrasm) reduction(- : centsm)
  for (int j = 1; j <= nyto; ++j) {
    for (int i = 1; i <= nxto; ++i) {
      const double diabat = 0.5 * T2(wekto, i, j) * (T2(sstm, i, j) + toc1);
      double sstnew = T2(sstm, i, j) + tdto * (T2(rhs, i, j) + hmoinv * (rrcpoc * T2(fnetoc, i, j) + diabat));
      const double xfoent = -(0.5 * dtoinv) * T2(wekto, i, j) * (T2(sstm, i, j) - toc1);
      const double dtonew = toc1 - sstnew;
      const double coneno = entfac * std::max(0.0, dtonew);
      T2(xfo, i, j) = xfoent - coneno;
      sstnew = sstnew + std::max(0.0, dtonew);
      cfrasm = cfrasm + (0.5 - sign(0.5, -dtonew));
      centsm = centsm - coneno;
      T2(sstm, i, j) = T2(sst, i, j);
      T2(sst, i, j) = sstnew;
    }
  }
#pragma omp parallel for schedule(static) reduction(+ : xfosum)
  for (int j = 1; j <= nyto; ++j) {
    double xfsi = 0.0;
    for (int i = 1; i <= nxto; ++i) xfsi = xfsi + T2(xfo, i, j);
    xfosum = xfosum + xfsi;
  }
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= nyto; ++j)
    for (int i = 1; i <= nxto; ++i) T2(xfo, i, j) = T2(xfo, i, j) - xfosum * ocnorm;
#define EN(i, j) entoc[IX2(i, j, nxpo)]
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= nypo - 1; ++j)
    for (int i = 2; i <= nxpo - 1; ++i)
      EN(i, j) = 0.25 * (T2(xfo, i - 1, j - 1) + T2(xfo, i, j - 1) + T2(xfo, i - 1, j) + T2(xfo, i, j));
  for (int i = 2; i <= nxpo - 1; ++i) {
    EN(i, 1) = 0.5 * (T2(xfo, i - 1, 1) + T2(xfo, i, 1));
    EN(i, nypo) = 0.5 * (T2(xfo, i - 1, nyto) + T2(xfo, i, nyto));
  }
  if (cyclic) {
    for (int j = 2; j <= nypo - 1; ++j) {
      EN(1, j) = 0.25 * (T2(xfo, nxto, j - 1) + T2(xfo, 1, j - 1) + T2(xfo, nxto, j) + T2(xfo, 1, j));
      EN(nxpo, j) = EN(1, j);
    }
    EN(1, 1) = 0.5 * (T2(xfo, nxto, 1) + T2(xfo, 1, 1));
    EN(1, nypo) = 0.5 * (T2(xfo, nxto, nyto) + T2(xfo, 1, nyto));
    EN(nxpo, 1) = EN(1, 1);
    EN(nxpo, nypo) = EN(1, nypo);
  } else {
    for (int j = 2; j <= nypo - 1; ++j) {
      EN(1, j) = 0.5 * (T2(xfo, 1, j - 1) + T2(xfo, 1, j));
      EN(nxpo, j) = 0.5 * (T2(xfo, nxto, j - 1) + T2(xfo, nxto, j));
    }
    EN(1, 1) = T2(xfo, 1, 1);
    EN(nxpo, 1) = T2(xfo, nxto, 1);
    EN(1, nypo) = T2(xfo, 1, nyto);
    EN(nxpo, nypo) = T2(xfo, nxto, nyto);
  }
  s.cfraoc = cfrasm * ocnorm;
  s.centoc = centsm * dxo * dyo;
  s.xon[0] = xintp(entoc.data(), nxpo, nypo);
  s.xon[0] = s.xon[0] * dxo * dyo;
  if (cyclic) {
    double ensums = 0.5 * EN(1, 1);
    double ensumn = 0.5 * EN(1, nypo);
    for (int i = 2; i <= nxpo - 1; ++i) {
      ensums = ensums + EN(i, 1);
      ensumn = ensumn + EN(i, nypo);
    }
    ensums = ensums + 0.5 * EN(nxpo, 1);
    ensumn = ensumn + 0.5 * EN(nxpo, nypo);
    s.enisoc[0] = dxo * ensums;
    s.eninoc[0] = dxo * ensumn;
  }
#undef EN
}

// ---------------------------------------------------------------- src/omlsubs.F:244-763
void Model::omladf(double *rhs, const double *po1) {
  const double uvgfac = c.ycexp * rdxof0;
  const double rhf0hm = 0.5 / (fnot * c.hmoc);
  const double d2tfac = c.st2d * dxom2;
  const double d4tfac = c.st4d * dxom2 * dxom2;
  const double tsbdy = c.tsbdy, tnbdy = c.tnbdy;
  const int nxd = nxto + 2;  // del2t(0:nxto+1,nyto)
  vec &del2t = work(wk_del2t, (size_t)nxd * nyto);
#define D2T(i, j) del2t[(size_t)(i) + (size_t)nxd * ((j)-1)]
#define PO1(i, j) po1[IX2(i, j, nxpo)]
#define TX(i, j) tauxo[IX2(i, j, nxpo)]
#define TY(i, j) tauyo[IX2(i, j, nxpo)]
#define SST(i, j) sst[IX2(i, j, nxto)]
#define SSTM(i, j) sstm[IX2(i, j, nxto)]
#define RHS(i, j) rhs[IX2(i, j, nxto)]
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= nyto - 1; ++j) {
    double um, tm, up, tp, hxadv, vm, vp, hyadv;
    if (cyclic) {
      um = -uvgfac * (PO1(1, j + 1) - PO1(1, j)) + rhf0hm * (TY(1, j + 1) + TY(1, j));
      tm = SST(1, j) + SST(nxto, j);
      D2T(1, j) = SSTM(1, j - 1) + SSTM(nxto, j) + SSTM(2, j) + SSTM(1, j + 1) - 4.0 * SSTM(1, j);
    } else {
      um = 0.0;
      tm = 0.0;
      D2T(1, j) = SSTM(1, j - 1) + SSTM(2, j) + SSTM(1, j + 1) - 3.0 * SSTM(1, j);
    }
    up = -uvgfac * (PO1(2, j + 1) - PO1(2, j)) + rhf0hm * (TY(2, j + 1) + TY(2, j));
    tp = SST(1, j) + SST(2, j);
    hxadv = hdxom1 * (up * tp - um * tm);
    vm = uvgfac * (PO1(2, j) - PO1(1, j)) - rhf0hm * (TX(2, j) + TX(1, j));
    vp = uvgfac * (PO1(2, j + 1) - PO1(1, j + 1)) - rhf0hm * (TX(2, j + 1) + TX(1, j + 1));
    hyadv = hdxom1 * (vp * (SST(1, j + 1) + SST(1, j)) - vm * (SST(1, j) + SST(1, j - 1)));
    RHS(1, j) = -(hxadv + hyadv);
    for (int i = 2; i <= nxto - 1; ++i) {
      um = up;
      tm = tp;
      up = -uvgfac * (PO1(i + 1, j + 1) - PO1(i + 1, j)) + rhf0hm * (TY(i + 1, j + 1) + TY(i + 1, j));
      tp = SST(i, j) + SST(i + 1, j);
      hxadv = hdxom1 * (up * tp - um * tm);
      vm = uvgfac * (PO1(i + 1, j) - PO1(i, j)) - rhf0hm * (TX(i + 1, j) + TX(i, j));
      vp = uvgfac * (PO1(i + 1, j + 1) - PO1(i, j + 1)) - rhf0hm * (TX(i + 1, j + 1) + TX(i, j + 1));
      hyadv = hdxom1 * (vp * (SST(i, j + 1) + SST(i, j)) - vm * (SST(i, j) + SST(i, j - 1)));
      RHS(i, j) = -(hxadv + hyadv);
      D2T(i, j) = SSTM(i, j - 1) + SSTM(i - 1, j) + SSTM(i + 1, j) + SSTM(i, j + 1) - 4.0 * SSTM(i, j);
    }
    um = up;
    tm = tp;
    if (cyclic) {
      up = -uvgfac * (PO1(nxto + 1, j + 1) - PO1(nxto + 1, j)) + rhf0hm * (TY(nxto + 1, j + 1) + TY(nxto + 1, j));
      tp = SST(1, j) + SST(nxto, j);
      D2T(nxto, j) = SSTM(nxto, j - 1) + SSTM(nxto - 1, j) + SSTM(1, j) + SSTM(nxto, j + 1) - 4.0 * SSTM(nxto, j);
    } else {
      up = 0.0;
      tp = 0.0;
      D2T(nxto, j) = SSTM(nxto, j - 1) + SSTM(nxto - 1, j) + SSTM(nxto, j + 1) - 3.0 * SSTM(nxto, j);
    }
    hxadv = hdxom1 * (up * tp - um * tm);
    vm = uvgfac * (PO1(nxto + 1, j) - PO1(nxto, j)) - rhf0hm * (TX(nxto + 1, j) + TX(nxto, j));
    vp = uvgfac * (PO1(nxto + 1, j + 1) - PO1(nxto, j + 1)) - rhf0hm * (TX(nxto + 1, j + 1) + TX(nxto, j + 1));
    hyadv = hdxom1 * (vp * (SST(nxto, j + 1) + SST(nxto, j)) - vm * (SST(nxto, j) + SST(nxto, j - 1)));
    RHS(nxto, j) = -(hxadv + hyadv);
    if (cyclic) {
      D2T(0, j) = D2T(nxto, j);
      D2T(nxto + 1, j) = D2T(1, j);
    } else {
      D2T(0, j) = D2T(1, j);
      D2T(nxto + 1, j) = D2T(nxto, j);
    }
  }
  // zonal boundaries, inner points (src/omlsubs.F:391-456)
#pragma omp parallel for schedule(static)
  for (int i = 2; i <= nxto - 1; ++i) {
    double um, tm, up, tp, hxadv, vm, vp, hyadv;
    um = -uvgfac * (PO1(i, 2) - PO1(i, 1)) + rhf0hm * (TY(i, 2) + TY(i, 1));
    up = -uvgfac * (PO1(i + 1, 2) - PO1(i + 1, 1)) + rhf0hm * (TY(i + 1, 2) + TY(i + 1, 1));
    hxadv = hdxom1 * (up * (SST(i + 1, 1) + SST(i, 1)) - um * (SST(i, 1) + SST(i - 1, 1)));
    vp = uvgfac * (PO1(i + 1, 2) - PO1(i, 2)) - rhf0hm * (TX(i + 1, 2) + TX(i, 2));
    tp = SST(i, 1) + SST(i, 2);
    if (sb_hflux) {
      vm = -rhf0hm * (TX(i + 1, 1) + TX(i, 1));
      tm = SST(i, 1) + tsbdy;
      hyadv = hdxom1 * (vp * tp - vm * tm);
      D2T(i, 1) = SSTM(i - 1, 1) + SSTM(i + 1, 1) + SSTM(i, 2) + tsbdy - 4.0 * SSTM(i, 1);
    } else {
      hyadv = hdxom1 * (vp * tp);
      D2T(i, 1) = SSTM(i - 1, 1) + SSTM(i + 1, 1) + SSTM(i, 2) - 3.0 * SSTM(i, 1);
    }
    RHS(i, 1) = -(hxadv + hyadv);
    um = -uvgfac * (PO1(i, nyto + 1) - PO1(i, nyto)) + rhf0hm * (TY(i, nyto + 1) + TY(i, nyto));
    up = -uvgfac * (PO1(i + 1, nyto + 1) - PO1(i + 1, nyto)) + rhf0hm * (TY(i + 1, nyto + 1) + TY(i + 1, nyto));
    hxadv = hdxom1 * (up * (SST(i + 1, nyto) + SST(i, nyto)) - um * (SST(i, nyto) + SST(i - 1, nyto)));
    vm = uvgfac * (PO1(i + 1, nyto) - PO1(i, nyto)) - rhf0hm * (TX(i + 1, nyto) + TX(i, nyto));
    tm = SST(i, nyto - 1) + SST(i, nyto);
    if (nb_hflux) {
      vp = -rhf0hm * (TX(i + 1, nyto + 1) + TX(i, nyto + 1));
      tp = SST(i, nyto) + tnbdy;
      hyadv = hdxom1 * (vp * tp - vm * tm);
      D2T(i, nyto) = SSTM(i, nyto - 1) + SSTM(i - 1, nyto) + tnbdy + SSTM(i + 1, nyto) - 4.0 * SSTM(i, nyto);
    } else {
      hyadv = hdxom1 * (-vm * tm);
      D2T(i, nyto) = SSTM(i, nyto - 1) + SSTM(i - 1, nyto) + SSTM(i + 1, nyto) - 3.0 * SSTM(i, nyto);
    }
    RHS(i, nyto) = -(hxadv + hyadv);
  }
  // corners (src/omlsubs.F:462-682)
  {
    double um, tm, up, tp, hxadv, vm, vp, hyadv;
    // SW
    if (cyclic) {
      um = -uvgfac * (PO1(1, 2) - PO1(1, 1)) + rhf0hm * (TY(1, 2) + TY(1, 1));
      tm = SST(1, 1) + SST(nxto, 1);
      if (sb_hflux)
        D2T(1, 1) = SSTM(nxto, 1) + SSTM(2, 1) + SSTM(1, 2) + tsbdy - 4.0 * SSTM(1, 1);
      else
        D2T(1, 1) = SSTM(nxto, 1) + SSTM(2, 1) + SSTM(1, 2) - 3.0 * SSTM(1, 1);
      D2T(nxto + 1, 1) = D2T(1, 1);
    } else {
      um = 0.0;
      tm = 0.0;
      if (sb_hflux)
        D2T(1, 1) = SSTM(2, 1) + SSTM(1, 2) + tsbdy - 3.0 * SSTM(1, 1);
      else
        D2T(1, 1) = SSTM(2, 1) + SSTM(1, 2) - 2.0 * SSTM(1, 1);
      D2T(0, 1) = D2T(1, 1);
    }
    up = -uvgfac * (PO1(2, 2) - PO1(2, 1)) + rhf0hm * (TY(2, 2) + TY(2, 1));
    tp = SST(1, 1) + SST(2, 1);
    hxadv = hdxom1 * (up * tp - um * tm);
    vp = uvgfac * (PO1(2, 2) - PO1(1, 2)) - rhf0hm * (TX(2, 2) + TX(1, 2));
    tp = SST(1, 1) + SST(1, 2);
    if (sb_hflux) {
      vm = -rhf0hm * (TX(2, 1) + TX(1, 1));
      tm = SST(1, 1) + tsbdy;
      hyadv = hdxom1 * (vp * tp - vm * tm);
    } else {
      hyadv = hdxom1 * (vp * tp);
    }
    RHS(1, 1) = -(hxadv + hyadv);
    // SE
    um = -uvgfac * (PO1(nxto, 2) - PO1(nxto, 1)) + rhf0hm * (TY(nxto, 2) + TY(nxto, 1));
    tm = SST(nxto - 1, 1) + SST(nxto, 1);
    if (cyclic) {
      up = -uvgfac * (PO1(nxto + 1, 2) - PO1(nxto + 1, 1)) + rhf0hm * (TY(nxto + 1, 2) + TY(nxto + 1, 1));
      tp = SST(1, 1) + SST(nxto, 1);
      if (sb_hflux)
        D2T(nxto, 1) = SSTM(nxto - 1, 1) + SSTM(1, 1) + SSTM(nxto, 2) + tsbdy - 4.0 * SSTM(nxto, 1);
      else
        D2T(nxto, 1) = SSTM(nxto - 1, 1) + SSTM(1, 1) + SSTM(nxto, 2) - 3.0 * SSTM(nxto, 1);
      D2T(0, 1) = D2T(nxto, 1);
    } else {
      up = 0.0;
      tp = 0.0;
      if (sb_hflux)
        D2T(nxto, 1) = SSTM(nxto - 1, 1) + SSTM(nxto, 2) + tsbdy - 3.0 * SSTM(nxto, 1);
      else
        D2T(nxto, 1) = SSTM(nxto - 1, 1) + SSTM(nxto, 2) - 2.0 * SSTM(nxto, 1);
      D2T(nxto + 1, 1) = D2T(nxto, 1);
    }
    hxadv = hdxom1 * (up * tp - um * tm);
    vp = uvgfac * (PO1(nxto + 1, 2) - PO1(nxto, 2)) - rhf0hm * (TX(nxto + 1, 2) + TX(nxto, 2));
    tp = SST(nxto, 1) + SST(nxto, 2);
    if (sb_hflux) {
      vm = -rhf0hm * (TX(nxto + 1, 1) + TX(nxto, 1));
      tm = SST(nxto, 1) + tsbdy;
      hyadv = hdxom1 * (vp * tp - vm * tm);
    } else {
      hyadv = hdxom1 * (vp * tp);
    }
    RHS(nxto, 1) = -(hxadv + hyadv);
    // NW
    if (cyclic) {
      um = -uvgfac * (PO1(1, nyto + 1) - PO1(1, nyto)) + rhf0hm * (TY(1, nyto + 1) + TY(1, nyto));
      tm = SST(1, nyto) + SST(nxto, nyto);
      if (nb_hflux)
        D2T(1, nyto) = SSTM(1, nyto - 1) + SSTM(nxto, nyto) + tnbdy + SSTM(2, nyto) - 4.0 * SSTM(1, nyto);
      else
        D2T(1, nyto) = SSTM(1, nyto - 1) + SSTM(nxto, nyto) + SSTM(2, nyto) - 3.0 * SSTM(1, nyto);
      D2T(nxto + 1, nyto) = D2T(1, nyto);
    } else {
      um = 0.0;
      tm = 0.0;
      if (nb_hflux)
        D2T(1, nyto) = SSTM(1, nyto - 1) + tnbdy + SSTM(2, nyto) - 3.0 * SSTM(1, nyto);
      else
        D2T(1, nyto) = SSTM(1, nyto - 1) + SSTM(2, nyto) - 2.0 * SSTM(1, nyto);
      D2T(0, nyto) = D2T(1, nyto);
    }
    up = -uvgfac * (PO1(2, nyto + 1) - PO1(2, nyto)) + rhf0hm * (TY(2, nyto + 1) + TY(2, nyto));
    tp = SST(1, nyto) + SST(2, nyto);
    hxadv = hdxom1 * (up * tp - um * tm);
    vm = uvgfac * (PO1(2, nyto) - PO1(1, nyto)) - rhf0hm * (TX(2, nyto) + TX(1, nyto));
    tm = SST(1, nyto - 1) + SST(1, nyto);
    if (nb_hflux) {
      vp = -rhf0hm * (TX(2, nyto + 1) + TX(1, nyto + 1));
      tp = SST(1, nyto) + tnbdy;
      hyadv = hdxom1 * (vp * tp - vm * tm);
    } else {
      hyadv = hdxom1 * (-vm * tm);
    }
    RHS(1, nyto) = -(hxadv + hyadv);
    // NE
    um = -uvgfac * (PO1(nxto, nyto + 1) - PO1(nxto, nyto)) + rhf0hm * (TY(nxto, nyto + 1) + TY(nxto, nyto));
    tm = SST(nxto - 1, nyto) + SST(nxto, nyto);
    if (cyclic) {
      up = -uvgfac * (PO1(nxto + 1, nyto + 1) - PO1(nxto + 1, nyto)) + rhf0hm * (TY(nxto + 1, nyto + 1) + TY(nxto + 1, nyto));
      tp = SST(1, nyto) + SST(nxto, nyto);
      if (nb_hflux)
        D2T(nxto, nyto) = SSTM(nxto, nyto - 1) + SSTM(nxto - 1, nyto) + SSTM(1, nyto) - 4.0 * SSTM(nxto, nyto) + tnbdy;
      else
        D2T(nxto, nyto) = SSTM(nxto, nyto - 1) + SSTM(nxto - 1, nyto) + SSTM(1, nyto) - 3.0 * SSTM(nxto, nyto);
      D2T(0, nyto) = D2T(nxto, nyto);
    } else {
      up = 0.0;
      tp = 0.0;
      if (nb_hflux)
        D2T(nxto, nyto) = SSTM(nxto, nyto - 1) + SSTM(nxto - 1, nyto) + tnbdy - 3.0 * SSTM(nxto, nyto);
      else
        D2T(nxto, nyto) = SSTM(nxto, nyto - 1) + SSTM(nxto - 1, nyto) - 2.0 * SSTM(nxto, nyto);
      D2T(nxto + 1, nyto) = D2T(nxto, nyto);
    }
    hxadv = hdxom1 * (up * tp - um * tm);
    vm = uvgfac * (PO1(nxto + 1, nyto) - PO1(nxto, nyto)) - rhf0hm * (TX(nxto + 1, nyto) + TX(nxto, nyto));
    tm = SST(nxto, nyto) + SST(nxto, nyto - 1);
    if (nb_hflux) {
      vp = -rhf0hm * (TX(nxto + 1, nyto + 1) + TX(nxto, nyto + 1));
      tp = SST(nxto, nyto) + tnbdy;
      hyadv = hdxom1 * (vp * tp - vm * tm);
    } else {
      hyadv = hdxom1 * (-vm * tm);
    }
    RHS(nxto, nyto) = -(hxadv + hyadv);
  }
  // monitoring (src/omlsubs.F:686-726)
  double vfsmsb = 0.0, tasmsb = 0.0, tdsmsb = 0.0, vfsmnb = 0.0, tasmnb = 0.0, tdsmnb = 0.0;
  if (sb_hflux)
    for (int i = 1; i <= nxto; ++i) {
      const double vm = -rhf0hm * (TX(i + 1, 1) + TX(i, 1));
      const double tm = SST(i, 1) + tsbdy;
      vfsmsb = vfsmsb + vm;
      tasmsb = tasmsb + vm * tm;
      tdsmsb = tdsmsb - (SSTM(i, 1) - tsbdy);
    }
  if (nb_hflux)
    for (int i = 1; i <= nxto; ++i) {
      const double vp = -rhf0hm * (TX(i + 1, nyto + 1) + TX(i, nyto + 1));
      const double tp = SST(i, nyto) + tnbdy;
      vfsmnb = vfsmnb - vp;
      tasmnb = tasmnb - vp * tp;
      tdsmnb = tdsmnb + (tnbdy - SSTM(i, nyto));
    }
  s.ttmads = hdxom1 * tasmsb / (double)nxto;
  s.vfmads = vfsmsb / (double)nxto;
  s.ttmdfs = d2tfac * tdsmsb / (double)nxto;
  s.ttmadn = hdxom1 * tasmnb / (double)nxto;
  s.vfmadn = vfsmnb / (double)nxto;
  s.ttmdfn = d2tfac * tdsmnb / (double)nxto;
  // diffusion (src/omlsubs.F:728-758)
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= nyto - 1; ++j)
    for (int i = 1; i <= nxto; ++i)
      RHS(i, j) = RHS(i, j) + d2tfac * D2T(i, j) -
                  d4tfac * (D2T(i, j - 1) + D2T(i - 1, j) + D2T(i + 1, j) + D2T(i, j + 1) - 4.0 * D2T(i, j));
  for (int i = 1; i <= nxto; ++i) {
    RHS(i, 1) = RHS(i, 1) + d2tfac * D2T(i, 1) - d4tfac * (D2T(i - 1, 1) + D2T(i + 1, 1) + D2T(i, 2) - 3.0 * D2T(i, 1));
    RHS(i, nyto) = RHS(i, nyto) + d2tfac * D2T(i, nyto) -
                   d4tfac * (D2T(i, nyto - 1) + D2T(i - 1, nyto) + D2T(i + 1, nyto) - 3.0 * D2T(i, nyto));
  }
#undef D2T
#undef PO1
#undef TX
#undef TY
#undef SST
#undef SSTM
#undef RHS
#undef T2
}

// ---------------------------------------------------------------- src/xfosubs.F:568-683
void Model::xforc_ocean_ekman() {
  const double hxofac = 0.5 * rdxof0;  // src/xfosubs.F: hxofac = 0.5d0*rdxof0
#define TX(i, j) tauxo[IX2(i, j, nxpo)]
#define TY(i, j) tauyo[IX2(i, j, nxpo)]
#define WT(i, j) wekto[IX2(i, j, nxto)]
#define WP(i, j) wekpo[IX2(i, j, nxpo)]
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= nyto; ++j)
    for (int i = 1; i <= nxto; ++i)
      WT(i, j) = hxofac * (TY(i + 1, j + 1) + TY(i + 1, j) - (TY(i, j + 1) + TY(i, j)) + TX(i + 1, j) + TX(i, j) -
                           (TX(i + 1, j + 1) + TX(i, j + 1)));
#pragma omp parallel for schedule(static)
  for (int jo = 2; jo <= nypo - 1; ++jo) {
    if (cyclic)
      WP(1, jo) = 0.25 * (WT(nxto, jo - 1) + WT(nxto, jo) + WT(1, jo - 1) + WT(1, jo));
    else
      WP(1, jo) = 0.5 * (WT(1, jo - 1) + WT(1, jo));
    for (int io = 2; io <= nxpo - 1; ++io)
      WP(io, jo) = 0.25 * (WT(io - 1, jo - 1) + WT(io - 1, jo) + WT(io, jo - 1) + WT(io, jo));
    if (cyclic)
      WP(nxpo, jo) = WP(1, jo);
    else
      WP(nxpo, jo) = 0.5 * (WT(nxto, jo - 1) + WT(nxto, jo));
  }
  if (cyclic) {
    WP(1, 1) = 0.5 * (WT(nxto, 1) + WT(1, 1));
    WP(1, nypo) = 0.5 * (WT(nxto, nyto) + WT(1, nyto));
  } else {
    WP(1, 1) = WT(1, 1);
    WP(1, nypo) = WT(1, nyto);
  }
  for (int io = 2; io <= nxpo - 1; ++io) {
    WP(io, 1) = 0.5 * (WT(io - 1, 1) + WT(io, 1));
    WP(io, nypo) = 0.5 * (WT(io - 1, nyto) + WT(io, nyto));
  }
  if (cyclic) {
    WP(nxpo, 1) = WP(1, 1);
    WP(nxpo, nypo) = WP(1, nypo);
  } else {
    WP(nxpo, 1) = WT(nxto, 1);
    WP(nxpo, nypo) = WT(nxto, nyto);
  }
  if (cyclic) {
    double txsums = 0.5 * (TX(1, 1) + TX(1, 2));
    double txsumn = 0.5 * (TX(1, nypo - 1) + TX(1, nypo));
    for (int io = 2; io <= nxpo - 1; ++io) {
      txsums = txsums + (TX(io, 1) + TX(io, 2));
      txsumn = txsumn + (TX(io, nypo - 1) + TX(io, nypo));
    }
    txsums = txsums + 0.5 * (TX(nxpo, 1) + TX(nxpo, 2));
    txsumn = txsumn + 0.5 * (TX(nxpo, nypo - 1) + TX(nxpo, nypo));
    s.txisoc = 0.5 * dxo * txsums;
    s.txinoc = 0.5 * dxo * txsumn;
  }
#undef TX
#undef TY
#undef WT
#undef WP
}

// ---------------------------------------------------------------- src/q-gcm.F:1328-1366
void Model::tlavg_ocean() {
  const size_t n3 = (size_t)nxpo * nypo * nlo, nt = (size_t)nxto * nyto;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n3; ++i) {
    qo[i] = 0.5 * (qo[i] + qom[i]);
    po[i] = 0.5 * (po[i] + pom[i]);
  }
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nt; ++i) sst[i] = 0.5 * (sst[i] + sstm[i]);
  for (int k = 1; k <= nlo - 1; ++k) s.dpioc[k - 1] = 0.5 * (s.dpioc[k - 1] + s.dpiocp[k - 1]);
  if (cyclic)
    for (int k = 1; k <= nlo; ++k) {
      s.ocncs[k - 1] = 0.5 * (s.ocncs[k - 1] + s.ocncsp[k - 1]);
      s.ocncn[k - 1] = 0.5 * (s.ocncn[k - 1] + s.ocncnp[k - 1]);
    }
}

}  // namespace orc

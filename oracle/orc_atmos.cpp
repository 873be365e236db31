// TEST INFRASTRUCTURE ONLY -- CPU oracle (pinned against the translated reference, see orc_model.h).
// Atmosphere routines: restatement of src/qgasubs.F, src/atisubs.F, src/amlsubs.F,
// src/vorsubs.F:396-480 (atqzbd), src/q-gcm.F:738-749, :1370-1407.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "orc_model.h"

namespace orc {

#define IX2(i, j, nx) ((size_t)((i)-1) + (size_t)(nx) * (size_t)((j)-1))
#define IX3(i, j, k, nx, ny) ((size_t)((i)-1) + (size_t)(nx) * ((size_t)((j)-1) + (size_t)(ny) * (size_t)((k)-1)))
#define MAT(a, i, j, ld) (a)[((i)-1) + (size_t)(ld) * ((j)-1)]

// ---------------------------------------------------------------- src/atisubs.F:301-395
void Model::hscyat(double *wrk, const double *bat) {
  const double ftnorm = 1.0 / nxta;
  vec scratch(2 * (size_t)nxta + 4), gam(nypa + 1), uvec(nypa + 1);
#define W(i, j) wrk[IX2(i, j, nxpa)]
  for (int j = 2; j <= nypa - 1; ++j) rfftf(planat, &W(1, j), scratch.data());
  for (int i = 1; i <= nxta; ++i) {
    double betinv = 1.0 / bat[i - 1];
    uvec[2] = W(i, 2) * betinv;
    for (int j = 3; j <= nypa - 1; ++j) {
      gam[j] = aat * betinv;
      betinv = 1.0 / (bat[i - 1] - aat * gam[j]);
      uvec[j] = (W(i, j) - aat * uvec[j - 1]) * betinv;
    }
    for (int j = nypa - 2; j >= 2; --j) uvec[j] = uvec[j] - gam[j + 1] * uvec[j + 1];
    for (int j = 2; j <= nypa - 1; ++j) W(i, j) = ftnorm * uvec[j];
  }
  for (int j = 2; j <= nypa - 1; ++j) {
    rfftb(planat, &W(1, j), scratch.data());
    W(nxpa, j) = W(1, j);
  }
  for (int i = 1; i <= nxpa; ++i) {
    W(i, 1) = 0.0;
    W(i, nypa) = 0.0;
  }
#undef W
}

// ---------------------------------------------------------------- src/q-gcm.F:1370-1407
void Model::tlavg_atmos() {
  const size_t n3 = (size_t)nxpa * nypa * nla, nt = (size_t)nxta * nyta;
  for (size_t i = 0; i < n3; ++i) {
    qa[i] = 0.5 * (qa[i] + qam[i]);
    pa[i] = 0.5 * (pa[i] + pam[i]);
  }
  for (size_t i = 0; i < nt; ++i) {
    ast[i] = 0.5 * (ast[i] + astm[i]);
    hmixa[i] = 0.5 * (hmixa[i] + hmixam[i]);
  }
  for (int k = 1; k <= nla - 1; ++k) s.dpiat[k - 1] = 0.5 * (s.dpiat[k - 1] + s.dpiatp[k - 1]);
  for (int k = 1; k <= nla; ++k) {
    s.atmcs[k - 1] = 0.5 * (s.atmcs[k - 1] + s.atmcsp[k - 1]);
    s.atmcn[k - 1] = 0.5 * (s.atmcn[k - 1] + s.atmcnp[k - 1]);
  }
}

// ---------------------------------------------------------------- src/vorsubs.F:396-480
void Model::atqzbd(double *q, const double *p) {
  const double zbfaca = c.bccoat * dxam2 / (0.5 * c.bccoat + 1.0) / fnot;
  const double betays = beta * yparel[0];
  const double betayn = beta * yparel[nypa - 1];
#define Q(i, j, k) q[IX3(i, j, k, nxpa, nypa)]
#define P(i, j, k) p[IX3(i, j, k, nxpa, nypa)]
#define A(i, j) MAT(c.amatat, i, j, nla)
  double f0Am, f0Ac, f0Ap;
  f0Ac = fnot * A(1, 1);
  f0Ap = fnot * A(1, 2);
  for (int i = 1; i <= nxpa; ++i) {
    Q(i, 1, 1) = zbfaca * (P(i, 2, 1) - P(i, 1, 1)) - (f0Ac * P(i, 1, 1) + f0Ap * P(i, 1, 2)) + betays + ddynat[IX2(i, 1, nxpa)];
    Q(i, nypa, 1) = zbfaca * (P(i, nypa - 1, 1) - P(i, nypa, 1)) - (f0Ac * P(i, nypa, 1) + f0Ap * P(i, nypa, 2)) + betayn +
                    ddynat[IX2(i, nypa, nxpa)];
  }
  for (int k = 2; k <= nla - 1; ++k) {
    f0Am = fnot * A(k, k - 1); f0Ac = fnot * A(k, k); f0Ap = fnot * A(k, k + 1);
    for (int i = 1; i <= nxpa; ++i) {
      Q(i, 1, k) = zbfaca * (P(i, 2, k) - P(i, 1, k)) - (f0Am * P(i, 1, k - 1) + f0Ac * P(i, 1, k) + f0Ap * P(i, 1, k + 1)) + betays;
      Q(i, nypa, k) = zbfaca * (P(i, nypa - 1, k) - P(i, nypa, k)) -
                      (f0Am * P(i, nypa, k - 1) + f0Ac * P(i, nypa, k) + f0Ap * P(i, nypa, k + 1)) + betayn;
    }
  }
  f0Am = fnot * A(nla, nla - 1);
  f0Ac = fnot * A(nla, nla);
  for (int i = 1; i <= nxpa; ++i) {
    // reference quirk (src/vorsubs.F:470): the southern row uses pa(i,2,nla), not pa(i,1,nla)
    Q(i, 1, nla) = zbfaca * (P(i, 2, nla) - P(i, 1, nla)) - (f0Am * P(i, 1, nla - 1) + f0Ac * P(i, 2, nla)) + betays;
    Q(i, nypa, nla) = zbfaca * (P(i, nypa - 1, nla) - P(i, nypa, nla)) - (f0Am * P(i, nypa, nla - 1) + f0Ac * P(i, nypa, nla)) + betayn;
  }
#undef Q
#undef P
#undef A
}

// ---------------------------------------------------------------- src/q-gcm.F:738-749
void Model::qcomp_atmos() {
  qcomp(qa.data(), pa.data(), c.amatat, yparel.data(), dxam2, nxpa, nypa, nla, ddynat.data(), 1);
  qcomp(qam.data(), pam.data(), c.amatat, yparel.data(), dxam2, nxpa, nypa, nla, ddynat.data(), 1);
  atqzbd(qa.data(), pa.data());
  atqzbd(qam.data(), pam.data());
  merqcy(qa.data(), pa.data(), c.amatat, yparel.data(), dxam2, nxpa, nypa, nla, ddynat.data(), 1);
  merqcy(qam.data(), pam.data(), c.amatat, yparel.data(), dxam2, nxpa, nypa, nla, ddynat.data(), 1);
}


static inline double sign(double a, double b) { return b >= 0.0 ? std::fabs(a) : -std::fabs(a); }

// ---------------------------------------------------------------- src/qgasubs.F:45-148
void Model::qgastep() {
  const size_t np = (size_t)nxpa * nypa;
  vec del2p(np), dqdt(np * nla);
  const double adfaca = 1.0 / (12.0 * dxa * dya * fnot);
  const double zbfaca = c.bccoat * dxam2 / (0.5 * c.bccoat + 1.0);
  double fohfac[QGCM_NLMAX];
  for (int k = 1; k <= nla; ++k) fohfac[k - 1] = fnot / c.hat[k - 1];
#define PAM(i, j, k) pam[IX3(i, j, k, nxpa, nypa)]
#define D2(i, j) del2p[IX2(i, j, nxpa)]
  for (int k = 1; k <= nla; ++k) {
    for (int j = 2; j <= nypa - 1; ++j) {
      D2(1, j) = (PAM(1, j - 1, k) + PAM(nxpa - 1, j, k) + PAM(2, j, k) + PAM(1, j + 1, k) - 4.0 * PAM(1, j, k)) * dxam2;
      for (int i = 2; i <= nxpa - 1; ++i)
        D2(i, j) = (PAM(i, j - 1, k) + PAM(i - 1, j, k) + PAM(i + 1, j, k) + PAM(i, j + 1, k) - 4.0 * PAM(i, j, k)) * dxam2;
      D2(nxpa, j) = D2(1, j);
    }
    for (int i = 1; i <= nxpa; ++i) {
      D2(i, 1) = zbfaca * (PAM(i, 2, k) - PAM(i, 1, k));
      D2(i, nypa) = zbfaca * (PAM(i, nypa - 1, k) - PAM(i, nypa, k));
    }
    atadif(&dqdt[np * (k - 1)], del2p.data(), c.ah4at[k - 1], zbfaca, &pa[np * (k - 1)], &qa[np * (k - 1)], adfaca, k);
  }
#define DQ(i, j, k) dqdt[IX3(i, j, k, nxpa, nypa)]
#define QA(i, j, k) qa[IX3(i, j, k, nxpa, nypa)]
#define QAM(i, j, k) qam[IX3(i, j, k, nxpa, nypa)]
  for (int j = 2; j <= nypa - 1; ++j) {
    double qdot[QGCM_NLMAX];
    for (int i = 1; i <= nxpa; ++i) {
      qdot[0] = DQ(i, j, 1) + fohfac[0] * (entat[IX2(i, j, nxpa)] - wekpa[IX2(i, j, nxpa)]);
      qdot[1] = DQ(i, j, 2) - fohfac[1] * entat[IX2(i, j, nxpa)];
      for (int k = 3; k <= nla; ++k) qdot[k - 1] = DQ(i, j, k);
      for (int k = 1; k <= nla; ++k) {
        const double qold = QA(i, j, k);
        QA(i, j, k) = QAM(i, j, k) + tdta * qdot[k - 1];
        QAM(i, j, k) = qold;
      }
    }
  }
  for (int k = 1; k <= nla; ++k)
    for (int i = 1; i <= nxpa; ++i) {
      QAM(i, 1, k) = QA(i, 1, k);
      QAM(i, nypa, k) = QA(i, nypa, k);
    }
#undef PAM
#undef D2
#undef DQ
#undef QA
#undef QAM
}

// ---------------------------------------------------------------- src/qgasubs.F:156-317
void Model::atadif(double *dqdt, const double *d2p, double ah4atk, double zbfaca, const double *p,
                   const double *q, double adfaca, int k) {
  const size_t np = (size_t)nxpa * nypa;
  vec d4p(np);
  const double ah4fac = ah4atk / fnot;
#define D2(i, j) d2p[IX2(i, j, nxpa)]
#define D4(i, j) d4p[IX2(i, j, nxpa)]
#define P(i, j) p[IX2(i, j, nxpa)]
#define Q(i, j) q[IX2(i, j, nxpa)]
#define DQ(i, j) dqdt[IX2(i, j, nxpa)]
  double aj5sms = 0.5 * Q(1, 1) * (P(2, 2) - P(nxpa - 1, 2));
  double aj9sms = 0.5 * Q(1, 2) * (P(2, 2) - P(nxpa - 1, 2));
  for (int i = 2; i <= nxpa - 1; ++i) {
    aj5sms = aj5sms + Q(i, 1) * (P(i + 1, 2) - P(i - 1, 2));
    aj9sms = aj9sms + Q(i, 2) * (P(i + 1, 2) - P(i - 1, 2));
  }
  aj5sms = aj5sms + 0.5 * Q(nxpa, 1) * (P(2, 2) - P(nxpa - 1, 2));
  aj9sms = aj9sms + 0.5 * Q(nxpa, 2) * (P(2, 2) - P(nxpa - 1, 2));
  s.ajisat[k - 1] = dxa * dya * (fnot * adfaca * (aj5sms + 2.0 * aj9sms));
  for (int i = 1; i <= nxpa; ++i) {
    D4(i, 1) = zbfaca * (D2(i, 2) - D2(i, 1));
    D4(i, nypa) = zbfaca * (D2(i, nypa - 1) - D2(i, nypa));
  }
  for (int j = 2; j <= nypa - 1; ++j) {
    D4(1, j) = dxam2 * (D2(1, j - 1) + D2(nxpa - 1, j) + D2(2, j) + D2(1, j + 1) - 4.0 * D2(1, j));
    for (int i = 2; i <= nxpa - 1; ++i)
      D4(i, j) = dxam2 * (D2(i, j - 1) + D2(i - 1, j) + D2(i + 1, j) + D2(i, j + 1) - 4.0 * D2(i, j));
    D4(nxpa, j) = D4(1, j);
  }
  for (int j = 2; j <= nypa - 1; ++j) {
    double d6p = dxam2 * (D4(1, j - 1) + D4(nxpa - 1, j) + D4(2, j) + D4(1, j + 1) - 4.0 * D4(1, j));
    DQ(1, j) = adfaca * ((Q(2, j) - Q(nxpa - 1, j)) * (P(1, j + 1) - P(1, j - 1)) +
                         (Q(1, j - 1) - Q(1, j + 1)) * (P(2, j) - P(nxpa - 1, j)) +
                         Q(2, j) * (P(2, j + 1) - P(2, j - 1)) -
                         Q(nxpa - 1, j) * (P(nxpa - 1, j + 1) - P(nxpa - 1, j - 1)) -
                         Q(1, j + 1) * (P(2, j + 1) - P(nxpa - 1, j + 1)) +
                         Q(1, j - 1) * (P(2, j - 1) - P(nxpa - 1, j - 1)) +
                         P(1, j + 1) * (Q(2, j + 1) - Q(nxpa - 1, j + 1)) -
                         P(1, j - 1) * (Q(2, j - 1) - Q(nxpa - 1, j - 1)) -
                         P(2, j) * (Q(2, j + 1) - Q(2, j - 1)) +
                         P(nxpa - 1, j) * (Q(nxpa - 1, j + 1) - Q(nxpa - 1, j - 1))) -
               ah4fac * d6p;
    for (int i = 2; i <= nxpa - 1; ++i) {
      d6p = dxam2 * (D4(i, j - 1) + D4(i - 1, j) + D4(i + 1, j) + D4(i, j + 1) - 4.0 * D4(i, j));
      DQ(i, j) = adfaca * ((Q(i + 1, j) - Q(i - 1, j)) * (P(i, j + 1) - P(i, j - 1)) +
                           (Q(i, j - 1) - Q(i, j + 1)) * (P(i + 1, j) - P(i - 1, j)) +
                           Q(i + 1, j) * (P(i + 1, j + 1) - P(i + 1, j - 1)) -
                           Q(i - 1, j) * (P(i - 1, j + 1) - P(i - 1, j - 1)) -
                           Q(i, j + 1) * (P(i + 1, j + 1) - P(i - 1, j + 1)) +
                           Q(i, j - 1) * (P(i + 1, j - 1) - P(i - 1, j - 1)) +
                           P(i, j + 1) * (Q(i + 1, j + 1) - Q(i - 1, j + 1)) -
                           P(i, j - 1) * (Q(i + 1, j - 1) - Q(i - 1, j - 1)) -
                           P(i + 1, j) * (Q(i + 1, j + 1) - Q(i + 1, j - 1)) +
                           P(i - 1, j) * (Q(i - 1, j + 1) - Q(i - 1, j - 1))) -
                 ah4fac * d6p;
    }
    DQ(nxpa, j) = DQ(1, j);
  }
  double aj5smn = -0.5 * Q(1, nypa) * (P(2, nypa - 1) - P(nxpa - 1, nypa - 1));
  double aj9smn = -0.5 * Q(1, nypa - 1) * (P(2, nypa - 1) - P(nxpa - 1, nypa - 1));
  for (int i = 2; i <= nxpa - 1; ++i) {
    aj5smn = aj5smn - Q(i, nypa) * (P(i + 1, nypa - 1) - P(i - 1, nypa - 1));
    aj9smn = aj9smn - Q(i, nypa - 1) * (P(i + 1, nypa - 1) - P(i - 1, nypa - 1));
  }
  aj5smn = aj5smn - 0.5 * Q(nxpa, nypa) * (P(2, nypa - 1) - P(nxpa - 1, nypa - 1));
  aj9smn = aj9smn - 0.5 * Q(nxpa, nypa - 1) * (P(2, nypa - 1) - P(nxpa - 1, nypa - 1));
  s.ajinat[k - 1] = dxa * dya * (fnot * adfaca * (aj5smn + 2.0 * aj9smn));
  double ah5sms = 0.5 * (D4(1, 2) - D4(1, 1));
  double ah5smn = 0.5 * (D4(1, nypa) - D4(1, nypa - 1));
  for (int i = 2; i <= nxpa - 1; ++i) {
    ah5sms = ah5sms + (D4(i, 2) - D4(i, 1));
    ah5smn = ah5smn + (D4(i, nypa) - D4(i, nypa - 1));
  }
  ah5sms = ah5sms + 0.5 * (D4(nxpa, 2) - D4(nxpa, 1));
  ah5smn = ah5smn + 0.5 * (D4(nxpa, nypa) - D4(nxpa, nypa - 1));
  s.ap5sat[k - 1] = ah4atk * ah5sms;
  s.ap5nat[k - 1] = ah4atk * ah5smn;
#undef D2
#undef D4
#undef P
#undef Q
#undef DQ
}

// ---------------------------------------------------------------- src/atisubs.F:60-293
void Model::atinvq() {
  const double ecrita = 1.0e-13;
  const size_t np = (size_t)nxpa * nypa;
  vec wrk(np * nla), bat(nxta);
  double ainhom[QGCM_NLMAX], ayis[QGCM_NLMAX], ayin[QGCM_NLMAX];
#define WRK(i, j, m) wrk[IX3(i, j, m, nxpa, nypa)]
#define QA(i, j, k) qa[IX3(i, j, k, nxpa, nypa)]
#define PA(i, j, k) pa[IX3(i, j, k, nxpa, nypa)]
#define PAM(i, j, k) pam[IX3(i, j, k, nxpa, nypa)]
#define CTL2M(k, m) MAT(c.ctl2mat, k, m, nla)
#define CTM2L(m, k) MAT(c.ctm2lat, m, k, nla)
  for (int j = 2; j <= nypa - 1; ++j) {
    const double betay = beta * yparel[j - 1];
    double ql[QGCM_NLMAX];
    for (int i = 1; i <= nxpa; ++i) {
      for (int k = 1; k <= nla; ++k) ql[k - 1] = QA(i, j, k) - betay;
      ql[0] = ql[0] - ddynat[IX2(i, j, nxpa)];
      for (int m = 1; m <= nla; ++m) {
        double qm = 0.0;
        for (int k = 1; k <= nla; ++k) qm = qm + CTL2M(k, m) * ql[k - 1];
        WRK(i, j, m) = fnot * qm;
      }
    }
  }
  for (int m = 1; m <= nla; ++m) {
    for (int i = 1; i <= nxta; ++i) bat[i - 1] = bd2at[i - 1] - c.rdm2at[m - 1];
    hscyat(&wrk[np * (m - 1)], bat.data());
    ainhom[m - 1] = xintp(&wrk[np * (m - 1)], nxpa, nypa);
    ainhom[m - 1] = ainhom[m - 1] * dxa * dya;
    s.xinhom_at[m - 1] = ainhom[m - 1];
    ayis[m - 1] = 0.5 * WRK(1, 2, m);
    ayin[m - 1] = -0.5 * WRK(1, nypa - 1, m);
    for (int i = 2; i <= nxpa - 1; ++i) {
      ayis[m - 1] = ayis[m - 1] + WRK(i, 2, m);
      ayin[m - 1] = ayin[m - 1] - WRK(i, nypa - 1, m);
    }
    ayis[m - 1] = ayis[m - 1] + 0.5 * WRK(nxpa, 2, m);
    ayin[m - 1] = ayin[m - 1] - 0.5 * WRK(nxpa, nypa - 1, m);
    ayis[m - 1] = ayis[m - 1] * (dxa / dya);
    ayin[m - 1] = ayin[m - 1] * (dxa / dya);
  }
  double rhss[QGCM_NLMAX], rhsn[QGCM_NLMAX], atsnew[QGCM_NLMAX], atnnew[QGCM_NLMAX];
  double clhss[QGCM_NLMAX], clhsn[QGCM_NLMAX], c1[QGCM_NLMAX], c2[QGCM_NLMAX], c3;
  double aipmod[QGCM_NLMAX], aiplay[QGCM_NLMAX];
  const double entfac = 0.5 * dya * fnot * fnot;
  const double *hat = c.hat;
  rhss[0] = -(entfac / hat[0]) * s.enisat[0] - (fnot / hat[0]) * s.txisat + s.ajisat[0] + s.ap5sat[0];
  rhsn[0] = -(entfac / hat[0]) * s.eninat[0] + (fnot / hat[0]) * s.txinat + s.ajinat[0] - s.ap5nat[0];
  for (int k = 2; k <= nla - 1; ++k) {
    rhss[k - 1] = -(entfac / hat[k - 1]) * (s.enisat[k - 1] - s.enisat[k - 2]) + s.ajisat[k - 1] + s.ap5sat[k - 1];
    rhsn[k - 1] = -(entfac / hat[k - 1]) * (s.eninat[k - 1] - s.eninat[k - 2]) + s.ajinat[k - 1] - s.ap5nat[k - 1];
  }
  rhss[nla - 1] = (entfac / hat[nla - 1]) * s.enisat[nla - 2] + s.ajisat[nla - 1] + s.ap5sat[nla - 1];
  rhsn[nla - 1] = (entfac / hat[nla - 1]) * s.eninat[nla - 2] + s.ajinat[nla - 1] - s.ap5nat[nla - 1];
  for (int k = 1; k <= nla; ++k) {
    atsnew[k - 1] = s.atmcsp[k - 1] + tdta * rhss[k - 1];
    atnnew[k - 1] = s.atmcnp[k - 1] + tdta * rhsn[k - 1];
    s.atmcsp[k - 1] = s.atmcs[k - 1];
    s.atmcnp[k - 1] = s.atmcn[k - 1];
    s.atmcs[k - 1] = atsnew[k - 1];
    s.atmcn[k - 1] = atnnew[k - 1];
  }
  for (int m = 1; m <= nla; ++m) {
    clhss[m - 1] = 0.0;
    clhsn[m - 1] = 0.0;
    for (int k = 1; k <= nla; ++k) {
      clhss[m - 1] = clhss[m - 1] + CTL2M(k, m) * atsnew[k - 1];
      clhsn[m - 1] = clhsn[m - 1] + CTL2M(k, m) * atnnew[k - 1];
    }
    clhss[m - 1] = clhss[m - 1] + ayis[m - 1];
    clhsn[m - 1] = clhsn[m - 1] - ayin[m - 1];
  }
  c3 = clhss[0] * s.hbsiat;
  for (int m = 1; m <= nla - 1; ++m) {
    c1[m - 1] = s.hc2nat[m - 1] * clhss[m] - s.hc2sat[m - 1] * clhsn[m];
    c2[m - 1] = s.hc1sat[m - 1] * clhsn[m] - s.hc1nat[m - 1] * clhss[m];
  }
  aipmod[0] = ainhom[0] + c3 * s.aipbha;
  for (int m = 2; m <= nla; ++m) aipmod[m - 1] = ainhom[m - 1] + (c1[m - 2] + c2[m - 2]) * s.aipcha[m - 2];
  for (int k = 1; k <= nla; ++k) {
    double pl = 0.0;
    for (int m = 1; m <= nla; ++m) pl = pl + CTM2L(m, k) * aipmod[m - 1];
    aiplay[k - 1] = pl;
  }
  for (int k = 1; k <= nla - 1; ++k) {
    const double est1 = aiplay[k - 1] - aiplay[k];
    const double est2 = s.dpiatp[k - 1] - tdta * c.gpat[k - 1] * s.xan[k - 1];
    const double edif = est1 - est2;
    const double esum = std::fabs(est1) + std::fabs(est2);
    s.ermasa[k - 1] = edif;
    if (esum > (ecrita * xla * yla * tdta * c.gpat[k - 1]))
      s.emfrat[k - 1] = 2.0 * edif / esum;
    else
      s.emfrat[k - 1] = 0.0;
    s.dpiatp[k - 1] = s.dpiat[k - 1];
    s.dpiat[k - 1] = aiplay[k - 1] - aiplay[k];
  }
  for (int j = 1; j <= nypa; ++j) {
    double homcor[QGCM_NLMAX], pm[QGCM_NLMAX];
    homcor[0] = c3 * pbhat[j - 1];
    for (int m = 2; m <= nla; ++m)
      homcor[m - 1] = c1[m - 2] * pch1at[IX2(j, m - 1, nypa)] + c2[m - 2] * pch2at[IX2(j, m - 1, nypa)];
    for (int i = 1; i <= nxpa; ++i) {
      for (int m = 1; m <= nla; ++m) pm[m - 1] = WRK(i, j, m) + homcor[m - 1];
      for (int k = 1; k <= nla; ++k) {
        PAM(i, j, k) = PA(i, j, k);
        double pl = 0.0;
        for (int m = 1; m <= nla; ++m) pl = pl + CTM2L(m, k) * pm[m - 1];
        PA(i, j, k) = pl;
      }
    }
  }
#undef WRK
#undef QA
#undef PA
#undef PAM
}

// ---------------------------------------------------------------- src/amlsubs.F:47-238
void Model::aml() {
  const size_t nt = (size_t)nxta * nyta;
  vec tmrhs(nt), hmrhs(nt), xfa(nt);
  const double hmat = c.hmat, hmamin = c.hmamin;
  const double hmainv = 1.0 / hmat;
  const double hdrcdt = c.hmadmp * rrcpat * tdta;
  const double tat1 = c.tat[0];
  const double diabcr = tat1 - 2.0 * hdrcdt;
  const double entfac = 1.0 / (tdta * (c.tat[1] - c.tat[0]));
  const double xcexp = c.xcexp;
  const double xbfac = xcexp * c.bface;
  double afacdp[QGCM_NLMAX];
  for (int l = 1; l <= nla - 1; ++l) afacdp[l - 1] = c.aface[l - 1] / c.gpat[l - 1];
  amladf(tmrhs.data(), hmrhs.data(), pa.data());
  double cfrasm = 0.0, centsm = 0.0;
#define T2(a, i, j) a[IX2(i, j, nxta)]
  for (int j = 1; j <= nyta; ++j) {
    for (int i = 1; i <= nxta; ++i) {
      double hnew, dtfix;
      if (T2(astm, i, j) <= diabcr) {
        const double dhdiab = hdrcdt * (T2(hmixam, i, j) - hmat) / (tat1 - T2(astm, i, j));
        hnew = T2(hmixam, i, j) + tdta * T2(hmrhs, i, j) - dhdiab;
        const double dhfix = std::max(hmamin - hnew, 0.0);
        hnew = hnew + dhfix;
        dtfix = dhfix * (tat1 - T2(astm, i, j)) / T2(hmixam, i, j);
      } else {
        hnew = hmat;
        dtfix = 0.0;
      }
      const double trhtot = T2(tmrhs, i, j) + rrcpat * T2(fnetat, i, j) / T2(hmixam, i, j) - hmainv * T2(wekta, i, j) * T2(astm, i, j);
      double astnew = T2(astm, i, j) + tdta * trhtot + dtfix;
      const double xfaent = xbfac * (T2(hmixam, i, j) - hmat) + c.dface * (xcexp * T2(astm, i, j) + T2(xc1ast, i, j));
      const double dtanew = tat1 - astnew;
      const double conena = entfac * T2(hmixa, i, j) * std::min(0.0, dtanew);
      T2(xfa, i, j) = xfaent - xcexp * conena;
      astnew = astnew + std::min(0.0, dtanew);
      cfrasm = cfrasm + (0.5 - sign(0.5, dtanew));
      centsm = centsm - conena;
      T2(astm, i, j) = T2(ast, i, j);
      T2(ast, i, j) = astnew;
      T2(hmixam, i, j) = T2(hmixa, i, j);
      T2(hmixa, i, j) = hnew;
    }
  }
#define EN(i, j) entat[IX2(i, j, nxpa)]
  for (int j = 2; j <= nypa - 1; ++j) {
    EN(1, j) = 0.25 * (T2(xfa, nxta, j - 1) + T2(xfa, 1, j - 1) + T2(xfa, nxta, j) + T2(xfa, 1, j));
    for (int i = 2; i <= nxpa - 1; ++i)
      EN(i, j) = 0.25 * (T2(xfa, i - 1, j - 1) + T2(xfa, i, j - 1) + T2(xfa, i - 1, j) + T2(xfa, i, j));
    EN(nxpa, j) = EN(1, j);
  }
  EN(1, 1) = 0.5 * (T2(xfa, nxta, 1) + T2(xfa, 1, 1));
  EN(1, nypa) = 0.5 * (T2(xfa, nxta, nyta) + T2(xfa, 1, nyta));
  for (int i = 2; i <= nxpa - 1; ++i) {
    EN(i, 1) = 0.5 * (T2(xfa, i - 1, 1) + T2(xfa, i, 1));
    EN(i, nypa) = 0.5 * (T2(xfa, i - 1, nyta) + T2(xfa, i, nyta));
  }
  EN(nxpa, 1) = EN(1, 1);
  EN(nxpa, nypa) = EN(1, nypa);
  for (int j = 1; j <= nypa; ++j)
    for (int i = 1; i <= nxpa; ++i) {
      double adpsum = 0.0;
      for (int l = 1; l <= nla - 1; ++l)
        adpsum = adpsum + afacdp[l - 1] * (pam[IX3(i, j, l, nxpa, nypa)] - pam[IX3(i, j, l + 1, nxpa, nypa)]);
      EN(i, j) = EN(i, j) + adpsum + c.cface * dtopat[IX2(i, j, nxpa)];
    }
  s.cfraat = cfrasm * atnorm;
  s.centat = centsm * dxa * dya;
  s.xan[0] = xintp(entat.data(), nxpa, nypa);
  s.xan[0] = s.xan[0] * dxa * dya;
  double ensums = 0.5 * EN(1, 1);
  double ensumn = 0.5 * EN(1, nypa);
  for (int i = 2; i <= nxpa - 1; ++i) {
    ensums = ensums + EN(i, 1);
    ensumn = ensumn + EN(i, nypa);
  }
  ensums = ensums + 0.5 * EN(nxpa, 1);
  ensumn = ensumn + 0.5 * EN(nxpa, nypa);
  s.enisat[0] = dxa * ensums;
  s.eninat[0] = dxa * ensumn;
#undef EN
#undef T2
}

// ---------------------------------------------------------------- src/amlsubs.F:246-563
// The reference spells out W/E columns, S/N rows and the four corners separately; the
// atmosphere is always x-periodic, so here one loop with wrapped neighbour indices
// (im, ip) evaluates the same expressions in the same association for every case.
void Model::amladf(double *tmrhs, double *hmrhs, const double *pa1) {
  const double d2tfac = c.at2d * dxam2;
  const double d4tfac = c.at4d * dxam2 * dxam2;
  const double hmdfac = c.ahmd * dxam2;
  const double hmat = c.hmat;
  const int nxd = nxta + 2;
  vec del2t((size_t)nxd * nyta);
#define D2T(i, j) del2t[(size_t)(i) + (size_t)nxd * ((j)-1)]
#define PA1(i, j) pa1[IX2(i, j, nxpa)]
#define AST(i, j) ast[IX2(i, j, nxta)]
#define ASTM(i, j) astm[IX2(i, j, nxta)]
#define HM(i, j) hmixa[IX2(i, j, nxta)]
#define HMM(i, j) hmixam[IX2(i, j, nxta)]
#define UE(i, j) uekat[IX2(i, j, nxpa)]
#define VE(i, j) vekat[IX2(i, j, nxta)]
  for (int j = 1; j <= nyta; ++j) {
    for (int i = 1; i <= nxta; ++i) {
      const int im = (i == 1) ? nxta : i - 1;
      const int ip = (i == nxta) ? 1 : i + 1;
      const double um = -rdxaf0 * (PA1(i, j + 1) - PA1(i, j)) + UE(i, j);
      const double up = -rdxaf0 * (PA1(i + 1, j + 1) - PA1(i + 1, j)) + UE(i + 1, j);
      const double tm = AST(i, j) + AST(im, j), tp = AST(i, j) + AST(ip, j);
      const double hm = HM(i, j) + HM(im, j), hp = HM(i, j) + HM(ip, j);
      const double xadvt = hdxam1 * (up * tp - um * tm);
      const double xadvh = hdxam1 * (up * hp - um * hm);
      double yadvt, yadvh, d2, d2h;
      if (j == 1) {
        const double vm = VE(i, 1);
        const double vp = rdxaf0 * (PA1(i + 1, 2) - PA1(i, 2)) + VE(i, 2);
        yadvt = hdxam1 * vp * (AST(i, 2) + AST(i, 1));
        yadvh = hdxam1 * (vp * (HM(i, 2) + HM(i, 1)) - vm * (HM(i, 1) + hmat));
        d2 = ASTM(im, 1) + ASTM(ip, 1) + ASTM(i, 2) - 3.0 * ASTM(i, 1);
        d2h = hmat + HMM(im, 1) + HMM(ip, 1) + HMM(i, 2) - 4.0 * HMM(i, 1);
      } else if (j == nyta) {
        const double vm = rdxaf0 * (PA1(i + 1, nyta) - PA1(i, nyta)) + VE(i, nyta);
        const double vp = VE(i, nyta + 1);
        yadvt = hdxam1 * (-vm * (AST(i, nyta) + AST(i, nyta - 1)));
        yadvh = hdxam1 * (vp * (hmat + HM(i, nyta)) - vm * (HM(i, nyta) + HM(i, nyta - 1)));
        d2 = ASTM(i, nyta - 1) + ASTM(im, nyta) + ASTM(ip, nyta) - 3.0 * ASTM(i, nyta);
        d2h = HMM(i, nyta - 1) + HMM(im, nyta) + HMM(ip, nyta) + hmat - 4.0 * HMM(i, nyta);
      } else {
        const double vm = rdxaf0 * (PA1(i + 1, j) - PA1(i, j)) + VE(i, j);
        const double vp = rdxaf0 * (PA1(i + 1, j + 1) - PA1(i, j + 1)) + VE(i, j + 1);
        yadvt = hdxam1 * (vp * (AST(i, j + 1) + AST(i, j)) - vm * (AST(i, j) + AST(i, j - 1)));
        yadvh = hdxam1 * (vp * (HM(i, j + 1) + HM(i, j)) - vm * (HM(i, j) + HM(i, j - 1)));
        d2 = ASTM(i, j - 1) + ASTM(im, j) + ASTM(ip, j) + ASTM(i, j + 1) - 4.0 * ASTM(i, j);
        d2h = HMM(i, j - 1) + HMM(im, j) + HMM(ip, j) + HMM(i, j + 1) - 4.0 * HMM(i, j);
      }
      tmrhs[IX2(i, j, nxta)] = -(xadvt + yadvt);
      D2T(i, j) = d2;
      hmrhs[IX2(i, j, nxta)] = -(xadvh + yadvh) + hmdfac * d2h;
    }
    D2T(0, j) = D2T(nxta, j);
    D2T(nxta + 1, j) = D2T(1, j);
  }
  for (int j = 2; j <= nyta - 1; ++j)
    for (int i = 1; i <= nxta; ++i)
      tmrhs[IX2(i, j, nxta)] = tmrhs[IX2(i, j, nxta)] + d2tfac * D2T(i, j) -
                               d4tfac * (D2T(i, j - 1) + D2T(i - 1, j) + D2T(i + 1, j) + D2T(i, j + 1) - 4.0 * D2T(i, j));
  for (int i = 1; i <= nxta; ++i) {
    tmrhs[IX2(i, 1, nxta)] = tmrhs[IX2(i, 1, nxta)] + d2tfac * D2T(i, 1) -
                             d4tfac * (D2T(i - 1, 1) + D2T(i + 1, 1) + D2T(i, 2) - 3.0 * D2T(i, 1));
    tmrhs[IX2(i, nyta, nxta)] = tmrhs[IX2(i, nyta, nxta)] + d2tfac * D2T(i, nyta) -
                                d4tfac * (D2T(i, nyta - 1) + D2T(i - 1, nyta) + D2T(i + 1, nyta) - 3.0 * D2T(i, nyta));
  }
#undef D2T
#undef PA1
#undef AST
#undef ASTM
#undef HM
#undef HMM
#undef UE
#undef VE
}

}  // namespace orc

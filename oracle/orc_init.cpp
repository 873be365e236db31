// TEST INFRASTRUCTURE ONLY -- CPU oracle (pinned against the translated reference, see orc_model.h).
// Restatement of constr (src/conhoms.F:44-314) and homsol (src/conhoms.F:318-818),
// and the main-loop body of src/q-gcm.F:1220-1408.
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "orc_model.h"

namespace orc {

#define IX2(i, j, nx) ((size_t)((i)-1) + (size_t)(nx) * (size_t)((j)-1))
#define IX3(i, j, k, nx, ny) ((size_t)((i)-1) + (size_t)(nx) * ((size_t)((j)-1) + (size_t)(ny) * (size_t)((k)-1)))
#define MAT(a, i, j, ld) (a)[((i)-1) + (size_t)(ld) * ((j)-1)]

// line integrals of p and dp/dy along the zonal boundaries + A-matrix combination
// (src/conhoms.F:123-190 ocean, :230-303 atmosphere; identical algebra)
static void constr_lines(const double *p, const double *pm, int nxp, int nyp, int nl, const double *amat,
                         double dx, double dy, double fnot, double *cs, double *cn, double *csp, double *cnp) {
  double pins[QGCM_NLMAX], pinn[QGCM_NLMAX], pinsp[QGCM_NLMAX], pinnp[QGCM_NLMAX];
#define P(i, j, k) p[IX3(i, j, k, nxp, nyp)]
#define PM(i, j, k) pm[IX3(i, j, k, nxp, nyp)]
  for (int k = 1; k <= nl; ++k) {
    pinsp[k - 1] = 0.5 * PM(1, 1, k);
    pinnp[k - 1] = 0.5 * PM(1, nyp, k);
    pins[k - 1] = 0.5 * P(1, 1, k);
    pinn[k - 1] = 0.5 * P(1, nyp, k);
    csp[k - 1] = 0.5 * (PM(1, 2, k) - PM(1, 1, k));
    cnp[k - 1] = 0.5 * (PM(1, nyp, k) - PM(1, nyp - 1, k));
    cs[k - 1] = 0.5 * (P(1, 2, k) - P(1, 1, k));
    cn[k - 1] = 0.5 * (P(1, nyp, k) - P(1, nyp - 1, k));
    for (int i = 2; i <= nxp - 1; ++i) {
      pinsp[k - 1] = pinsp[k - 1] + PM(i, 1, k);
      pinnp[k - 1] = pinnp[k - 1] + PM(i, nyp, k);
      pins[k - 1] = pins[k - 1] + P(i, 1, k);
      pinn[k - 1] = pinn[k - 1] + P(i, nyp, k);
      csp[k - 1] = csp[k - 1] + (PM(i, 2, k) - PM(i, 1, k));
      cnp[k - 1] = cnp[k - 1] + (PM(i, nyp, k) - PM(i, nyp - 1, k));
      cs[k - 1] = cs[k - 1] + (P(i, 2, k) - P(i, 1, k));
      cn[k - 1] = cn[k - 1] + (P(i, nyp, k) - P(i, nyp - 1, k));
    }
    pinsp[k - 1] = pinsp[k - 1] + 0.5 * PM(nxp, 1, k);
    pinnp[k - 1] = pinnp[k - 1] + 0.5 * PM(nxp, nyp, k);
    pins[k - 1] = pins[k - 1] + 0.5 * P(nxp, 1, k);
    pinn[k - 1] = pinn[k - 1] + 0.5 * P(nxp, nyp, k);
    csp[k - 1] = csp[k - 1] + 0.5 * (PM(nxp, 2, k) - PM(nxp, 1, k));
    cnp[k - 1] = cnp[k - 1] + 0.5 * (PM(nxp, nyp, k) - PM(nxp, nyp - 1, k));
    cs[k - 1] = cs[k - 1] + 0.5 * (P(nxp, 2, k) - P(nxp, 1, k));
    cn[k - 1] = cn[k - 1] + 0.5 * (P(nxp, nyp, k) - P(nxp, nyp - 1, k));
    csp[k - 1] = csp[k - 1] * (dx / dy);
    cnp[k - 1] = cnp[k - 1] * (dx / dy);
    cs[k - 1] = cs[k - 1] * (dx / dy);
    cn[k - 1] = cn[k - 1] * (dx / dy);
    pinsp[k - 1] = dx * pinsp[k - 1];
    pinnp[k - 1] = dx * pinnp[k - 1];
    pins[k - 1] = dx * pins[k - 1];
    pinn[k - 1] = dx * pinn[k - 1];
  }
  for (int k = 1; k <= nl; ++k) {
    double apsp = 0.0, apnp = 0.0, aps = 0.0, apn = 0.0;
    for (int j = 1; j <= nl; ++j) {
      apsp = apsp + MAT(amat, k, j, nl) * pinsp[j - 1];
      apnp = apnp + MAT(amat, k, j, nl) * pinnp[j - 1];
      aps = aps + MAT(amat, k, j, nl) * pins[j - 1];
      apn = apn + MAT(amat, k, j, nl) * pinn[j - 1];
    }
    csp[k - 1] = -csp[k - 1] + 0.5 * dy * fnot * fnot * apsp;
    cnp[k - 1] = cnp[k - 1] + 0.5 * dy * fnot * fnot * apnp;
    cs[k - 1] = -cs[k - 1] + 0.5 * dy * fnot * fnot * aps;
    cn[k - 1] = cn[k - 1] + 0.5 * dy * fnot * fnot * apn;
  }
#undef P
#undef PM
}

// ---------------------------------------------------------------- src/conhoms.F:44-314
void Model::constr() {
  if (!atmos_only) {
    const size_t np = (size_t)nxpo * nypo;
    vec w1(np), w2(np);
    for (int k = 1; k <= nlo - 1; ++k) {
      for (size_t i = 0; i < np; ++i) {
        w1[i] = pom[np * k + i] - pom[np * (k - 1) + i];
        w2[i] = po[np * k + i] - po[np * (k - 1) + i];
      }
      s.dpiocp[k - 1] = xintp(w1.data(), nxpo, nypo) * dxo * dyo;
      s.dpioc[k - 1] = xintp(w2.data(), nxpo, nypo) * dxo * dyo;
    }
    if (cyclic)
      constr_lines(po.data(), pom.data(), nxpo, nypo, nlo, c.amatoc, dxo, dyo, fnot, s.ocncs, s.ocncn, s.ocncsp, s.ocncnp);
  }
  if (!ocean_only) {
    const size_t np = (size_t)nxpa * nypa;
    vec w1(np), w2(np);
    for (int k = 1; k <= nla - 1; ++k) {
      for (size_t i = 0; i < np; ++i) {
        w1[i] = pam[np * (k - 1) + i] - pam[np * k + i];
        w2[i] = pa[np * (k - 1) + i] - pa[np * k + i];
      }
      s.dpiatp[k - 1] = xintp(w1.data(), nxpa, nypa) * dxa * dya;
      s.dpiat[k - 1] = xintp(w2.data(), nxpa, nypa) * dxa * dya;
    }
    constr_lines(pa.data(), pam.data(), nxpa, nypa, nla, c.amatat, dxa, dya, fnot, s.atmcs, s.atmcn, s.atmcsp, s.atmcnp);
  }
}

// cyclic-channel homogeneous solutions, shared by ocean (src/conhoms.F:386-543) and
// atmosphere (src/conhoms.F:655-810)
template <class Solve>
static void homsol_channel(int nxp, int nyp, int nxt, int nl, const double *yp, double xl, double yl, double dx,
                           double dy, const double *bd2, const double *rdm2, Solve solve, double *pch1,
                           double *pch2, double *pbh, double *hc1s, double *hc2s, double *hc1n, double *hc2n,
                           double *aipch, double *hbsi, double *aipbh) {
  for (int j = 1; j <= nyp; ++j) pbh[j - 1] = (double)(nyp - j) / (double)(nyp - 1);
  *hbsi = yl / xl;
  *aipbh = 0.5 * xl * yl;
  const size_t np = (size_t)nxp * nyp;
  vec wk1(np), wk2(np), b(nxt);
  for (int m = 1; m <= nl - 1; ++m) {
    for (int i = 1; i <= nxt; ++i) b[i - 1] = bd2[i - 1] - rdm2[m];
#define PCH1(j, m) pch1[IX2(j, m, nyp)]
#define PCH2(j, m) pch2[IX2(j, m, nyp)]
    for (int j = 1; j <= nyp; ++j) {
      PCH1(j, m) = (yp[nyp - 1] - yp[j - 1]) / yl;
      PCH2(j, m) = (yp[j - 1] - yp[0]) / yl;
      for (int i = 1; i <= nxp; ++i) {
        wk1[IX2(i, j, nxp)] = PCH1(j, m);
        wk2[IX2(i, j, nxp)] = PCH2(j, m);
      }
    }
    solve(wk1.data(), b.data());
    solve(wk2.data(), b.data());
    for (int j = 1; j <= nyp; ++j) {
      for (int i = 1; i <= nxp; ++i) {
        wk1[IX2(i, j, nxp)] = PCH1(j, m) + rdm2[m] * wk1[IX2(i, j, nxp)];
        wk2[IX2(i, j, nxp)] = PCH2(j, m) + rdm2[m] * wk2[IX2(i, j, nxp)];
      }
      PCH1(j, m) = wk1[IX2(1, j, nxp)];
      PCH2(j, m) = wk2[IX2(1, j, nxp)];
    }
    const double aipch1 = Model::xintp(wk1.data(), nxp, nyp);
    const double aipch2 = Model::xintp(wk2.data(), nxp, nyp);
    aipch[m - 1] = 0.5 * (aipch1 + aipch2) * dx * dy;
    double pch1ys = (PCH1(2, m) - PCH1(1, m)) / dy;
    double pch2ys = (PCH2(2, m) - PCH2(1, m)) / dy;
    double pch1yn = (PCH1(nyp, m) - PCH1(nyp - 1, m)) / dy;
    double pch2yn = (PCH2(nyp, m) - PCH2(nyp - 1, m)) / dy;
    pch1ys = -pch1ys + 0.5 * dy * rdm2[m] * PCH1(1, m);
    pch2ys = -pch2ys + 0.5 * dy * rdm2[m] * PCH2(1, m);
    pch1yn = pch1yn + 0.5 * dy * rdm2[m] * PCH1(nyp, m);
    pch2yn = pch2yn + 0.5 * dy * rdm2[m] * PCH2(nyp, m);
    pch1ys = xl * pch1ys;
    pch2ys = xl * pch2ys;
    pch1yn = xl * pch1yn;
    pch2yn = xl * pch2yn;
    const double pchdet = pch1ys * pch2yn - pch2ys * pch1yn;
    hc1s[m - 1] = pch1ys / pchdet;
    hc2s[m - 1] = pch2ys / pchdet;
    hc1n[m - 1] = pch1yn / pchdet;
    hc2n[m - 1] = pch2yn / pchdet;
#undef PCH1
#undef PCH2
  }
}

// ---------------------------------------------------------------- src/conhoms.F:318-818
void Model::homsol() {
  if (!atmos_only) {
    if (cyclic) {
      homsol_channel(nxpo, nypo, nxto, nlo, ypo.data(), xlo, ylo, dxo, dyo, bd2oc.data(), c.rdm2oc,
                     [this](double *w, const double *b) { hscyoc(w, b); }, pch1oc.data(), pch2oc.data(),
                     pbhoc.data(), s.hc1soc, s.hc2soc, s.hc1noc, s.hc2noc, s.aipcho, &s.hbsioc, &s.aipbho);
    } else {
      const size_t np = (size_t)nxpo * nypo;
      vec boc(nxto);
      for (int m = 1; m <= nlo - 1; ++m) {
        for (int i = 1; i <= nxto; ++i) boc[i - 1] = bd2oc[i - 1] - c.rdm2oc[m];
        double *oh = &ochom[np * (m - 1)];
        for (size_t i = 0; i < np; ++i) oh[i] = 1.0;
        hsbxoc(oh, boc.data());
        for (size_t i = 0; i < np; ++i) oh[i] = 1.0 + c.rdm2oc[m] * oh[i];
        s.aipohs[m - 1] = xintp(oh, nxpo, nypo) * dxo * dyo;
      }
      for (int k = 1; k <= nlo - 1; ++k) {
        for (int m = 1; m <= nlo; ++m)
          MAT(s.cdiffo, m, k, nlo) = MAT(c.ctm2loc, m, k + 1, nlo) - MAT(c.ctm2loc, m, k, nlo);
        for (int m = 1; m <= nlo - 1; ++m)
          MAT(s.cdhoc, k, m, nlo - 1) = (MAT(c.ctm2loc, m + 1, k + 1, nlo) - MAT(c.ctm2loc, m + 1, k, nlo)) * s.aipohs[m - 1];
      }
    }
  }
  if (!ocean_only) {
    homsol_channel(nxpa, nypa, nxta, nla, ypa.data(), xla, yla, dxa, dya, bd2at.data(), c.rdm2at,
                   [this](double *w, const double *b) { hscyat(w, b); }, pch1at.data(), pch2at.data(), pbhat.data(),
                   s.hc1sat, s.hc2sat, s.hc1nat, s.hc2nat, s.aipcha, &s.hbsiat, &s.aipbha);
  }
}

// ---------------------------------------------------------------- src/q-gcm.F:1220-1408
// nstr == 1: mod(nt,1).eq.1 is never true in the reference, so the
// shipped NAtl 1 km deck never steps the ocean; here the ocean steps every nt instead.
void Model::run(int64_t nt_first, int64_t nt_last) {
  for (int64_t nt = nt_first; nt <= nt_last; ++nt) {
    const bool ocstep = (nstr == 1) ? true : (nt % nstr == 1);
    if (ocstep) {
      if (!ocean_only) xforc();
      if (!atmos_only) {
        oml();
        qgostep();
        ocinvq();
        ocqbdy(qo.data(), po.data());
      }
    }
    if (!ocean_only) {
      aml();
      qgastep();
      atinvq();
      atqzbd(qa.data(), pa.data());
    }
    if (!atmos_only && ((nt - 1) % (25 * (int64_t)nstr) == 0)) tlavg_ocean();
    if (!ocean_only && ((nt - 1) % 100 == 0)) tlavg_atmos();
  }
}

}  // namespace orc

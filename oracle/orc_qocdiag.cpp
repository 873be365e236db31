// TEST INFRASTRUCTURE ONLY -- CPU oracle (pinned against the translated reference, see orc_model.h).
// Restatement of qocdiag_out (src/qocdiag.F:303-683) without its netCDF calls: per layer the
// vorticity tendency and its Jacobian, del-4th, del-6th and forcing/drag terms, sub-sampled
// by nsko into the vectors the reference hands to nf_put_vara_double.
// out[((t*nlo + k)*jpwk + j)*ipwk + i], t = 0 dqdt, 1 qotjac, 2 qt2dif, 3 qt4dif, 4 qotent.
#include <algorithm>
#include <cmath>

#include "orc_model.h"

namespace orc {

#define IX2(i, j, nx) ((size_t)((i)-1) + (size_t)(nx) * (size_t)((j)-1))
#define IX3(i, j, k, nx, ny) ((size_t)((i)-1) + (size_t)(nx) * ((size_t)((j)-1) + (size_t)(ny) * (size_t)((k)-1)))

void Model::qocdiag(int nsko, double *out) {
  // src/qocdiag.F:352-372
  int mwk = nxpo % nsko;
  const int ipwk = std::min(mwk, 1) + (nxpo - mwk) / nsko;
  mwk = nypo % nsko;
  const int jpwk = std::min(mwk, 1) + (nypo - mwk) / nsko;
  const double adfaco = 1.0 / (12.0 * dxo * dyo * fnot);
  const double bcfaco = c.bccooc * dxom2 / (0.5 * c.bccooc + 1.0);
  double fohfac[QGCM_NLMAX];
  for (int k = 1; k <= nlo; ++k) fohfac[k - 1] = fnot / c.hoc[k - 1];
  const double bdrfac = 0.5 * (fnot < 0.0 ? -1.0 : 1.0) * c.delek / c.hoc[nlo - 1];
  const double rdto = 1.0 / dto;
  const size_t np = (size_t)nxpo * nypo;
  vec del2p(np), del4p(np), qt2dif(np), qt4dif(np), qotjac(np), qotent(np), dqdt(np);
#define P(a, i, j) a[IX3(i, j, k, nxpo, nypo)]
#define D2(i, j) del2p[IX2(i, j, nxpo)]
#define D4(i, j) del4p[IX2(i, j, nxpo)]
  for (int k = 1; k <= nlo; ++k) {
    const double ah2fac = c.ah2oc[k - 1] / fnot, ah4fac = c.ah4oc[k - 1] / fnot;
    // del-sqd(pom), :399-436
    for (int i = 1; i <= nxpo; ++i) {
      D2(i, 1) = bcfaco * (P(pom, i, 2) - P(pom, i, 1));
      D2(i, nypo) = bcfaco * (P(pom, i, nypo - 1) - P(pom, i, nypo));
    }
#pragma omp parallel for schedule(static)
    for (int j = 2; j <= nypo - 1; ++j) {
      if (cyclic)
        D2(1, j) = (P(pom, 1, j - 1) + P(pom, nxpo - 1, j) + P(pom, 2, j) + P(pom, 1, j + 1) - 4.0 * P(pom, 1, j)) * dxom2;
      else
        D2(1, j) = bcfaco * (P(pom, 2, j) - P(pom, 1, j));
      for (int i = 2; i <= nxpo - 1; ++i)
        D2(i, j) = (P(pom, i, j - 1) + P(pom, i - 1, j) + P(pom, i + 1, j) + P(pom, i, j + 1) - 4.0 * P(pom, i, j)) * dxom2;
      if (cyclic)
        D2(nxpo, j) = D2(1, j);
      else
        D2(nxpo, j) = bcfaco * (P(pom, nxpo - 1, j) - P(pom, nxpo, j));
    }
    // del-4th(pom), :442-477
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
        D4(i, j) = (D2(i, j - 1) + D2(i - 1, j) + D2(i + 1, j) + D2(i, j + 1) - 4.0 * D2(i, j)) * dxom2;
      if (cyclic)
        D4(nxpo, j) = D4(1, j);
      else
        D4(nxpo, j) = bcfaco * (D2(nxpo - 1, j) - D2(nxpo, j));
    }
    // :483-493
    for (size_t n = 0; n < np; ++n) qt2dif[n] = qt4dif[n] = qotjac[n] = qotent[n] = dqdt[n] = 0.0;
    // :499-603
#pragma omp parallel for schedule(static)
    for (int j = 2; j <= nypo - 1; ++j) {
      auto forcing = [&](int i) {
        double e;
        if (k == 1) e = fohfac[0] * (wekpo[IX2(i, j, nxpo)] - entoc[IX2(i, j, nxpo)]);
        else if (k == 2) e = fohfac[1] * entoc[IX2(i, j, nxpo)];
        else e = 0.0;
        if (k == nlo) e = e - bdrfac * D2(i, j);
        return e;
      };
      if (cyclic) {
        const int w = nxpo - 1;
        const double d6p = dxom2 * (D4(1, j - 1) + D4(w, j) + D4(2, j) + D4(1, j + 1) - 4.0 * D4(1, j));
        qt2dif[IX2(1, j, nxpo)] = ah2fac * D4(1, j);
        qt4dif[IX2(1, j, nxpo)] = -ah4fac * d6p;
        qotjac[IX2(1, j, nxpo)] =
            adfaco * ((P(qo, 2, j) - P(qo, w, j)) * (P(po, 1, j + 1) - P(po, 1, j - 1)) +
                      (P(qo, 1, j - 1) - P(qo, 1, j + 1)) * (P(po, 2, j) - P(po, w, j)) +
                      P(qo, 2, j) * (P(po, 2, j + 1) - P(po, 2, j - 1)) - P(qo, w, j) * (P(po, w, j + 1) - P(po, w, j - 1)) -
                      P(qo, 1, j + 1) * (P(po, 2, j + 1) - P(po, w, j + 1)) + P(qo, 1, j - 1) * (P(po, 2, j - 1) - P(po, w, j - 1)) +
                      P(po, 1, j + 1) * (P(qo, 2, j + 1) - P(qo, w, j + 1)) - P(po, 1, j - 1) * (P(qo, 2, j - 1) - P(qo, w, j - 1)) -
                      P(po, 2, j) * (P(qo, 2, j + 1) - P(qo, 2, j - 1)) + P(po, w, j) * (P(qo, w, j + 1) - P(qo, w, j - 1)));
        qotent[IX2(1, j, nxpo)] = forcing(1);
        dqdt[IX2(1, j, nxpo)] = qotjac[IX2(1, j, nxpo)] + qt2dif[IX2(1, j, nxpo)] + qt4dif[IX2(1, j, nxpo)] + qotent[IX2(1, j, nxpo)];
      } else {
        dqdt[IX2(1, j, nxpo)] = rdto * (P(qo, 1, j) - P(qom, 1, j));
      }
      for (int i = 2; i <= nxpo - 1; ++i) {
        const double d6p = dxom2 * (D4(i, j - 1) + D4(i - 1, j) + D4(i + 1, j) + D4(i, j + 1) - 4.0 * D4(i, j));
        qt2dif[IX2(i, j, nxpo)] = ah2fac * D4(i, j);
        qt4dif[IX2(i, j, nxpo)] = -ah4fac * d6p;
        qotjac[IX2(i, j, nxpo)] =
            adfaco * ((P(qo, i + 1, j) - P(qo, i - 1, j)) * (P(po, i, j + 1) - P(po, i, j - 1)) +
                      (P(qo, i, j - 1) - P(qo, i, j + 1)) * (P(po, i + 1, j) - P(po, i - 1, j)) +
                      P(qo, i + 1, j) * (P(po, i + 1, j + 1) - P(po, i + 1, j - 1)) -
                      P(qo, i - 1, j) * (P(po, i - 1, j + 1) - P(po, i - 1, j - 1)) -
                      P(qo, i, j + 1) * (P(po, i + 1, j + 1) - P(po, i - 1, j + 1)) +
                      P(qo, i, j - 1) * (P(po, i + 1, j - 1) - P(po, i - 1, j - 1)) +
                      P(po, i, j + 1) * (P(qo, i + 1, j + 1) - P(qo, i - 1, j + 1)) -
                      P(po, i, j - 1) * (P(qo, i + 1, j - 1) - P(qo, i - 1, j - 1)) -
                      P(po, i + 1, j) * (P(qo, i + 1, j + 1) - P(qo, i + 1, j - 1)) +
                      P(po, i - 1, j) * (P(qo, i - 1, j + 1) - P(qo, i - 1, j - 1)));
        qotent[IX2(i, j, nxpo)] = forcing(i);
        dqdt[IX2(i, j, nxpo)] = qotjac[IX2(i, j, nxpo)] + qt2dif[IX2(i, j, nxpo)] + qt4dif[IX2(i, j, nxpo)] + qotent[IX2(i, j, nxpo)];
      }
      if (cyclic) {
        qt2dif[IX2(nxpo, j, nxpo)] = qt2dif[IX2(1, j, nxpo)];
        qt4dif[IX2(nxpo, j, nxpo)] = qt4dif[IX2(1, j, nxpo)];
        qotjac[IX2(nxpo, j, nxpo)] = qotjac[IX2(1, j, nxpo)];
        qotent[IX2(nxpo, j, nxpo)] = qotent[IX2(1, j, nxpo)];
        dqdt[IX2(nxpo, j, nxpo)] = dqdt[IX2(1, j, nxpo)];
      } else {
        dqdt[IX2(nxpo, j, nxpo)] = rdto * (P(qo, nxpo, j) - P(qom, nxpo, j));
      }
    }
    // zonal boundaries: time difference of qo, :607-612
    for (int i = 1; i <= nxpo; ++i) {
      dqdt[IX2(i, 1, nxpo)] = rdto * (P(qo, i, 1) - P(qom, i, 1));
      dqdt[IX2(i, nypo, nxpo)] = rdto * (P(qo, i, nypo) - P(qom, i, nypo));
    }
    // sub-sampling, :616-672
    const vec *terms[5] = {&dqdt, &qotjac, &qt2dif, &qt4dif, &qotent};
    for (int t = 0; t < 5; ++t)
      for (int j = 1; j <= jpwk; ++j)
        for (int i = 1; i <= ipwk; ++i)
          out[(((size_t)t * nlo + (k - 1)) * jpwk + (j - 1)) * ipwk + (i - 1)] = (*terms[t])[IX2(1 + (i - 1) * nsko, 1 + (j - 1) * nsko, nxpo)];
  }
#undef P
#undef D2
#undef D4
}

}  // namespace orc

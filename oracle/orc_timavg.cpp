// TEST INFRASTRUCTURE ONLY -- CPU oracle (pinned against the translated reference, see orc_model.h).
// Restatement of the running-sum accumulators of src/timavge.F: tavini (:108-273), tavatm
// (:278-419), tavocn (:425-617) and the fork's avg_ocn_k247 (:624-660).  tavout's netCDF
// dump stays Fortran; the sums and the contribution counts are what it reads.
#include <algorithm>

#include "orc_model.h"

namespace orc {

#define IX2(i, j, nx) ((size_t)((i)-1) + (size_t)(nx) * (size_t)((j)-1))
#define IX3(i, j, k, nx, ny) ((size_t)((i)-1) + (size_t)(nx) * ((size_t)((j)-1) + (size_t)(ny) * (size_t)((k)-1)))

// src/timavge.F:108-273
void Model::tavini(int which) {
  // which: 1 = atmosphere sums, 2 = ocean sums, 3 = both (the reference's tavini)
  if (!ocean_only && (which & 1)) {
    for (vec *v : {&txatav, &tyatav}) v->assign((size_t)nxpa * nypa, 0.0);
    for (vec *v : {&wtatav, &fmatav, &astav}) v->assign((size_t)nxta * nyta, 0.0);
    for (vec *v : {&uufa, &tufa, &utufa}) v->assign((size_t)nxpa * nyta, 0.0);
    for (vec *v : {&vvfa, &tvfa, &vtvfa}) v->assign((size_t)nxta * nypa, 0.0);
    for (vec *v : {&patav, &qatav}) v->assign((size_t)nxpa * nypa * nla, 0.0);
    nsumat = 0;
  }
  if (!atmos_only && (which & 2)) {
    for (vec *v : {&txocav, &tyocav, &wpocav}) v->assign((size_t)nxpo * nypo, 0.0);
    for (vec *v : {&wtocav, &fmocav, &sstav}) v->assign((size_t)nxto * nyto, 0.0);
    for (vec *v : {&uufo, &tufo, &utufo}) v->assign((size_t)nxpo * nyto, 0.0);
    for (vec *v : {&vvfo, &tvfo, &vtvfo}) v->assign((size_t)nxto * nypo, 0.0);
    for (vec *v : {&pocav, &qocav, &po_avg}) v->assign((size_t)nxpo * nypo * nlo, 0.0);
    nsumoc = nsum_ocavg = 0;
  }
}

// src/timavge.F:278-419
void Model::tavatm() {
  if (ocean_only) return;
  if (txatav.empty()) tavini(1);
  const double rhf0hm = 0.5 / (fnot * c.hmat);
#pragma omp parallel
  {
#pragma omp for schedule(static) nowait
    for (int j = 1; j <= nypa; ++j)
      for (int i = 1; i <= nxpa; ++i) {
        txatav[IX2(i, j, nxpa)] = txatav[IX2(i, j, nxpa)] + tauxa[IX2(i, j, nxpa)];
        tyatav[IX2(i, j, nxpa)] = tyatav[IX2(i, j, nxpa)] + tauya[IX2(i, j, nxpa)];
      }
#pragma omp for schedule(static) nowait
    for (int j = 1; j <= nyta; ++j)
      for (int i = 1; i <= nxta; ++i) {
        wtatav[IX2(i, j, nxta)] = wtatav[IX2(i, j, nxta)] + wekta[IX2(i, j, nxta)];
        fmatav[IX2(i, j, nxta)] = fmatav[IX2(i, j, nxta)] + fnetat[IX2(i, j, nxta)];
        astav[IX2(i, j, nxta)] = astav[IX2(i, j, nxta)] + ast[IX2(i, j, nxta)];
      }
    // zonal advection (:343-360)
#pragma omp for schedule(static) nowait
    for (int j = 1; j <= nyta; ++j) {
      vec tuf(nxpa + 1);
      tuf[1] = 0.5 * (ast[IX2(1, j, nxta)] + ast[IX2(nxta, j, nxta)]);
      for (int i = 2; i <= nxpa - 1; ++i) tuf[i] = 0.5 * (ast[IX2(i, j, nxta)] + ast[IX2(i - 1, j, nxta)]);
      tuf[nxpa] = 0.5 * (ast[IX2(1, j, nxta)] + ast[IX2(nxta, j, nxta)]);
      for (int i = 1; i <= nxpa; ++i) {
        const double uuf = -rdxaf0 * (pa[IX3(i, j + 1, 1, nxpa, nypa)] - pa[IX3(i, j, 1, nxpa, nypa)]) -
                           rhf0hm * (tauya[IX2(i, j + 1, nxpa)] + tauya[IX2(i, j, nxpa)]);
        uufa[IX2(i, j, nxpa)] = uufa[IX2(i, j, nxpa)] + uuf;
        tufa[IX2(i, j, nxpa)] = tufa[IX2(i, j, nxpa)] + tuf[i];
        utufa[IX2(i, j, nxpa)] = utufa[IX2(i, j, nxpa)] + uuf * tuf[i];
      }
    }
    // meridional advection, inner rows (:365-377)
#pragma omp for schedule(static) nowait
    for (int j = 2; j <= nyta; ++j)
      for (int i = 1; i <= nxta; ++i) {
        const double vvf = rdxaf0 * (pa[IX3(i + 1, j, 1, nxpa, nypa)] - pa[IX3(i, j, 1, nxpa, nypa)]) +
                           rhf0hm * (tauxa[IX2(i + 1, j, nxpa)] + tauxa[IX2(i, j, nxpa)]);
        const double tvf = 0.5 * (ast[IX2(i, j, nxta)] + ast[IX2(i, j - 1, nxta)]);
        const double vtvf = vvf * tvf;
        vvfa[IX2(i, j, nxta)] = vvfa[IX2(i, j, nxta)] + vvf;
        tvfa[IX2(i, j, nxta)] = tvfa[IX2(i, j, nxta)] + tvf;
        vtvfa[IX2(i, j, nxta)] = vtvfa[IX2(i, j, nxta)] + vtvf;
      }
    // zonal boundaries (:385-399): v = 0, boundary temperature = interior point
#pragma omp for schedule(static) nowait
    for (int i = 1; i <= nxta; ++i) {
      vvfa[IX2(i, 1, nxta)] = vvfa[IX2(i, 1, nxta)] + 0.0;
      tvfa[IX2(i, 1, nxta)] = tvfa[IX2(i, 1, nxta)] + ast[IX2(i, 1, nxta)];
      vtvfa[IX2(i, 1, nxta)] = vtvfa[IX2(i, 1, nxta)] + 0.0;
      vvfa[IX2(i, nypa, nxta)] = vvfa[IX2(i, nypa, nxta)] + 0.0;
      tvfa[IX2(i, nypa, nxta)] = tvfa[IX2(i, nypa, nxta)] + ast[IX2(i, nypa - 1, nxta)];
      vtvfa[IX2(i, nypa, nxta)] = vtvfa[IX2(i, nypa, nxta)] + 0.0;
    }
    for (int k = 1; k <= nla; ++k) {
#pragma omp for schedule(static) nowait
      for (int j = 1; j <= nypa; ++j)
        for (int i = 1; i <= nxpa; ++i) {
          patav[IX3(i, j, k, nxpa, nypa)] = patav[IX3(i, j, k, nxpa, nypa)] + pa[IX3(i, j, k, nxpa, nypa)];
          qatav[IX3(i, j, k, nxpa, nypa)] = qatav[IX3(i, j, k, nxpa, nypa)] + qa[IX3(i, j, k, nxpa, nypa)];
        }
    }
  }
  nsumat = nsumat + 1;
}

// src/timavge.F:425-617
void Model::tavocn() {
  if (atmos_only) return;
  if (txocav.empty()) tavini(2);
  const double uvgfac = c.ycexp * rdxof0;
  const double rhf0hm = 0.5 / (fnot * c.hmoc);
  const double tsbdy = c.tsbdy, tnbdy = c.tnbdy;
#pragma omp parallel
  {
#pragma omp for schedule(static) nowait
    for (int j = 1; j <= nypo; ++j)
      for (int i = 1; i <= nxpo; ++i) {
        txocav[IX2(i, j, nxpo)] = txocav[IX2(i, j, nxpo)] + tauxo[IX2(i, j, nxpo)];
        tyocav[IX2(i, j, nxpo)] = tyocav[IX2(i, j, nxpo)] + tauyo[IX2(i, j, nxpo)];
        wpocav[IX2(i, j, nxpo)] = wpocav[IX2(i, j, nxpo)] + wekpo[IX2(i, j, nxpo)];
      }
#pragma omp for schedule(static) nowait
    for (int j = 1; j <= nyto; ++j)
      for (int i = 1; i <= nxto; ++i) {
        wtocav[IX2(i, j, nxto)] = wtocav[IX2(i, j, nxto)] + wekto[IX2(i, j, nxto)];
        fmocav[IX2(i, j, nxto)] = fmocav[IX2(i, j, nxto)] + fnetoc[IX2(i, j, nxto)];
        sstav[IX2(i, j, nxto)] = sstav[IX2(i, j, nxto)] + sst[IX2(i, j, nxto)];
      }
    // zonal advection (:487-531)
#pragma omp for schedule(static) nowait
    for (int j = 1; j <= nyto; ++j) {
      vec uuf(nxpo + 1), tuf(nxpo + 1), utuf(nxpo + 1);
      if (cyclic) {
        uuf[1] = -uvgfac * (po[IX3(1, j + 1, 1, nxpo, nypo)] - po[IX3(1, j, 1, nxpo, nypo)]) +
                 rhf0hm * (tauyo[IX2(1, j + 1, nxpo)] + tauyo[IX2(1, j, nxpo)]);
        tuf[1] = 0.5 * (sst[IX2(1, j, nxto)] + sst[IX2(nxto, j, nxto)]);
        utuf[1] = uuf[1] * tuf[1];
      } else {
        uuf[1] = 0.0;
        tuf[1] = sst[IX2(1, j, nxto)];
        utuf[1] = 0.0;
      }
      for (int i = 2; i <= nxpo - 1; ++i) {
        uuf[i] = -uvgfac * (po[IX3(i, j + 1, 1, nxpo, nypo)] - po[IX3(i, j, 1, nxpo, nypo)]) +
                 rhf0hm * (tauyo[IX2(i, j + 1, nxpo)] + tauyo[IX2(i, j, nxpo)]);
        tuf[i] = 0.5 * (sst[IX2(i, j, nxto)] + sst[IX2(i - 1, j, nxto)]);
        utuf[i] = uuf[i] * tuf[i];
      }
      if (cyclic) {
        uuf[nxpo] = uuf[1];
        tuf[nxpo] = tuf[1];
        utuf[nxpo] = utuf[1];
      } else {
        uuf[nxpo] = 0.0;
        tuf[nxpo] = sst[IX2(nxto, j, nxto)];
        utuf[nxpo] = 0.0;
      }
      for (int i = 1; i <= nxpo; ++i) {
        uufo[IX2(i, j, nxpo)] = uufo[IX2(i, j, nxpo)] + uuf[i];
        tufo[IX2(i, j, nxpo)] = tufo[IX2(i, j, nxpo)] + tuf[i];
        utufo[IX2(i, j, nxpo)] = utufo[IX2(i, j, nxpo)] + utuf[i];
      }
    }
    // meridional advection, inner rows (:537-548)
#pragma omp for schedule(static) nowait
    for (int j = 2; j <= nyto; ++j)
      for (int i = 1; i <= nxto; ++i) {
        const double vvf = uvgfac * (po[IX3(i + 1, j, 1, nxpo, nypo)] - po[IX3(i, j, 1, nxpo, nypo)]) -
                           rhf0hm * (tauxo[IX2(i + 1, j, nxpo)] + tauxo[IX2(i, j, nxpo)]);
        const double tvf = 0.5 * (sst[IX2(i, j, nxto)] + sst[IX2(i, j - 1, nxto)]);
        const double vtvf = vvf * tvf;
        vvfo[IX2(i, j, nxto)] = vvfo[IX2(i, j, nxto)] + vvf;
        tvfo[IX2(i, j, nxto)] = tvfo[IX2(i, j, nxto)] + tvf;
        vtvfo[IX2(i, j, nxto)] = vtvfo[IX2(i, j, nxto)] + vtvf;
      }
    // zonal boundaries (:553-592)
#pragma omp for schedule(static) nowait
    for (int i = 1; i <= nxto; ++i) {
      double vvf, tvf, vtvf;
      if (sb_hflux) {
        vvf = -rhf0hm * (tauxo[IX2(i + 1, 1, nxpo)] + tauxo[IX2(i, 1, nxpo)]);
        tvf = 0.5 * (sst[IX2(i, 1, nxto)] + tsbdy);
        vtvf = vvf * tvf;
      } else {
        vvf = 0.0;
        tvf = sst[IX2(i, 1, nxto)];
        vtvf = 0.0;
      }
      vvfo[IX2(i, 1, nxto)] = vvfo[IX2(i, 1, nxto)] + vvf;
      tvfo[IX2(i, 1, nxto)] = tvfo[IX2(i, 1, nxto)] + tvf;
      vtvfo[IX2(i, 1, nxto)] = vtvfo[IX2(i, 1, nxto)] + vtvf;
      if (nb_hflux) {
        vvf = -rhf0hm * (tauxo[IX2(i + 1, nyto + 1, nxpo)] + tauxo[IX2(i, nyto + 1, nxpo)]);
        tvf = 0.5 * (sst[IX2(i, nyto, nxto)] + tnbdy);
        vtvf = vvf * tvf;
      } else {
        vvf = 0.0;
        tvf = sst[IX2(i, nyto, nxto)];
        vtvf = 0.0;
      }
      vvfo[IX2(i, nypo, nxto)] = vvfo[IX2(i, nypo, nxto)] + vvf;
      tvfo[IX2(i, nypo, nxto)] = tvfo[IX2(i, nypo, nxto)] + tvf;
      vtvfo[IX2(i, nypo, nxto)] = vtvfo[IX2(i, nypo, nxto)] + vtvf;
    }
    for (int k = 1; k <= nlo; ++k) {
#pragma omp for schedule(static) nowait
      for (int j = 1; j <= nypo; ++j)
        for (int i = 1; i <= nxpo; ++i) {
          pocav[IX3(i, j, k, nxpo, nypo)] = pocav[IX3(i, j, k, nxpo, nypo)] + po[IX3(i, j, k, nxpo, nypo)];
          qocav[IX3(i, j, k, nxpo, nypo)] = qocav[IX3(i, j, k, nxpo, nypo)] + qo[IX3(i, j, k, nxpo, nypo)];
        }
    }
  }
  nsumoc = nsumoc + 1;
}

// src/timavge.F:624-660
void Model::avg_ocn_k247() {
  if (atmos_only) return;
  if (po_avg.empty()) tavini(2);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < po_avg.size(); ++i) po_avg[i] = po_avg[i] + po[i];
  nsum_ocavg = nsum_ocavg + 1;
}

}  // namespace orc

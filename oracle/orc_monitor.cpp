// TEST INFRASTRUCTURE ONLY -- CPU oracle (pinned against the translated reference, see orc_model.h).
// Restatement of the ocean section of monnc_comp (src/monitor_diag.F:480-840) with its helpers
// del4bx (:900-1015), del4ch (:1020-1155) and genint (:1160-1210).  Same loops, same
// expression association (including the reference's own ugdot expression, :676-677, whose
// pom(i,j,k) terms cancel), without the OpenMP directives' reduction order.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "orc_model.h"

namespace orc {

#define IX2(i, j, nx) ((size_t)((i)-1) + (size_t)(nx) * (size_t)((j)-1))
#define IX3(i, j, k, nx, ny) ((size_t)((i)-1) + (size_t)(nx) * ((size_t)((j)-1) + (size_t)(ny) * (size_t)((k)-1)))

// src/monitor_diag.F:1160-1210
static double genint(const double *val, int nx, int ny, double facwe, double facsn) {
  double answer = 0.0;
  for (int j = 2; j <= ny - 1; ++j) {
    double sumi = facwe * val[IX2(1, j, nx)];
    for (int i = 2; i <= nx - 1; ++i) sumi = sumi + val[IX2(i, j, nx)];
    sumi = sumi + facwe * val[IX2(nx, j, nx)];
    answer = answer + sumi;
  }
  double xxs = facwe * val[IX2(1, 1, nx)], xxn = facwe * val[IX2(1, ny, nx)];
  for (int i = 2; i <= nx - 1; ++i) {
    xxs = xxs + val[IX2(i, 1, nx)];
    xxn = xxn + val[IX2(i, ny, nx)];
  }
  xxs = xxs + facwe * val[IX2(nx, 1, nx)];
  xxn = xxn + facwe * val[IX2(nx, ny, nx)];
  return answer + facsn * (xxs + xxn);
}

// one application of the Laplacian of del4bx (:925-968) / del4ch (:1048-1093)
static void lap_onesided(const double *arr, int nx, int ny, double dxm2, bool cyclic, double *del2) {
#define A(i, j) arr[IX2(i, j, nx)]
#define D(i, j) del2[IX2(i, j, nx)]
  for (int j = 2; j <= ny - 1; ++j) {
    if (cyclic)
      D(1, j) = dxm2 * (A(1, j - 1) + A(nx, j) + A(2, j) + A(1, j + 1) - 4.0 * A(1, j));
    else
      D(1, j) = dxm2 * (A(3, j) - 2.0 * A(2, j) + A(1, j) + A(1, j - 1) - 2.0 * A(1, j) + A(1, j + 1));
    for (int i = 2; i <= nx - 1; ++i) D(i, j) = dxm2 * (A(i, j - 1) + A(i - 1, j) + A(i + 1, j) + A(i, j + 1) - 4.0 * A(i, j));
    if (cyclic)
      D(nx, j) = dxm2 * (A(nx, j - 1) + A(nx - 1, j) + A(1, j) + A(nx, j + 1) - 4.0 * A(nx, j));
    else
      D(nx, j) = dxm2 * (A(nx, j) - 2.0 * A(nx - 1, j) + A(nx - 2, j) + A(nx, j - 1) - 2.0 * A(nx, j) + A(nx, j + 1));
  }
  for (int i = 2; i <= nx - 1; ++i) {
    D(i, 1) = dxm2 * (A(i - 1, 1) - 2.0 * A(i, 1) + A(i + 1, 1) + A(i, 3) - 2.0 * A(i, 2) + A(i, 1));
    D(i, ny) = dxm2 * (A(i - 1, ny) - 2.0 * A(i, ny) + A(i + 1, ny) + A(i, ny) - 2.0 * A(i, ny - 1) + A(i, ny - 2));
  }
  if (cyclic) {
    D(1, 1) = dxm2 * (A(nx, 1) - 2.0 * A(1, 1) + A(2, 1) + A(1, 3) - 2.0 * A(1, 2) + A(1, 1));
    D(1, ny) = dxm2 * (A(nx, ny) - 2.0 * A(1, ny) + A(2, ny) + A(1, ny) - 2.0 * A(1, ny - 1) + A(1, ny - 2));
    D(nx, 1) = dxm2 * (A(nx - 1, 1) - 2.0 * A(nx, 1) + A(1, 1) + A(nx, 3) - 2.0 * A(nx, 2) + A(nx, 1));
    D(nx, ny) = dxm2 * (A(nx - 1, ny) - 2.0 * A(nx, ny) + A(1, ny) + A(nx, ny) - 2.0 * A(nx, ny - 1) + A(nx, ny - 2));
  } else {
    D(1, 1) = dxm2 * (A(3, 1) - 2.0 * A(2, 1) + A(1, 1) + A(1, 3) - 2.0 * A(1, 2) + A(1, 1));
    D(nx, 1) = dxm2 * (A(nx, 1) - 2.0 * A(nx - 1, 1) + A(nx - 2, 1) + A(nx, 3) - 2.0 * A(nx, 2) + A(nx, 1));
    D(1, ny) = dxm2 * (A(3, ny) - 2.0 * A(2, ny) + A(1, ny) + A(1, ny) - 2.0 * A(1, ny - 1) + A(1, ny - 2));
    D(nx, ny) = dxm2 * (A(nx, ny) - 2.0 * A(nx - 1, ny) + A(nx - 2, ny) + A(nx, ny) - 2.0 * A(nx, ny - 1) + A(nx, ny - 2));
  }
#undef A
#undef D
}

void Model::monnc_ocean(qgcm_monitor_ocean *r) {
  std::memset(r, 0, sizeof(*r));
  if (atmos_only) return;
  const double rhooc = c.rhooc, cpoc = c.cpoc, delek = c.delek;
  const size_t np = (size_t)nxpo * nypo;
  vec octwk1((size_t)nxto * nypo), octwk2((size_t)nxto * nypo), octwk3((size_t)nxto * nypo), octwk4((size_t)nxto * nypo);
  vec ocpwk1(np), ocpwk2(np), ocpwk3(np), ocpwk4(np), etaoc(np), ugoc((size_t)nxpo * nyto), vgoc((size_t)nxto * nypo);
  // reference pressures on the equatorward side (:174-183)
  double poref[QGCM_NLMAX] = {0};
  for (int k = 1; k <= nlo; ++k) poref[k - 1] = (fnot > 0.0) ? po[IX3(1, 1, k, nxpo, nypo)] : po[IX3(1, nypo, k, nxpo, nypo)];
  // Ekman velocity (:498-527)
  r->wetmoc = genint(wekto.data(), nxto, nyto, 1.0, 1.0);
  for (int j = 1; j <= nyto; ++j)
    for (int i = 1; i <= nxto; ++i) octwk1[IX2(i, j, nxto)] = std::fabs(wekto[IX2(i, j, nxto)]);
  r->watmoc = genint(octwk1.data(), nxto, nyto, 1.0, 1.0);
  r->wepmoc = genint(wekpo.data(), nxpo, nypo, 0.5, 0.5);
  for (size_t n = 0; n < np; ++n) ocpwk1[n] = std::fabs(wekpo[n]);
  r->wapmoc = genint(ocpwk1.data(), nxpo, nypo, 0.5, 0.5);
  r->wetmoc = r->wetmoc * ocnorm;
  r->watmoc = r->watmoc * ocnorm;
  r->wepmoc = r->wepmoc * ocnorm;
  r->wapmoc = r->wapmoc * ocnorm;
  // entrainment (:529-543)
  r->entmoc = genint(entoc.data(), nxpo, nypo, 0.5, 0.5);
  for (size_t n = 0; n < np; ++n) ocpwk1[n] = std::fabs(entoc[n]);
  r->enamoc = genint(ocpwk1.data(), nxpo, nypo, 0.5, 0.5);
  r->entmoc = r->entmoc * ocnorm;
  r->enamoc = r->enamoc * ocnorm;
  // interface displacements (:545-583)
  for (int k = 1; k <= nlo - 1; ++k) {
    const double rgpoc = 1.0 / c.gpoc[k - 1];
    for (int j = 1; j <= nypo; ++j)
      for (int i = 1; i <= nxpo; ++i) {
        const double eta = rgpoc * (po[IX3(i, j, k + 1, nxpo, nypo)] - po[IX3(i, j, k, nxpo, nypo)]);
        const double etadot = (rgpoc / dto) * (po[IX3(i, j, k, nxpo, nypo)] - po[IX3(i, j, k + 1, nxpo, nypo)] -
                                               pom[IX3(i, j, k, nxpo, nypo)] + pom[IX3(i, j, k + 1, nxpo, nypo)]);
        etaoc[IX2(i, j, nxpo)] = eta;
        ocpwk1[IX2(i, j, nxpo)] = eta * eta;
        ocpwk2[IX2(i, j, nxpo)] = eta * etadot;
        ocpwk3[IX2(i, j, nxpo)] = eta * entoc[IX2(i, j, nxpo)];
      }
    const double etaint = genint(etaoc.data(), nxpo, nypo, 0.5, 0.5);
    double et2now = genint(ocpwk1.data(), nxpo, nypo, 0.5, 0.5);
    const double et2dot = genint(ocpwk2.data(), nxpo, nypo, 0.5, 0.5);
    r->etamoc[k - 1] = etaint * ocnorm;
    et2now = et2now * ocnorm;
    r->ddtpeoc[k - 1] = rhooc * c.gpoc[k - 1] * et2dot;
    r->et2moc[k - 1] = et2now;
    if (k == 1) {
      const double pkeint = genint(ocpwk3.data(), nxpo, nypo, 0.5, 0.5);
      r->pkenoc = rhooc * c.gpoc[0] * pkeint * ocnorm;
    }
  }
  // wind work (:588-617)
  for (int j = 1; j <= nyto; ++j)
    for (int i = 1; i <= nxpo; ++i) {
      const double ugeos = -rdxof0 * (po[IX3(i, j + 1, 1, nxpo, nypo)] - po[IX3(i, j, 1, nxpo, nypo)]);
      const double tauxav = 0.5 * (tauxo[IX2(i, j + 1, nxpo)] + tauxo[IX2(i, j, nxpo)]);
      ocpwk1[IX2(i, j, nxpo)] = ugeos * tauxav;
    }
  const double utaux = genint(ocpwk1.data(), nxpo, nyto, 0.5, 1.0);
  for (int j = 1; j <= nypo; ++j)
    for (int i = 1; i <= nxto; ++i) {
      const double vgeos = rdxof0 * (po[IX3(i + 1, j, 1, nxpo, nypo)] - po[IX3(i, j, 1, nxpo, nypo)]);
      const double tauyav = 0.5 * (tauyo[IX2(i + 1, j, nxpo)] + tauyo[IX2(i, j, nxpo)]);
      octwk1[IX2(i, j, nxto)] = vgeos * tauyav;
    }
  const double vtauy = genint(octwk1.data(), nxto, nypo, 1.0, 0.5);
  r->utauoc = rhooc * (vtauy + utaux) * ocnorm;
  // layer kinetic energy, dissipation, stream function (:620-752)
  for (int k = 1; k <= nlo; ++k) {
    for (int j = 1; j <= nyto; ++j)
      for (int i = 1; i <= nxpo; ++i) ugoc[IX2(i, j, nxpo)] = -rdxof0 * (pom[IX3(i, j + 1, k, nxpo, nypo)] - pom[IX3(i, j, k, nxpo, nypo)]);
    lap_onesided(ugoc.data(), nxpo, nyto, dxom2, cyclic, ocpwk1.data());
    lap_onesided(ocpwk1.data(), nxpo, nyto, dxom2, cyclic, ocpwk2.data());
    for (int j = 1; j <= nypo; ++j)
      for (int i = 1; i <= nxto; ++i) vgoc[IX2(i, j, nxto)] = rdxof0 * (pom[IX3(i + 1, j, k, nxpo, nypo)] - pom[IX3(i, j, k, nxpo, nypo)]);
    lap_onesided(vgoc.data(), nxto, nypo, dxom2, cyclic, octwk1.data());
    lap_onesided(octwk1.data(), nxto, nypo, dxom2, cyclic, octwk2.data());
    double pomin = 1.0e30, pomax = -1.0e30;
    for (int j = 1; j <= nypo; ++j)
      for (int i = 1; i <= nxpo; ++i) {
        pomin = std::min(pomin, po[IX3(i, j, k, nxpo, nypo)]);
        pomax = std::max(pomax, po[IX3(i, j, k, nxpo, nypo)]);
      }
    vec ujeto(nyto + 1);
    for (int j = 1; j <= nyto; ++j) {
      double ujet = 0.0, ugeos = 0.0;
      for (int i = 1; i <= nxpo; ++i) {
        ugeos = -rdxof0 * (po[IX3(i, j + 1, k, nxpo, nypo)] - po[IX3(i, j, k, nxpo, nypo)]);
        const double ugdot = -(rdxof0 / dto) * (po[IX3(i, j + 1, k, nxpo, nypo)] - pom[IX3(i, j, k, nxpo, nypo)] -
                                                pom[IX3(i, j + 1, k, nxpo, nypo)] + pom[IX3(i, j, k, nxpo, nypo)]);
        ujet = ujet + ugeos;
        ocpwk1[IX2(i, j, nxpo)] = ugeos * ocpwk1[IX2(i, j, nxpo)];
        ocpwk2[IX2(i, j, nxpo)] = ugeos * ocpwk2[IX2(i, j, nxpo)];
        ocpwk3[IX2(i, j, nxpo)] = ugeos * ugeos;
        ocpwk4[IX2(i, j, nxpo)] = ugeos * ugdot;
      }
      ujet = ujet - ugeos;
      ujeto[j] = std::fabs(ujet) / (double)nxto;
    }
    r->ocjpos[k - 1] = 0;
    r->ocjval[k - 1] = 0.0;
    for (int j = 1; j <= nyto; ++j)
      if (ujeto[j] > r->ocjval[k - 1]) {
        r->ocjpos[k - 1] = j;
        r->ocjval[k - 1] = ujeto[j];
      }
    const double u2diss = genint(ocpwk1.data(), nxpo, nyto, 0.5, 1.0);
    const double u4diss = genint(ocpwk2.data(), nxpo, nyto, 0.5, 1.0);
    const double uke = genint(ocpwk3.data(), nxpo, nyto, 0.5, 1.0);
    const double ukedot = genint(ocpwk4.data(), nxpo, nyto, 0.5, 1.0);
    for (int j = 1; j <= nypo; ++j)
      for (int i = 1; i <= nxto; ++i) {
        const double vgeos = rdxof0 * (po[IX3(i + 1, j, k, nxpo, nypo)] - po[IX3(i, j, k, nxpo, nypo)]);
        const double vgdot = (rdxof0 / dto) * (po[IX3(i + 1, j, k, nxpo, nypo)] - po[IX3(i, j, k, nxpo, nypo)] -
                                               pom[IX3(i + 1, j, k, nxpo, nypo)] + pom[IX3(i, j, k, nxpo, nypo)]);
        octwk1[IX2(i, j, nxto)] = vgeos * octwk1[IX2(i, j, nxto)];
        octwk2[IX2(i, j, nxto)] = vgeos * octwk2[IX2(i, j, nxto)];
        octwk3[IX2(i, j, nxto)] = vgeos * vgeos;
        octwk4[IX2(i, j, nxto)] = vgeos * vgdot;
      }
    const double pint = genint(&po[IX3(1, 1, k, nxpo, nypo)], nxpo, nypo, 0.5, 0.5);
    const double qint = genint(&qo[IX3(1, 1, k, nxpo, nypo)], nxpo, nypo, 0.5, 0.5);
    const double v2diss = genint(octwk1.data(), nxto, nypo, 1.0, 0.5);
    const double v4diss = genint(octwk2.data(), nxto, nypo, 1.0, 0.5);
    const double vke = genint(octwk3.data(), nxto, nypo, 1.0, 0.5);
    const double vkedot = genint(octwk4.data(), nxto, nypo, 1.0, 0.5);
    r->pavgoc[k - 1] = pint * ocnorm;
    r->qavgoc[k - 1] = qint * ocnorm;
    r->ah2doc[k - 1] = -rhooc * c.ah2oc[k - 1] * c.hoc[k - 1] * (u2diss + v2diss) * ocnorm;
    r->ah4doc[k - 1] = rhooc * c.ah4oc[k - 1] * c.hoc[k - 1] * (u4diss + v4diss) * ocnorm;
    r->kealoc[k - 1] = 0.5 * rhooc * c.hoc[k - 1] * (uke + vke) * ocnorm;
    r->ddtkeoc[k - 1] = rhooc * c.hoc[k - 1] * (ukedot + vkedot) * ocnorm;
    double psiext = std::min(pomin / fnot, pomax / fnot);
    r->osfmin[k - 1] = 1.0e-6 * c.hoc[k - 1] * (psiext - poref[k - 1] / fnot);
    psiext = std::max(pomin / fnot, pomax / fnot);
    r->osfmax[k - 1] = 1.0e-6 * c.hoc[k - 1] * (psiext - poref[k - 1] / fnot);
    r->occirc[k - 1] = 1.0e-6 * c.hoc[k - 1] * (po[IX3(1, 1, k, nxpo, nypo)] - po[IX3(1, nypo, k, nxpo, nypo)]) / fnot;
  }
  // bottom drag (:755-783)
  for (int j = 1; j <= nyto; ++j)
    for (int i = 1; i <= nxpo; ++i) {
      const double ugeos = -rdxof0 * (pom[IX3(i, j + 1, nlo, nxpo, nypo)] - pom[IX3(i, j, nlo, nxpo, nypo)]);
      ocpwk1[IX2(i, j, nxpo)] = ugeos * ugeos;
    }
  const double u2d = genint(ocpwk1.data(), nxpo, nyto, 0.5, 1.0);
  for (int j = 1; j <= nypo; ++j)
    for (int i = 1; i <= nxto; ++i) {
      const double vgeos = rdxof0 * (pom[IX3(i + 1, j, nlo, nxpo, nypo)] - pom[IX3(i, j, nlo, nxpo, nypo)]);
      octwk1[IX2(i, j, nxto)] = vgeos * vgeos;
    }
  const double v2d = genint(octwk1.data(), nxto, nypo, 1.0, 0.5);
  r->btdgoc = 0.5 * rhooc * delek * std::fabs(fnot) * (u2d + v2d) * ocnorm;
  // mixed layer (:788-812)
  r->sstmin = 1.0e30;
  r->sstmax = -1.0e30;
  for (int j = 1; j <= nyto; ++j)
    for (int i = 1; i <= nxto; ++i) {
      r->sstmin = std::min(r->sstmin, sst[IX2(i, j, nxto)]);
      r->sstmax = std::max(r->sstmax, sst[IX2(i, j, nxto)]);
      octwk1[IX2(i, j, nxto)] = sst[IX2(i, j, nxto)] * wekto[IX2(i, j, nxto)];
    }
  r->hfmloc = genint(octwk1.data(), nxto, nyto, 1.0, 1.0);
  r->tmlmoc = genint(sst.data(), nxto, nyto, 1.0, 1.0);
  r->tmlmoc = r->tmlmoc * ocnorm;
  r->hfmloc = rhooc * cpoc * r->hfmloc * ocnorm;
  // total circulation (:818-821)
  r->occtot = 0.0;
  for (int k = 1; k <= nlo; ++k) r->occtot = r->occtot + r->occirc[k - 1];
  couroc(r);      // :826
}

// src/monitor_diag.F:1450-1925.  The reference walks every T cell with a recurrence on the
// western/eastern face velocities (um, up) and the southern/northern ones (vm, vp), with the
// boundary faces set by the configuration (no normal flow; Ekman outflow for the mixed layer
// under sb_hflux / nb_hflux; periodic in a channel); restated here as one loop over the cells
// with the face values spelled out.  Extrema do not depend on the order of the cells.  (In the
// mixed-layer rows j = 1 and j = nyto the reference leaves the western face out of umin/umax,
// :1566-1593, :1667-1690; the eastern face of the same row carries the same value in a channel
// and both are zero in a box, so the extrema are the same.)
void Model::couroc(qgcm_monitor_ocean *r) {
  const double uvgfac = c.ycexp * rdxof0, rhf0hm = 0.5 / (fnot * c.hmoc);
  for (int k = 0; k <= nlo; ++k) {       // k = 0: mixed layer (po(:,:,1) + Ekman), k >= 1: QG layer k
    const bool ml = (k == 0);
    const int kk = ml ? 1 : k;
    const double ug = ml ? uvgfac : rdxof0, rh = ml ? rhf0hm : 0.0;
    auto uface = [&](int f, int j) {     // face f = 1..nxto+1 of T row j
      if (!cyclic && (f == 1 || f == nxto + 1)) return 0.0;
      double u = -ug * (po[IX3(f, j + 1, kk, nxpo, nypo)] - po[IX3(f, j, kk, nxpo, nypo)]);
      if (ml) u = u + rh * (tauyo[IX2(f, j + 1, nxpo)] + tauyo[IX2(f, j, nxpo)]);
      return u;
    };
    auto vface = [&](int i, int jf) {    // face jf = 1..nyto+1 of T column i
      if (jf == 1) return (ml && sb_hflux) ? -rh * (tauxo[IX2(i + 1, 1, nxpo)] + tauxo[IX2(i, 1, nxpo)]) : 0.0;
      if (jf == nyto + 1) return (ml && nb_hflux) ? -rh * (tauxo[IX2(i + 1, nyto + 1, nxpo)] + tauxo[IX2(i, nyto + 1, nxpo)]) : 0.0;
      double v = ug * (po[IX3(i + 1, jf, kk, nxpo, nypo)] - po[IX3(i, jf, kk, nxpo, nypo)]);
      if (ml) v = v - rh * (tauxo[IX2(i + 1, jf, nxpo)] + tauxo[IX2(i, jf, nxpo)]);
      return v;
    };
    double umin = 1.0e30, umax = -1.0e30, vmin = 1.0e30, vmax = -1.0e30, vsqmax = -1.0e30;
    for (int j = 1; j <= nyto; ++j) {
      double up = uface(1, j);
      umin = std::min(umin, up);
      umax = std::max(umax, up);
      for (int i = 1; i <= nxto; ++i) {
        const double um = up;
        up = uface(i + 1, j);
        const double vm = vface(i, j), vp = vface(i, j + 1);
        umin = std::min(umin, up);
        umax = std::max(umax, up);
        vmin = std::min(vmin, std::min(vm, vp));
        vmax = std::max(vmax, std::max(vm, vp));
        const double velsqd = (um + up) * (um + up) + (vm + vp) * (vm + vp);
        vsqmax = std::max(vsqmax, velsqd);
      }
    }
    if (ml) {
      r->umminoc = umin; r->ummaxoc = umax; r->vmminoc = vmin; r->vmmaxoc = vmax;
      r->cnmloc = hdxom1 * dto * std::sqrt(vsqmax);
    } else {
      r->ugminoc[k - 1] = umin; r->ugmaxoc[k - 1] = umax; r->vgminoc[k - 1] = vmin; r->vgmaxoc[k - 1] = vmax;
      r->cnqgoc[k - 1] = hdxom1 * dto * std::sqrt(vsqmax);
    }
  }
}

// atmosphere section of monnc_comp, src/monitor_diag.F:186-478
void Model::monnc_atmos(qgcm_monitor_atmos *r) {
  std::memset(r, 0, sizeof(*r));
  if (ocean_only) return;
  const double rhoat = c.rhoat, cpat = c.cpat;
  const size_t np = (size_t)nxpa * nypa;
  vec attwk1((size_t)nxta * nypa), attwk2((size_t)nxta * nypa), attwk3((size_t)nxta * nypa);
  vec atpwk1(np), atpwk2(np), atpwk3(np), etaat(np), ugat((size_t)nxpa * nyta), vgat((size_t)nxta * nypa);
  // Ekman velocity (:204-226)
  r->wetmat = genint(wekta.data(), nxta, nyta, 1.0, 1.0);
  for (int j = 1; j <= nyta; ++j)
    for (int i = 1; i <= nxta; ++i) attwk1[IX2(i, j, nxta)] = std::fabs(wekta[IX2(i, j, nxta)]);
  r->watmat = genint(attwk1.data(), nxta, nyta, 1.0, 1.0);
  r->wepmat = genint(wekpa.data(), nxpa, nypa, 0.5, 0.5);
  for (size_t n = 0; n < np; ++n) atpwk1[n] = std::fabs(wekpa[n]);
  r->wapmat = genint(atpwk1.data(), nxpa, nypa, 0.5, 0.5);
  r->wetmat = r->wetmat * atnorm;
  r->watmat = r->watmat * atnorm;
  r->wepmat = r->wepmat * atnorm;
  r->wapmat = r->wapmat * atnorm;
  // entrainment across interface 1 (:233-249)
  r->entmat[0] = genint(entat.data(), nxpa, nypa, 0.5, 0.5);
  for (size_t n = 0; n < np; ++n) atpwk1[n] = std::fabs(entat[n]);
  r->enamat[0] = genint(atpwk1.data(), nxpa, nypa, 0.5, 0.5);
  r->entmat[0] = r->entmat[0] * atnorm;
  r->enamat[0] = r->enamat[0] * atnorm;
  // interface displacements (:254-290)
  for (int k = 1; k <= nla - 1; ++k) {
    const double rgpat = 1.0 / c.gpat[k - 1];
    for (int j = 1; j <= nypa; ++j)
      for (int i = 1; i <= nxpa; ++i) {
        const double eta = rgpat * (pa[IX3(i, j, k, nxpa, nypa)] - pa[IX3(i, j, k + 1, nxpa, nypa)]);
        const double etadot = (rgpat / dta) * (pa[IX3(i, j, k, nxpa, nypa)] - pa[IX3(i, j, k + 1, nxpa, nypa)] -
                                               pam[IX3(i, j, k, nxpa, nypa)] + pam[IX3(i, j, k + 1, nxpa, nypa)]);
        etaat[IX2(i, j, nxpa)] = eta;
        atpwk1[IX2(i, j, nxpa)] = eta * eta;
        atpwk2[IX2(i, j, nxpa)] = eta * etadot;
        atpwk3[IX2(i, j, nxpa)] = eta * entat[IX2(i, j, nxpa)];
      }
    const double etaint = genint(etaat.data(), nxpa, nypa, 0.5, 0.5);
    double et2now = genint(atpwk1.data(), nxpa, nypa, 0.5, 0.5);
    const double et2dot = genint(atpwk2.data(), nxpa, nypa, 0.5, 0.5);
    r->etamat[k - 1] = etaint * atnorm;
    et2now = et2now * atnorm;
    r->pkenat[k - 1] = 0.0;
    r->ddtpeat[k - 1] = rhoat * c.gpat[k - 1] * et2dot;
    r->et2mat[k - 1] = et2now;
    if (k == 1) {
      const double pkeint = genint(atpwk3.data(), nxpa, nypa, 0.5, 0.5);
      r->pkenat[0] = rhoat * c.gpat[0] * pkeint * atnorm;
    }
  }
  // wind work (:298-326)
  for (int j = 1; j <= nyta; ++j)
    for (int i = 1; i <= nxpa; ++i) {
      const double ugeos = -rdxaf0 * (pa[IX3(i, j + 1, 1, nxpa, nypa)] - pa[IX3(i, j, 1, nxpa, nypa)]);
      const double tauxav = 0.5 * (tauxa[IX2(i, j + 1, nxpa)] + tauxa[IX2(i, j, nxpa)]);
      atpwk1[IX2(i, j, nxpa)] = ugeos * tauxav;
    }
  const double utaux = genint(atpwk1.data(), nxpa, nyta, 0.5, 1.0);
  for (int j = 1; j <= nypa; ++j)
    for (int i = 1; i <= nxta; ++i) {
      const double vgeos = rdxaf0 * (pa[IX3(i + 1, j, 1, nxpa, nypa)] - pa[IX3(i, j, 1, nxpa, nypa)]);
      const double tauyav = 0.5 * (tauya[IX2(i + 1, j, nxpa)] + tauya[IX2(i, j, nxpa)]);
      attwk1[IX2(i, j, nxta)] = vgeos * tauyav;
    }
  const double vtauy = genint(attwk1.data(), nxta, nypa, 1.0, 0.5);
  r->utauat = rhoat * (vtauy + utaux) * atnorm;
  // layers (:330-418)
  for (int k = 1; k <= nla; ++k) {
    for (int j = 1; j <= nyta; ++j)
      for (int i = 1; i <= nxpa; ++i) ugat[IX2(i, j, nxpa)] = -rdxaf0 * (pam[IX3(i, j + 1, k, nxpa, nypa)] - pam[IX3(i, j, k, nxpa, nypa)]);
    lap_onesided(ugat.data(), nxpa, nyta, dxam2, true, atpwk3.data());      // del4ch: del-sqd in atpwk3,
    lap_onesided(atpwk3.data(), nxpa, nyta, dxam2, true, atpwk1.data());    // del-4th in atpwk1
    for (int j = 1; j <= nypa; ++j)
      for (int i = 1; i <= nxta; ++i) vgat[IX2(i, j, nxta)] = rdxaf0 * (pam[IX3(i + 1, j, k, nxpa, nypa)] - pam[IX3(i, j, k, nxpa, nypa)]);
    lap_onesided(vgat.data(), nxta, nypa, dxam2, true, attwk3.data());      // del-sqd in attwk3,
    lap_onesided(attwk3.data(), nxta, nypa, dxam2, true, attwk2.data());    // del-4th in attwk2
    vec ujeta(nyta + 1);
    for (int j = 1; j <= nyta; ++j) {
      double ujet = 0.0, ugeos = 0.0;
      for (int i = 1; i <= nxpa; ++i) {
        ugeos = -rdxaf0 * (pa[IX3(i, j + 1, k, nxpa, nypa)] - pa[IX3(i, j, k, nxpa, nypa)]);
        const double ugdot = -(rdxaf0 / dta) * (pa[IX3(i, j + 1, k, nxpa, nypa)] - pa[IX3(i, j, k, nxpa, nypa)] -
                                                pam[IX3(i, j + 1, k, nxpa, nypa)] + pam[IX3(i, j, k, nxpa, nypa)]);
        ujet = ujet + ugeos;
        atpwk1[IX2(i, j, nxpa)] = ugeos * atpwk1[IX2(i, j, nxpa)];
        atpwk2[IX2(i, j, nxpa)] = ugeos * ugeos;
        atpwk3[IX2(i, j, nxpa)] = ugeos * ugdot;
      }
      ujet = ujet - ugeos;
      ujeta[j] = std::fabs(ujet) / (double)nxta;
    }
    r->atstpos[k - 1] = 0;
    r->atstval[k - 1] = 0.0;
    for (int j = 1; j <= nyta; ++j)
      if (ujeta[j] > r->atstval[k - 1]) {
        r->atstpos[k - 1] = j;
        r->atstval[k - 1] = ujeta[j];
      }
    const double u4diss = genint(atpwk1.data(), nxpa, nyta, 0.5, 1.0);
    const double uke = genint(atpwk2.data(), nxpa, nyta, 0.5, 1.0);
    const double ukedot = genint(atpwk3.data(), nxpa, nyta, 0.5, 1.0);
    for (int j = 1; j <= nypa; ++j)
      for (int i = 1; i <= nxta; ++i) {
        const double vgeos = rdxaf0 * (pa[IX3(i + 1, j, k, nxpa, nypa)] - pa[IX3(i, j, k, nxpa, nypa)]);
        // vgdot is computed at :387-388 but never stored: attwk3 keeps del-sqd of the lagged v
        attwk1[IX2(i, j, nxta)] = vgeos * attwk2[IX2(i, j, nxta)];
        attwk2[IX2(i, j, nxta)] = vgeos * vgeos;
      }
    const double v4diss = genint(attwk1.data(), nxta, nypa, 1.0, 0.5);
    const double vke = genint(attwk2.data(), nxta, nypa, 1.0, 0.5);
    const double vkedot = genint(attwk3.data(), nxta, nypa, 1.0, 0.5);
    const double pint = genint(&pa[IX3(1, 1, k, nxpa, nypa)], nxpa, nypa, 0.5, 0.5);
    const double qint = genint(&qa[IX3(1, 1, k, nxpa, nypa)], nxpa, nypa, 0.5, 0.5);
    r->pavgat[k - 1] = pint * atnorm;
    r->qavgat[k - 1] = qint * atnorm;
    r->ah4dat[k - 1] = rhoat * c.ah4at[k - 1] * c.hat[k - 1] * (u4diss + v4diss) * atnorm;
    r->kealat[k - 1] = 0.5 * rhoat * c.hat[k - 1] * (uke + vke) * atnorm;
    r->ddtkeat[k - 1] = rhoat * c.hat[k - 1] * (ukedot + vkedot) * atnorm;
  }
  // mixed layer (:424-451)
  r->tmlmat = genint(ast.data(), nxta, nyta, 1.0, 1.0);
  r->hmlmat = genint(hmixa.data(), nxta, nyta, 1.0, 1.0);
  r->astmin = 1.0e30;
  r->astmax = -1.0e30;
  for (int j = 1; j <= nyta; ++j)
    for (int i = 1; i <= nxta; ++i) {
      r->astmin = std::min(r->astmin, ast[IX2(i, j, nxta)]);
      r->astmax = std::max(r->astmax, ast[IX2(i, j, nxta)]);
      attwk1[IX2(i, j, nxta)] = ast[IX2(i, j, nxta)] * hmixa[IX2(i, j, nxta)];
    }
  r->hcmlat = genint(attwk1.data(), nxta, nyta, 1.0, 1.0);
  r->tmlmat = r->tmlmat * atnorm;
  r->hmlmat = r->hmlmat * atnorm;
  r->hcmlat = rhoat * cpat * r->hcmlat * atnorm;
  // mean over the ocean (:454-461)
  const int nxaooc = nxto / ndxr, nyaooc = nyto / ndxr;
  double tmaooc = 0.0;
  for (int j = ny1; j <= ny1 + nyaooc - 1; ++j)
    for (int i = nx1; i <= nx1 + nxaooc - 1; ++i) tmaooc = tmaooc + ast[IX2(i, j, nxta)];
  r->tmaooc = tmaooc / (double)(nxaooc * nyaooc);
  // outgoing long wave radiation (:464-470); davgat as in src/topsubs.F:429-430
  r->davgat = xintp(dtopat.data(), nxpa, nypa) * atnorm;
  double olrtop = c.Bup[nla - 1] * (r->hmlmat - c.hmat) + c.Cup[nla - 1] * r->davgat + c.Dup[nla - 1] * r->tmlmat;
  for (int i = 1; i <= nla - 1; ++i) olrtop = olrtop + c.Aup[(nla - 1) + nla * (i - 1)] * r->etamat[i - 1];
  r->olrtop = olrtop;
  courat(r);      // :475
}

// src/monitor_diag.F:1215-1445: as couroc for the periodic atmosphere; the mixed layer moves with
// the geostrophic wind of layer 1 plus the Ekman velocities uekat, vekat, which also set v on
// the zonal boundaries; the QG layers have v = 0 there
void Model::courat(qgcm_monitor_atmos *r) {
  for (int k = 0; k <= nla; ++k) {
    const bool ml = (k == 0);
    const int kk = ml ? 1 : k;
    auto uface = [&](int f, int j) {
      double u = -rdxaf0 * (pa[IX3(f, j + 1, kk, nxpa, nypa)] - pa[IX3(f, j, kk, nxpa, nypa)]);
      if (ml) u = u + uekat[IX2(f, j, nxpa)];
      return u;
    };
    auto vface = [&](int i, int jf) {
      if (jf == 1 || jf == nyta + 1) return ml ? vekat[IX2(i, jf, nxta)] : 0.0;
      double v = rdxaf0 * (pa[IX3(i + 1, jf, kk, nxpa, nypa)] - pa[IX3(i, jf, kk, nxpa, nypa)]);
      if (ml) v = v + vekat[IX2(i, jf, nxta)];
      return v;
    };
    double umin = 1.0e30, umax = -1.0e30, vmin = 1.0e30, vmax = -1.0e30, vsqmax = -1.0e30;
    for (int j = 1; j <= nyta; ++j) {
      double up = uface(1, j);
      umin = std::min(umin, up);
      umax = std::max(umax, up);
      for (int i = 1; i <= nxta; ++i) {
        const double um = up;
        up = uface(i + 1, j);
        const double vm = vface(i, j), vp = vface(i, j + 1);
        umin = std::min(umin, up);
        umax = std::max(umax, up);
        vmin = std::min(vmin, std::min(vm, vp));
        vmax = std::max(vmax, std::max(vm, vp));
        const double velsqd = (um + up) * (um + up) + (vm + vp) * (vm + vp);
        vsqmax = std::max(vsqmax, velsqd);
      }
    }
    if (ml) {
      r->umminat = umin; r->ummaxat = umax; r->vmminat = vmin; r->vmmaxat = vmax;
      r->cnmlat = hdxam1 * dta * std::sqrt(vsqmax);
    } else {
      r->ugminat[k - 1] = umin; r->ugmaxat[k - 1] = umax; r->vgminat[k - 1] = vmin; r->vgmaxat[k - 1] = vmax;
      r->cnqgat[k - 1] = hdxam1 * dta * std::sqrt(vsqmax);
    }
  }
}

}  // namespace orc

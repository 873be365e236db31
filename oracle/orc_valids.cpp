// TEST INFRASTRUCTURE ONLY -- CPU oracle (pinned against the translated reference, see orc_model.h).
// Restatement of the extreme-value scan and thickness check of valids
// (src/valsubs.F:43-630) without its diagnostic print-outs.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "orc_model.h"

namespace orc {

#define IX2(i, j, nx) ((size_t)((i)-1) + (size_t)(nx) * (size_t)((j)-1))
#define IX3(i, j, k, nx, ny) ((size_t)((i)-1) + (size_t)(nx) * ((size_t)((j)-1) + (size_t)(ny) * (size_t)((k)-1)))

static void scan(const vec &f, double &lo, double &hi) {
  lo = 1.0e30;     // bignum, src/valsubs.F:77
  hi = -1.0e30;
  for (double v : f) {
    lo = std::min(lo, v);
    hi = std::max(hi, v);
  }
}

void Model::valids(qgcm_valids_report *r) {
  // src/valsubs.F:77-81, :98-99
  const double tauext = 10.0, wtaext = 1.0, wtoext = 1.0e-3, astext = 90.0, patext = 1.0e7, qatext = 0.05;
  const double sstext = 75.0, pocext = 1.0e4, qocext = 0.05, thkmin = 100.0, critpc = 20.0;
  std::memset(r, 0, sizeof(*r));
  bool solnok = true;
  auto bad = [](double lo, double hi, double ext) { return std::fabs(lo) >= ext || std::fabs(hi) >= ext; };
  if (!ocean_only) {   // src/valsubs.F:118-262
    scan(pa, r->patmin, r->patmax);
    scan(qa, r->qatmin, r->qatmax);
    scan(ast, r->astmin, r->astmax);
    scan(wekta, r->wtamin, r->wtamax);
    scan(tauxa, r->txamin, r->txamax);
    scan(tauya, r->tyamin, r->tyamax);
    if (bad(r->patmin, r->patmax, patext)) solnok = false;
    if (bad(r->qatmin, r->qatmax, qatext)) solnok = false;
    if (bad(r->astmin, r->astmax, astext)) solnok = false;
    if (bad(r->wtamin, r->wtamax, wtaext)) solnok = false;
    if (bad(r->txamin, r->txamax, tauext) || bad(r->tyamin, r->tyamax, tauext)) solnok = false;
  }
  if (!atmos_only) {   // src/valsubs.F:264-524
    scan(po, r->pocmin, r->pocmax);
    scan(qo, r->qocmin, r->qocmax);
    scan(sst, r->sstmin, r->sstmax);
    scan(wekto, r->wtomin, r->wtomax);
    if (bad(r->pocmin, r->pocmax, pocext)) solnok = false;
    if (bad(r->qocmin, r->qocmax, qocext)) solnok = false;
    if (bad(r->sstmin, r->sstmax, sstext)) solnok = false;
    if (bad(r->wtomin, r->wtomax, wtoext)) solnok = false;
    double rgpoc[QGCM_NLMAX], etaoc[QGCM_NLMAX], hfbad[QGCM_NLMAX];
    for (int k = 1; k <= nlo - 1; ++k) rgpoc[k - 1] = 1.0 / c.gpoc[k - 1];
    double hfmint = 1.0e30, hfmaxt = -1.0e30, hfmini = 1.0e30, hfmaxi = -1.0e30, hfminb = 1.0e30, hfmaxb = -1.0e30;
    // dtopoc = H_nlo/f0 * ddynoc (src/topsubs.F:454)
    const double dtopfac = c.hoc[nlo - 1] / fnot;
    for (int pass = 0; pass < 2; ++pass) {
      const double hfmina = std::min(hfmint, std::min(hfmini, hfminb));
      if (pass == 1) {
        for (int k = 0; k < nlo; ++k) hfbad[k] = 0.0;
        if (!(hfmina <= thkmin)) break;
      }
      for (int j = 1; j <= nypo; ++j) {
        const double wtj = (j == 1 || j == nypo) ? 0.5 : 1.0;
        for (int i = 1; i <= nxpo; ++i) {
          const double wti = (i == 1 || i == nxpo) ? 0.5 : 1.0;
          for (int k = 1; k <= nlo - 1; ++k)
            etaoc[k - 1] = rgpoc[k - 1] * (po[IX3(i, j, k + 1, nxpo, nypo)] - po[IX3(i, j, k, nxpo, nypo)]);
          double hfull = c.hoc[0] - etaoc[0];
          if (pass == 0) { hfmint = std::min(hfmint, hfull); hfmaxt = std::max(hfmaxt, hfull); }
          else if (hfull < thkmin) hfbad[0] += wti * wtj;
          for (int k = 2; k <= nlo - 1; ++k) {
            hfull = c.hoc[k - 1] - etaoc[k - 1] + etaoc[k - 2];
            if (pass == 0) { hfmini = std::min(hfmini, hfull); hfmaxi = std::max(hfmaxi, hfull); }
            else if (hfull < thkmin) hfbad[k - 1] += wti * wtj;
          }
          hfull = c.hoc[nlo - 1] + etaoc[nlo - 2] - dtopfac * ddynoc[IX2(i, j, nxpo)];
          if (pass == 0) { hfminb = std::min(hfminb, hfull); hfmaxb = std::max(hfmaxb, hfull); }
          else if (hfull < thkmin) hfbad[nlo - 1] += wti * wtj;
        }
      }
    }
    r->hfmint = hfmint; r->hfmaxt = hfmaxt; r->hfmini = hfmini; r->hfmaxi = hfmaxi; r->hfminb = hfminb; r->hfmaxb = hfmaxb;
    bool pcfail = false;
    for (int k = 0; k < nlo; ++k) {
      r->hfbad[k] = 100.0 * hfbad[k] * ocnorm;
      if (r->hfbad[k] > critpc) pcfail = true;
    }
    if (pcfail) solnok = false;   // spfail = .false. (src/valsubs.F:98)
  }
  r->solnok = solnok ? 1 : 0;
}

}  // namespace orc

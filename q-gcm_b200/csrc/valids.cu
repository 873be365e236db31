// Device-side validity scan: valids (src/valsubs.F:43-630).  Extreme values of the prognostic
// and forcing fields against the reference's thresholds (:77-81) and the perturbed ocean
// layer thicknesses (:380-524).  Called every 0.25 model days (src/q-gcm.F:1278), so the
// kernels are plain grid-stride reductions; only a few hundred partials come back to the host.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "qgcm_internal.h"

namespace qg {

constexpr int VB = 296;      // blocks per reduction (2 per SM)

// block-wide min and max; result valid in thread 0
__device__ __forceinline__ void block_minmax(double &lo, double &hi, double (*red)[8]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { red[0][w] = lo; red[1][w] = hi; }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 1; i < 8; ++i) { lo = fmin(lo, red[0][i]); hi = fmax(hi, red[1][i]); }
}

// min/max of field rows [j0, j1), columns [0, nx), nl layers; out[2*b] = min, out[2*b+1] = max
__global__ void __launch_bounds__(256) k_minmax(const double *f, int nx, int j0, int j1, int nl, int ld, size_t lsz, double *out) {
  __shared__ double red[2][8];
  double lo = 1.0e30, hi = -1.0e30;     // bignum, src/valsubs.F:77
  const int rows = j1 - j0;
  for (int k = 0; k < nl; ++k)
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
      const double *row = f + (size_t)k * lsz + (size_t)(j0 + r) * ld;
      for (int i = threadIdx.x; i < nx; i += 256) {
        const double v = row[i];
        lo = fmin(lo, v);
        hi = fmax(hi, v);
      }
    }
  block_minmax(lo, hi, red);
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = lo; out[2 * blockIdx.x + 1] = hi; }
}

struct ThickArgs {
  Grid g;
  int nl, j0, j1;
  double h[NLMAX], rgp[NLMAX];
  double dtopfac;            // dtopoc = dtopfac * ddynoc  (ddynoc = f0*dtopoc/H_nlo, src/topsubs.F:454)
  double thkmin;
  const double *p, *ddyn;
  double *out;               // [VB][6 + NLMAX]: min/max top, intermediate, bottom; weighted count of thin points per layer
};

// full layer thicknesses h_k - eta_k + eta_{k-1} (src/valsubs.F:403-422) and the area-weighted
// count of points thinner than thkmin (:437-474)
__global__ void __launch_bounds__(256) k_thickness(ThickArgs a) {
  __shared__ double red[2][8];
  __shared__ double cnt[NLMAX][8];
  const Grid &g = a.g;
  const int nl = a.nl;
  double mt = 1.0e30, xt = -1.0e30, mi = 1.0e30, xi = -1.0e30, mb = 1.0e30, xb = -1.0e30;
  double bad[NLMAX];
  for (int k = 0; k < NLMAX; ++k) bad[k] = 0.0;
  for (int j = a.j0 + blockIdx.x; j < a.j1; j += gridDim.x) {
    const int jg = g.jg0 + j;
    const double wtj = (jg == 0 || jg == g.nyp_g - 1) ? 0.5 : 1.0;
    for (int i = threadIdx.x; i < g.nxp; i += 256) {
      const double wti = (i == 0 || i == g.nxp - 1) ? 0.5 : 1.0;
      const size_t c = (size_t)j * g.ld + i;
      double eta[NLMAX];
      for (int k = 0; k < nl - 1; ++k) eta[k] = a.rgp[k] * (a.p[(k + 1) * g.lsz + c] - a.p[k * g.lsz + c]);
      double hf = a.h[0] - eta[0];
      mt = fmin(mt, hf); xt = fmax(xt, hf);
      if (hf < a.thkmin) bad[0] += wti * wtj;
      for (int k = 1; k < nl - 1; ++k) {
        hf = a.h[k] - eta[k] + eta[k - 1];
        mi = fmin(mi, hf); xi = fmax(xi, hf);
        if (hf < a.thkmin) bad[k] += wti * wtj;
      }
      hf = a.h[nl - 1] + eta[nl - 2] - a.dtopfac * a.ddyn[c];
      mb = fmin(mb, hf); xb = fmax(xb, hf);
      if (hf < a.thkmin) bad[nl - 1] += wti * wtj;
    }
  }
  double *o = a.out + (size_t)blockIdx.x * (6 + NLMAX);
  block_minmax(mt, xt, red); if (threadIdx.x == 0) { o[0] = mt; o[1] = xt; }
  block_minmax(mi, xi, red); if (threadIdx.x == 0) { o[2] = mi; o[3] = xi; }
  block_minmax(mb, xb, red); if (threadIdx.x == 0) { o[4] = mb; o[5] = xb; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = 0; k < nl; ++k) {
    double v = bad[k];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
    if (lane == 0) cnt[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < nl; ++k) {
      double t = 0.0;
      for (int i = 0; i < 8; ++i) t += cnt[k][i];
      o[6 + k] = t;
    }
}

static void minmax(qgcm_model *m, const char *name, int j0, int j1, double *lo, double *hi) {
  qgcm_model::Field &f = m->fields.at(name);
  const int nb = std::min(VB, std::max(1, j1 - j0));
  QG_LAUNCH(m, "k_minmax", nb, 256, 0, k_minmax, f.d, f.nx, j0, j1, f.nl, f.ld, f.lsz, m->d_val);
  std::vector<double> h(2 * nb);
  QG_CUDA(cudaMemcpyAsync(h.data(), m->d_val, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
  *lo = 1.0e30; *hi = -1.0e30;
  for (int b = 0; b < nb; ++b) { *lo = std::min(*lo, h[2 * b]); *hi = std::max(*hi, h[2 * b + 1]); }
}

void launch_valids(qgcm_model *m, qgcm_valids_report *r) {
  // thresholds, src/valsubs.F:77-81, :98-99
  const double tauext = 10.0, wtaext = 1.0, wtoext = 1.0e-3, astext = 90.0, patext = 1.0e7, qatext = 0.05;
  const double sstext = 75.0, pocext = 1.0e4, qocext = 0.05, thkmin = 100.0, critpc = 20.0;
  std::memset(r, 0, sizeof(*r));
  if (!m->d_val) m->d_val = (double *)dalloc(m, sizeof(double) * VB * (6 + NLMAX));
  bool ok = true;
  auto bad = [](double lo, double hi, double ext) { return std::fabs(lo) >= ext || std::fabs(hi) >= ext; };
  if (m->has_atmos) {
    const Grid &g = m->ga;
    minmax(m, "pa", 0, g.nyp, &r->patmin, &r->patmax);
    minmax(m, "qa", 0, g.nyp, &r->qatmin, &r->qatmax);
    minmax(m, "ast", 0, g.nyt, &r->astmin, &r->astmax);
    minmax(m, "wekta", 0, g.nyt, &r->wtamin, &r->wtamax);
    minmax(m, "tauxa", 0, g.nyp, &r->txamin, &r->txamax);
    minmax(m, "tauya", 0, g.nyp, &r->tyamin, &r->tyamax);
    if (bad(r->patmin, r->patmax, patext) || bad(r->qatmin, r->qatmax, qatext) || bad(r->astmin, r->astmax, astext) ||
        bad(r->wtamin, r->wtamax, wtaext) || bad(r->txamin, r->txamax, tauext) || bad(r->tyamin, r->tyamax, tauext))
      ok = false;
  }
  if (m->has_ocean) {
    const Grid &g = m->go;
    const int p0 = g.own0, p1 = g.own1, t1 = std::min(g.own1, g.nyt);
    minmax(m, "po", p0, p1, &r->pocmin, &r->pocmax);
    minmax(m, "qo", p0, p1, &r->qocmin, &r->qocmax);
    minmax(m, "sst", p0, t1, &r->sstmin, &r->sstmax);
    minmax(m, "wekto", p0, t1, &r->wtomin, &r->wtomax);
    if (bad(r->pocmin, r->pocmax, pocext) || bad(r->qocmin, r->qocmax, qocext) || bad(r->sstmin, r->sstmax, sstext) ||
        bad(r->wtomin, r->wtomax, wtoext))
      ok = false;
    ThickArgs a;
    a.g = g; a.nl = g.nl; a.j0 = p0; a.j1 = p1;
    for (int k = 0; k < NLMAX; ++k) { a.h[k] = m->lo.h[k]; a.rgp[k] = (k < g.nl - 1) ? 1.0 / m->lo.gp[k] : 0.0; }
    a.dtopfac = m->lo.h[g.nl - 1] / m->fnot;
    a.thkmin = thkmin;
    a.p = m->F("po"); a.ddyn = m->F("ddynoc");
    a.out = m->d_val;
    const int nb = std::min(VB, std::max(1, p1 - p0));
    QG_LAUNCH(m, "k_thickness", nb, 256, 0, k_thickness, a);
    std::vector<double> h((size_t)nb * (6 + NLMAX));
    QG_CUDA(cudaMemcpyAsync(h.data(), m->d_val, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, m->stream));
    QG_CUDA(cudaStreamSynchronize(m->stream));
    double mn[3] = {1.0e30, 1.0e30, 1.0e30}, mx[3] = {-1.0e30, -1.0e30, -1.0e30}, cnt[NLMAX] = {0};
    for (int b = 0; b < nb; ++b) {
      const double *o = &h[(size_t)b * (6 + NLMAX)];
      for (int q = 0; q < 3; ++q) { mn[q] = std::min(mn[q], o[2 * q]); mx[q] = std::max(mx[q], o[2 * q + 1]); }
      for (int k = 0; k < g.nl; ++k) cnt[k] += o[6 + k];
    }
    r->hfmint = mn[0]; r->hfmaxt = mx[0]; r->hfmini = mn[1]; r->hfmaxi = mx[1]; r->hfminb = mn[2]; r->hfmaxb = mx[2];
    const double hfmina = std::min(mn[0], std::min(mn[1], mn[2]));
    // the thin-point census is only taken when some thickness is at or below thkmin (src/valsubs.F:431)
    bool pcfail = false;
    for (int k = 0; k < g.nl; ++k) {
      r->hfbad[k] = (hfmina <= thkmin) ? 100.0 * cnt[k] * g.norm : 0.0;
      if (r->hfbad[k] > critpc) pcfail = true;
    }
    if (pcfail) ok = false;      // spfail = .false.: percentage criterion (src/valsubs.F:505-524)
  }
  r->solnok = ok ? 1 : 0;
}

}  // namespace qg

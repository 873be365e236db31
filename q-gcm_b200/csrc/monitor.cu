// Device-side ocean section of monnc_comp (src/monitor_diag.F:480-840): every
// diagnostic of the module `monitor` that monnc_comp derives from the ocean state is an area
// integral (genint, :1160-1210), an extremum or a row statistic.  The device produces, per
// grid row, the x-sums with genint's W/E weights (one block per row, fixed summation order)
// and the row extrema; the host adds the rows in genint's order (interior rows, then
// facsn*(south + north)) and applies the scalar factors.  Only Q x nyp doubles cross PCIe
// (1.9 MB at 1 km) instead of the 2.6 GB of fields the Fortran routine reads.
// The lagged velocities and their one-sided-boundary Laplacians (del4bx :900-1015, del4ch
// :1020-1155) live in four scratch fields of the monitor's own, allocated on first use.
#include <cmath>
#include <cstring>

#include "qgcm_internal.h"

namespace qg {

// fixed-order block reduction of K running sums; thread 0 stores them to out[q*pitch]
template <int K>
__device__ __forceinline__ void row_store(double (&v)[K], double *out, size_t pitch) {
  __shared__ double red[K][8];
#pragma unroll
  for (int q = 0; q < K; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < K; ++q) red[q][w] = v[q];
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    out[(size_t)threadIdx.x * pitch] = s;
  }
  __syncthreads();
}
__device__ __forceinline__ void row_minmax(double lo, double hi, double *out_lo, double *out_hi) {
  __shared__ double red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_down_sync(0xffffffffu, hi, o));
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[0][w] = lo; red[1][w] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { lo = fmin(lo, red[0][i]); hi = fmax(hi, red[1][i]); }
    *out_lo = lo;
    *out_hi = hi;
  }
  __syncthreads();
}

// T grid: sums of wekto, |wekto|, sst*wekto, sst; extrema of sst  (:498-509, :788-806)
__global__ void __launch_bounds__(256) k_mon_t(const double *wekt, const double *sst, int nxt, int ld, double *out, size_t pitch) {
  const int j = blockIdx.x;
  double v[4] = {0.0, 0.0, 0.0, 0.0}, lo = 1.0e30, hi = -1.0e30;
  for (int i = threadIdx.x; i < nxt; i += 256) {
    const double w = wekt[(size_t)j * ld + i], t = sst[(size_t)j * ld + i];
    v[0] += w; v[1] += fabs(w); v[2] += t * w; v[3] += t;
    lo = fmin(lo, t); hi = fmax(hi, t);
  }
  row_store<4>(v, out + j, pitch);
  row_minmax(lo, hi, out + 4 * pitch + j, out + 5 * pitch + j);
}

// p grid, W/E weight 0.5: wekpo, |wekpo|, entoc, |entoc|  (:511-543)
__global__ void __launch_bounds__(256) k_mon_p(const double *wekp, const double *ent, int nxp, int ld, double *out, size_t pitch) {
  const int j = blockIdx.x;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nxp; i += 256) {
    const double f = (i == 0 || i == nxp - 1) ? 0.5 : 1.0;
    const double w = wekp[(size_t)j * ld + i], e = ent[(size_t)j * ld + i];
    v[0] += f * w; v[1] += f * fabs(w); v[2] += f * e; v[3] += f * fabs(e);
  }
  row_store<4>(v, out + j, pitch);
}

// interface k (between layers k and k+1), p grid: eta, eta^2, eta*etadot, eta*entoc  (:547-583)
__global__ void __launch_bounds__(256) k_mon_eta(const double *p0, const double *p1, const double *pm0, const double *pm1, const double *ent,
                                                 double rgp, double rgpdt, int nxp, int ld, double *out, size_t pitch) {
  const int j = blockIdx.x;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nxp; i += 256) {
    const size_t o = (size_t)j * ld + i;
    const double f = (i == 0 || i == nxp - 1) ? 0.5 : 1.0;
    const double eta = rgp * (p1[o] - p0[o]);
    const double etadot = rgpdt * (p0[o] - p1[o] - pm0[o] + pm1[o]);
    v[0] += f * eta; v[1] += f * (eta * eta); v[2] += f * (eta * etadot); v[3] += f * (eta * ent[o]);
  }
  row_store<4>(v, out + j, pitch);
}

// layer k, p grid: sums of po, qo; extrema of po  (:666-675, :716-717)
__global__ void __launch_bounds__(256) k_mon_pq(const double *p, const double *q, int nxp, int ld, double *out, size_t pitch) {
  const int j = blockIdx.x;
  double v[2] = {0.0, 0.0}, lo = 1.0e30, hi = -1.0e30;
  for (int i = threadIdx.x; i < nxp; i += 256) {
    const size_t o = (size_t)j * ld + i;
    const double f = (i == 0 || i == nxp - 1) ? 0.5 : 1.0;
    v[0] += f * p[o]; v[1] += f * q[o];
    lo = fmin(lo, p[o]); hi = fmax(hi, p[o]);
  }
  row_store<2>(v, out + j, pitch);
  row_minmax(lo, hi, out + 2 * pitch + j, out + 3 * pitch + j);
}

// lagged geostrophic velocities of one layer: u on (nxp, nyt), v on (nxt, nyp)  (:624-630, :641-647)
__global__ void __launch_bounds__(256) k_mon_geo(const double *pm, double *ug, double *vg, Grid g, double rdxf0) {
  const int i = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
  if (i >= g.nxp) return;
  const size_t o = (size_t)j * g.ld + i;
  if (j < g.nyt) ug[o] = -rdxf0 * (pm[o + g.ld] - pm[o]);
  if (i < g.nxt) vg[o] = rdxf0 * (pm[o + 1] - pm[o]);
}

// one Laplacian of del4bx / del4ch on an (nx, ny) array: centred inside, one-sided second
// differences on solid boundaries, period nx in a channel
__global__ void __launch_bounds__(256) k_mon_lap(const double *__restrict__ a, double *__restrict__ d, int nx, int ny, int ld, int cyclic,
                                                 double dxm2) {
  const int i = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
  if (i >= nx) return;
  const double *r = a + (size_t)j * ld;
  double v;
  const bool inner_j = (j > 0 && j < ny - 1);
  const bool inner_i = (i > 0 && i < nx - 1);
  if (inner_j && (inner_i || cyclic)) {
    const int im = (i == 0) ? nx - 1 : i - 1, ip = (i == nx - 1) ? 0 : i + 1;
    v = dxm2 * (r[i - ld] + r[im] + r[ip] + r[i + ld] - 4.0 * r[i]);
  } else {
    // x part
    double s;
    if (inner_i || cyclic) {
      const int im = (i == 0) ? nx - 1 : i - 1, ip = (i == nx - 1) ? 0 : i + 1;
      s = r[im] - 2.0 * r[i] + r[ip];
    } else if (i == 0) {
      s = r[2] - 2.0 * r[1] + r[0];
    } else {
      s = r[nx - 1] - 2.0 * r[nx - 2] + r[nx - 3];
    }
    // y part, in the reference's order of terms
    if (inner_j) s = s + r[i - ld] - 2.0 * r[i] + r[i + ld];
    else if (j == 0) s = s + r[i + 2 * ld] - 2.0 * r[i + ld] + r[i];
    else s = s + r[i] - 2.0 * r[i - ld] + r[i - 2 * ld];
    v = dxm2 * s;
  }
  d[(size_t)j * ld + i] = v;
}

// layer k, u points (nxp, nyt), W/E weight 0.5: ug*del2(ugm), ug*del4(ugm), ug^2, ug*ugdot, ugm^2,
// ug*tauxav, and the zonal jet sum  (:593-601, :676-694, :759-766)
__global__ void __launch_bounds__(256) k_mon_u(const double *p, const double *pm, const double *ugm, const double *d2, const double *d4,
                                               const double *taux, double rdxf0, double rdxf0dt, int nxp, int ld, double *out, size_t pitch) {
  const int j = blockIdx.x;
  double v[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nxp; i += 256) {
    const size_t o = (size_t)j * ld + i;
    const double f = (i == 0 || i == nxp - 1) ? 0.5 : 1.0;
    const double ug = -rdxf0 * (p[o + ld] - p[o]);
    const double ugdot = -rdxf0dt * (p[o + ld] - pm[o] - pm[o + ld] + pm[o]);     // as written at :676-677
    v[0] += f * (ug * d2[o]); v[1] += f * (ug * d4[o]); v[2] += f * (ug * ug); v[3] += f * (ug * ugdot);
    v[4] += f * (ugm[o] * ugm[o]);
    v[5] += f * (ug * (0.5 * (taux[o + ld] + taux[o])));
    if (i < nxp - 1) v[6] += ug;            // the end columns are equal: counted once (:690-691)
  }
  row_store<7>(v, out + j, pitch);
}

// layer k, v points (nxt, nyp), W/E weight 1  (:606-614, :700-712, :771-778)
__global__ void __launch_bounds__(256) k_mon_v(const double *p, const double *pm, const double *vgm, const double *d2, const double *d4,
                                               const double *tauy, double rdxf0, double rdxf0dt, int nxt, int ld, double *out, size_t pitch) {
  const int j = blockIdx.x;
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nxt; i += 256) {
    const size_t o = (size_t)j * ld + i;
    const double vg = rdxf0 * (p[o + 1] - p[o]);
    const double vgdot = rdxf0dt * (p[o + 1] - p[o] - pm[o + 1] + pm[o]);
    v[0] += vg * d2[o]; v[1] += vg * d4[o]; v[2] += vg * vg; v[3] += vg * vgdot;
    v[4] += vgm[o] * vgm[o];
    v[5] += vg * (0.5 * (tauy[o + 1] + tauy[o]));
  }
  row_store<6>(v, out + j, pitch);
}

// genint's outer sum over the row sums (:1180-1207)
static double rows_int(const double *row, int ny, double facsn) {
  double answer = 0.0;
  for (int j = 1; j < ny - 1; ++j) answer = answer + row[j];
  return answer + facsn * (row[0] + row[ny - 1]);
}

void launch_monnc_ocean(qgcm_model *m, qgcm_monitor_ocean *r) {
  std::memset(r, 0, sizeof(*r));
  if (!m->has_ocean) return;
  if (m->nranks > 1) throw std::runtime_error("qgcm_monnc_ocean: the reductions are not combined across y-slabs yet (single GPU)");
  const Grid &g = m->go;
  const qgcm_config &c = m->cfg;
  const int nl = g.nl, nxp = g.nxp, nyp = g.nyp, nxt = g.nxt, nyt = g.nyt, ld = g.ld;
  const size_t pitch = nyp;
  // row-sum slots: T 6 | P 4 | eta 4 per interface | pq 4 per layer | u 7 per layer | v 6 per layer
  const int oT = 0, oP = 6, oE = 10, oQ = oE + 4 * (nl - 1), oU = oQ + 4 * nl, oV = oU + 7 * nl, nslot = oV + 6 * nl;
  const size_t need = (size_t)nslot * pitch;
  if (m->mon_elems < need) {
    m->d_mon = (double *)dalloc(m, sizeof(double) * need);
    m->mon_elems = need;
  }
  double *rows = m->d_mon;
  const double *po = m->F("po"), *pom = m->F("pom"), *qo = m->F("qo");
  const double rdxf0 = g.rdxf0, dto = m->dto;
  QG_LAUNCH(m, "k_mon_t", nyt, 256, 0, k_mon_t, m->F("wekto"), m->F("sst"), nxt, ld, rows + oT * pitch, pitch);
  QG_LAUNCH(m, "k_mon_p", nyp, 256, 0, k_mon_p, m->F("wekpo"), m->F("entoc"), nxp, ld, rows + oP * pitch, pitch);
  for (int k = 0; k < nl - 1; ++k) {
    const double rgp = 1.0 / m->lo.gp[k];
    QG_LAUNCH(m, "k_mon_eta", nyp, 256, 0, k_mon_eta, po + (size_t)k * g.lsz, po + (size_t)(k + 1) * g.lsz, pom + (size_t)k * g.lsz,
              pom + (size_t)(k + 1) * g.lsz, m->F("entoc"), rgp, rgp / dto, nxp, ld, rows + (size_t)(oE + 4 * k) * pitch, pitch);
  }
  if (!m->d_monf) m->d_monf = (double *)dalloc(m, sizeof(double) * 4 * g.lsz);
  double *ugm = m->d_monf, *vgm = ugm + g.lsz, *d2 = vgm + g.lsz, *d4 = d2 + g.lsz;
  const dim3 full((nxp + 255) / 256, nyp);
  for (int k = 0; k < nl; ++k) {
    const double *pk = po + (size_t)k * g.lsz, *pmk = pom + (size_t)k * g.lsz;
    QG_LAUNCH(m, "k_mon_pq", nyp, 256, 0, k_mon_pq, pk, qo + (size_t)k * g.lsz, nxp, ld, rows + (size_t)(oQ + 4 * k) * pitch, pitch);
    QG_LAUNCH(m, "k_mon_geo", full, 256, 0, k_mon_geo, pmk, ugm, vgm, g, rdxf0);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxp + 255) / 256, nyt), 256, 0, k_mon_lap, ugm, d2, nxp, nyt, ld, g.cyclic, g.dxm2);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxp + 255) / 256, nyt), 256, 0, k_mon_lap, d2, d4, nxp, nyt, ld, g.cyclic, g.dxm2);
    QG_LAUNCH(m, "k_mon_u", nyt, 256, 0, k_mon_u, pk, pmk, ugm, d2, d4, m->F("tauxo"), rdxf0, rdxf0 / dto, nxp, ld,
              rows + (size_t)(oU + 7 * k) * pitch, pitch);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxt + 255) / 256, nyp), 256, 0, k_mon_lap, vgm, d2, nxt, nyp, ld, g.cyclic, g.dxm2);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxt + 255) / 256, nyp), 256, 0, k_mon_lap, d2, d4, nxt, nyp, ld, g.cyclic, g.dxm2);
    QG_LAUNCH(m, "k_mon_v", nyp, 256, 0, k_mon_v, pk, pmk, vgm, d2, d4, m->F("tauyo"), rdxf0, rdxf0 / dto, nxt, ld,
              rows + (size_t)(oV + 6 * k) * pitch, pitch);
  }
  std::vector<double> h(need);
  double corner[2 * NLMAX];       // po(1,1,k), po(1,nypo,k)
  QG_CUDA(cudaMemcpyAsync(h.data(), rows, sizeof(double) * need, cudaMemcpyDeviceToHost, m->stream));
  for (int k = 0; k < nl; ++k) {
    QG_CUDA(cudaMemcpyAsync(&corner[2 * k], po + (size_t)k * g.lsz, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    QG_CUDA(cudaMemcpyAsync(&corner[2 * k + 1], po + (size_t)k * g.lsz + (size_t)(nyp - 1) * ld, sizeof(double), cudaMemcpyDeviceToHost,
                            m->stream));
  }
  QG_CUDA(cudaStreamSynchronize(m->stream));
  auto row = [&](int slot) { return h.data() + (size_t)slot * pitch; };
  const double ocnorm = g.norm, rhooc = c.rhooc, fnot = m->fnot;
  // Ekman velocity, entrainment (:498-543)
  r->wetmoc = rows_int(row(oT + 0), nyt, 1.0) * ocnorm;
  r->watmoc = rows_int(row(oT + 1), nyt, 1.0) * ocnorm;
  r->wepmoc = rows_int(row(oP + 0), nyp, 0.5) * ocnorm;
  r->wapmoc = rows_int(row(oP + 1), nyp, 0.5) * ocnorm;
  r->entmoc = rows_int(row(oP + 2), nyp, 0.5) * ocnorm;
  r->enamoc = rows_int(row(oP + 3), nyp, 0.5) * ocnorm;
  // interface displacements (:547-583)
  for (int k = 0; k < nl - 1; ++k) {
    r->etamoc[k] = rows_int(row(oE + 4 * k), nyp, 0.5) * ocnorm;
    r->et2moc[k] = rows_int(row(oE + 4 * k + 1), nyp, 0.5) * ocnorm;
    r->ddtpeoc[k] = rhooc * m->lo.gp[k] * rows_int(row(oE + 4 * k + 2), nyp, 0.5);
    if (k == 0) r->pkenoc = rhooc * m->lo.gp[0] * rows_int(row(oE + 3), nyp, 0.5) * ocnorm;
  }
  // wind work (:588-617)
  {
    const double utaux = rows_int(row(oU + 5), nyt, 1.0), vtauy = rows_int(row(oV + 5), nyp, 0.5);
    r->utauoc = rhooc * (vtauy + utaux) * ocnorm;
  }
  // layers (:620-752)
  for (int k = 0; k < nl; ++k) {
    const double hk = m->lo.h[k];
    const double *u = row(oU + 7 * k), *v = row(oV + 6 * k);
    const double u2diss = rows_int(u, nyt, 1.0), u4diss = rows_int(u + pitch, nyt, 1.0);
    const double uke = rows_int(u + 2 * pitch, nyt, 1.0), ukedot = rows_int(u + 3 * pitch, nyt, 1.0);
    const double v2diss = rows_int(v, nyp, 0.5), v4diss = rows_int(v + pitch, nyp, 0.5);
    const double vke = rows_int(v + 2 * pitch, nyp, 0.5), vkedot = rows_int(v + 3 * pitch, nyp, 0.5);
    r->pavgoc[k] = rows_int(row(oQ + 4 * k), nyp, 0.5) * ocnorm;
    r->qavgoc[k] = rows_int(row(oQ + 4 * k + 1), nyp, 0.5) * ocnorm;
    r->ah2doc[k] = -rhooc * m->lo.ah2[k] * hk * (u2diss + v2diss) * ocnorm;
    r->ah4doc[k] = rhooc * m->lo.ah4[k] * hk * (u4diss + v4diss) * ocnorm;
    r->kealoc[k] = 0.5 * rhooc * hk * (uke + vke) * ocnorm;
    r->ddtkeoc[k] = rhooc * hk * (ukedot + vkedot) * ocnorm;
    // jet position: largest |zonal mean u| (:696-704)
    const double *uj = u + 6 * pitch;
    r->ocjpos[k] = 0;
    r->ocjval[k] = 0.0;
    for (int j = 0; j < nyt; ++j) {
      const double val = std::fabs(uj[j]) / (double)nxt;
      if (val > r->ocjval[k]) { r->ocjpos[k] = j + 1; r->ocjval[k] = val; }
    }
    double pomin = 1.0e30, pomax = -1.0e30;
    const double *lo = row(oQ + 4 * k + 2), *hi = row(oQ + 4 * k + 3);
    for (int j = 0; j < nyp; ++j) { pomin = std::min(pomin, lo[j]); pomax = std::max(pomax, hi[j]); }
    const double poref = (fnot > 0.0) ? corner[2 * k] : corner[2 * k + 1];
    double psiext = std::min(pomin / fnot, pomax / fnot);
    r->osfmin[k] = 1.0e-6 * hk * (psiext - poref / fnot);
    psiext = std::max(pomin / fnot, pomax / fnot);
    r->osfmax[k] = 1.0e-6 * hk * (psiext - poref / fnot);
    r->occirc[k] = 1.0e-6 * hk * (corner[2 * k] - corner[2 * k + 1]) / fnot;
  }
  // bottom drag (:755-783)
  {
    const double u2 = rows_int(row(oU + 7 * (nl - 1) + 4), nyt, 1.0), v2 = rows_int(row(oV + 6 * (nl - 1) + 4), nyp, 0.5);
    r->btdgoc = 0.5 * rhooc * c.delek * std::fabs(fnot) * (u2 + v2) * ocnorm;
  }
  // mixed layer (:788-812)
  r->sstmin = 1.0e30;
  r->sstmax = -1.0e30;
  for (int j = 0; j < nyt; ++j) {
    r->sstmin = std::min(r->sstmin, row(oT + 4)[j]);
    r->sstmax = std::max(r->sstmax, row(oT + 5)[j]);
  }
  r->hfmloc = rhooc * c.cpoc * rows_int(row(oT + 2), nyt, 1.0) * ocnorm;
  r->tmlmoc = rows_int(row(oT + 3), nyt, 1.0) * ocnorm;
  r->occtot = 0.0;
  for (int k = 0; k < nl; ++k) r->occtot = r->occtot + r->occirc[k];
}

}  // namespace qg

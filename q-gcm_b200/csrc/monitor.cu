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
                                               const double *taux, double rdxf0, double rdxf0dt, int nxp, int ld, int atmos, double *out,
                                               size_t pitch) {
  const int j = blockIdx.x;
  double v[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nxp; i += 256) {
    const size_t o = (size_t)j * ld + i;
    const double f = (i == 0 || i == nxp - 1) ? 0.5 : 1.0;
    const double ug = -rdxf0 * (p[o + ld] - p[o]);
    // ocean: as written at :676-677 (its two pom(i,j,k) terms cancel); atmosphere: :354-355
    const double ugdot = atmos ? -rdxf0dt * (p[o + ld] - p[o] - pm[o + ld] + pm[o]) : -rdxf0dt * (p[o + ld] - pm[o] - pm[o + ld] + pm[o]);
    v[0] += f * (ug * d2[o]); v[1] += f * (ug * d4[o]); v[2] += f * (ug * ug); v[3] += f * (ug * ugdot);
    v[4] += f * (ugm[o] * ugm[o]);
    v[5] += f * (ug * (0.5 * (taux[o + ld] + taux[o])));
    if (i < nxp - 1) v[6] += ug;            // the end columns are equal: counted once (:690-691)
  }
  row_store<7>(v, out + j, pitch);
}

// layer k, v points (nxt, nyp), W/E weight 1  (:606-614, :700-712, :771-778)
__global__ void __launch_bounds__(256) k_mon_v(const double *p, const double *pm, const double *vgm, const double *d2, const double *d4,
                                               const double *tauy, double rdxf0, double rdxf0dt, int nxt, int ld, int atmos, double *out,
                                               size_t pitch) {
  const int j = blockIdx.x;
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nxt; i += 256) {
    const size_t o = (size_t)j * ld + i;
    const double vg = rdxf0 * (p[o + 1] - p[o]);
    const double vgdot = rdxf0dt * (p[o + 1] - p[o] - pm[o + 1] + pm[o]);
    // atmosphere: the reference integrates attwk3, which still holds del-sqd of the lagged v (:391, :404)
    v[0] += vg * d2[o]; v[1] += vg * d4[o]; v[2] += vg * vg; v[3] += atmos ? d2[o] : vg * vgdot;
    v[4] += vgm[o] * vgm[o];
    v[5] += vg * (0.5 * (tauy[o + 1] + tauy[o]));
  }
  row_store<6>(v, out + j, pitch);
}

// couroc (:1450-1925): face velocities of one layer on T row j -- western/eastern faces um, up
// and southern/northern faces vm, vp of every cell, with the boundary faces set by the
// configuration -- reduced to the row's extrema of u and v and its largest (um+up)^2 + (vm+vp)^2.
// ekman: the mixed layer (po(:,:,1) scaled by ycexp plus the Ekman drift); else a QG layer.
// uek, vek non-null: the atmosphere's mixed layer (courat :1240-1310), which adds the Ekman
// velocities at the faces and keeps vekat on the zonal boundaries.
__global__ void __launch_bounds__(256) k_mon_cour(const double *p, const double *taux, const double *tauy, const double *uek,
                                                  const double *vek, Grid g, int ekman, int sflux, int nflux, double ug, double rh,
                                                  double *out, size_t pitch) {
  const int j = blockIdx.x, ld = g.ld, nxt = g.nxt;
  const bool south = (j == 0 && g.wall_s()), north = (j == g.nyt - 1 && g.wall_n());
  const double *p0 = p + (size_t)j * ld, *p1 = p0 + ld;
  const double *tx0 = taux + (size_t)j * ld, *tx1 = tx0 + ld, *ty0 = tauy + (size_t)j * ld, *ty1 = ty0 + ld;
  auto uface = [&](int f) {
    if (!g.cyclic && (f == 0 || f == nxt)) return 0.0;
    double u = -ug * (p1[f] - p0[f]);
    if (ekman) u = u + rh * (ty1[f] + ty0[f]);
    if (uek) u = u + uek[(size_t)j * ld + f];
    return u;
  };
  double ulo = 1.0e30, uhi = -1.0e30, vlo = 1.0e30, vhi = -1.0e30, vsq = -1.0e30;
  for (int i = threadIdx.x; i < nxt; i += 256) {
    const double um = uface(i), up = uface(i + 1);
    double vm, vp;
    if (south) vm = vek ? vek[(size_t)j * ld + i] : ((ekman && sflux) ? -rh * (tx0[i + 1] + tx0[i]) : 0.0);
    else {
      vm = ug * (p0[i + 1] - p0[i]);
      if (ekman) vm = vm - rh * (tx0[i + 1] + tx0[i]);
      if (vek) vm = vm + vek[(size_t)j * ld + i];
    }
    if (north) vp = vek ? vek[(size_t)(j + 1) * ld + i] : ((ekman && nflux) ? -rh * (tx1[i + 1] + tx1[i]) : 0.0);
    else {
      vp = ug * (p1[i + 1] - p1[i]);
      if (ekman) vp = vp - rh * (tx1[i + 1] + tx1[i]);
      if (vek) vp = vp + vek[(size_t)(j + 1) * ld + i];
    }
    if (i == 0) { ulo = fmin(ulo, um); uhi = fmax(uhi, um); }
    ulo = fmin(ulo, up); uhi = fmax(uhi, up);
    vlo = fmin(vlo, fmin(vm, vp)); vhi = fmax(vhi, fmax(vm, vp));
    vsq = fmax(vsq, (um + up) * (um + up) + (vm + vp) * (vm + vp));
  }
  row_minmax(ulo, uhi, out + j, out + pitch + j);
  row_minmax(vlo, vhi, out + 2 * pitch + j, out + 3 * pitch + j);
  double dummy;
  row_minmax(0.0, vsq, &dummy, out + 4 * pitch + j);
}

// atmosphere T grid: sums of wekta, |wekta|, ast*hmixa, ast, hmixa and of ast over the columns
// above the ocean (zero outside the ocean's rows); extrema of ast  (:204-212, :424-461)
__global__ void __launch_bounds__(256) k_mon_ta(const double *wekt, const double *ast, const double *hmix, int nxt, int ld, int i0, int i1,
                                                int j0, int j1, double *out, size_t pitch) {
  const int j = blockIdx.x;
  const bool orow = (j >= j0 && j < j1);
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, lo = 1.0e30, hi = -1.0e30;
  for (int i = threadIdx.x; i < nxt; i += 256) {
    const size_t o = (size_t)j * ld + i;
    const double w = wekt[o], t = ast[o], h = hmix[o];
    v[0] += w; v[1] += fabs(w); v[2] += t * h; v[3] += t; v[4] += h;
    if (orow && i >= i0 && i < i1) v[5] += t;
    lo = fmin(lo, t); hi = fmax(hi, t);
  }
  row_store<6>(v, out + j, pitch);
  row_minmax(lo, hi, out + 6 * pitch + j, out + 7 * pitch + j);
}

// Slot layout of the per-row sums (pitch = local nyp):
//   T 6 (wekto, |wekto|, sst*wekto, sst, sst min, sst max)      rows: T
//   P 4 (wekpo, |wekpo|, entoc, |entoc|)                        rows: p
//   eta 4 per interface (eta, eta^2, eta*etadot, eta*entoc)     rows: p
//   pq 4 per layer (po, qo, po min, po max)                     rows: p
//   u 7 per layer (u*del2, u*del4, u^2, u*udot, um^2, u*taux, jet sum)   rows: T
//   v 6 per layer (v*del2, v*del4, v^2, v*vdot, vm^2, v*tauy)           rows: p
//   cour 5 per layer incl. the mixed layer (u min, u max, v min, v max, largest speed^2)   rows: T; extrema only
struct MonLayout {
  int nl, oT, oP, oE, oQ, oU, oV, oC, nsum, nslot;
  explicit MonLayout(int nl_) : nl(nl_) {
    oT = 0; oP = 6; oE = 10; oQ = oE + 4 * (nl - 1); oU = oQ + 4 * nl; oV = oU + 7 * nl; oC = oV + 6 * nl;
    nsum = oC;                       // slots below oC take part in the area sums
    nslot = oC + 5 * (nl + 1);
  }
  bool t_rows(int slot) const { return slot < oP || (slot >= oU && slot < oV); }   // defined on nyt rows
};

// launches the row kernels of one rank and brings the row sums and the two corner values of
// po per layer to the host
static void mon_rows(qgcm_model *m, const MonLayout &L, std::vector<double> &h, double *corner) {
  const Grid &g = m->go;
  const int nl = g.nl, nxp = g.nxp, nyp = g.nyp, nxt = g.nxt, nyt = g.nyt, ld = g.ld;
  const size_t pitch = nyp;
  const size_t need = (size_t)L.nslot * pitch;
  if (m->mon_elems < need) {
    m->d_mon = (double *)dalloc(m, sizeof(double) * need);
    m->mon_elems = need;
  }
  double *rows = m->d_mon;
  const double *po = m->F("po"), *pom = m->F("pom"), *qo = m->F("qo");
  const double rdxf0 = g.rdxf0, dto = m->dto;
  QG_LAUNCH(m, "k_mon_t", nyt, 256, 0, k_mon_t, m->F("wekto"), m->F("sst"), nxt, ld, rows + L.oT * pitch, pitch);
  QG_LAUNCH(m, "k_mon_p", nyp, 256, 0, k_mon_p, m->F("wekpo"), m->F("entoc"), nxp, ld, rows + L.oP * pitch, pitch);
  for (int k = 0; k < nl - 1; ++k) {
    const double rgp = 1.0 / m->lo.gp[k];
    QG_LAUNCH(m, "k_mon_eta", nyp, 256, 0, k_mon_eta, po + (size_t)k * g.lsz, po + (size_t)(k + 1) * g.lsz, pom + (size_t)k * g.lsz,
              pom + (size_t)(k + 1) * g.lsz, m->F("entoc"), rgp, rgp / dto, nxp, ld, rows + (size_t)(L.oE + 4 * k) * pitch, pitch);
  }
  if (!m->d_monf) m->d_monf = (double *)dalloc(m, sizeof(double) * 4 * g.lsz);
  double *ugm = m->d_monf, *vgm = ugm + g.lsz, *d2 = vgm + g.lsz, *d4 = d2 + g.lsz;
  const dim3 full((nxp + 255) / 256, nyp);
  for (int k = 0; k < nl; ++k) {
    const double *pk = po + (size_t)k * g.lsz, *pmk = pom + (size_t)k * g.lsz;
    QG_LAUNCH(m, "k_mon_pq", nyp, 256, 0, k_mon_pq, pk, qo + (size_t)k * g.lsz, nxp, ld, rows + (size_t)(L.oQ + 4 * k) * pitch, pitch);
    QG_LAUNCH(m, "k_mon_geo", full, 256, 0, k_mon_geo, pmk, ugm, vgm, g, rdxf0);
    // at the inner edge of a y-slab the one-sided formulas spoil del2 on the edge row and del4
    // on two rows: halo rows (three per edge), never owned ones
    QG_LAUNCH(m, "k_mon_lap", dim3((nxp + 255) / 256, nyt), 256, 0, k_mon_lap, ugm, d2, nxp, nyt, ld, g.cyclic, g.dxm2);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxp + 255) / 256, nyt), 256, 0, k_mon_lap, d2, d4, nxp, nyt, ld, g.cyclic, g.dxm2);
    QG_LAUNCH(m, "k_mon_u", nyt, 256, 0, k_mon_u, pk, pmk, ugm, d2, d4, m->F("tauxo"), rdxf0, rdxf0 / dto, nxp, ld, 0,
              rows + (size_t)(L.oU + 7 * k) * pitch, pitch);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxt + 255) / 256, nyp), 256, 0, k_mon_lap, vgm, d2, nxt, nyp, ld, g.cyclic, g.dxm2);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxt + 255) / 256, nyp), 256, 0, k_mon_lap, d2, d4, nxt, nyp, ld, g.cyclic, g.dxm2);
    QG_LAUNCH(m, "k_mon_v", nyp, 256, 0, k_mon_v, pk, pmk, vgm, d2, d4, m->F("tauyo"), rdxf0, rdxf0 / dto, nxt, ld, 0,
              rows + (size_t)(L.oV + 6 * k) * pitch, pitch);
  }
  {
    const double rh = 0.5 / (m->fnot * m->cfg.hmoc);
    QG_LAUNCH(m, "k_mon_cour", nyt, 256, 0, k_mon_cour, po, m->F("tauxo"), m->F("tauyo"), (const double *)nullptr, (const double *)nullptr, g, 1,
              (int)m->sb_hflux, (int)m->nb_hflux,
              m->cfg.ycexp * rdxf0, rh, rows + (size_t)L.oC * pitch, pitch);
    for (int k = 0; k < nl; ++k)
      QG_LAUNCH(m, "k_mon_cour", nyt, 256, 0, k_mon_cour, po + (size_t)k * g.lsz, m->F("tauxo"), m->F("tauyo"), (const double *)nullptr,
                (const double *)nullptr, g, 0, 0, 0, rdxf0, 0.0,
                rows + (size_t)(L.oC + 5 * (k + 1)) * pitch, pitch);
  }
  h.resize(need);
  QG_CUDA(cudaMemcpyAsync(h.data(), rows, sizeof(double) * need, cudaMemcpyDeviceToHost, m->stream));
  for (int k = 0; k < nl; ++k) {
    corner[2 * k] = corner[2 * k + 1] = 0.0;
    if (g.wall_s())
      QG_CUDA(cudaMemcpyAsync(&corner[2 * k], po + (size_t)k * g.lsz, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    if (g.wall_n())
      QG_CUDA(cudaMemcpyAsync(&corner[2 * k + 1], po + (size_t)k * g.lsz + (size_t)(nyp - 1) * ld, sizeof(double), cudaMemcpyDeviceToHost,
                              m->stream));
  }
  QG_CUDA(cudaStreamSynchronize(m->stream));
}

// What a rank contributes: per sum slot the sum over the interior rows it owns (genint's inner
// rows, :1180-1190) and the values of the global southern / northern row if it owns them; its
// extrema, jet candidates and corner values in a one-hot block per rank.  Every entry combines
// across ranks by addition.
struct MonShare {
  std::vector<double> v;
  int nsum, ngat;           // 3 per sum slot + 2*nl corners | per rank: sst min/max, po min/max, jet value/row per layer
};
static void mon_share(qgcm_model *m, const MonLayout &L, const std::vector<double> &h, const double *corner, MonShare &S) {
  const Grid &g = m->go;
  const int nl = g.nl, nranks = m->nranks, rank = m->rank;
  const size_t pitch = g.nyp;
  S.nsum = 3 * L.nsum + 2 * nl;
  S.ngat = 2 + 4 * nl + 5 * (nl + 1);
  S.v.assign((size_t)S.nsum + (size_t)nranks * S.ngat, 0.0);
  double *gat = S.v.data() + S.nsum + (size_t)rank * S.ngat;
  auto owned = [&](bool trows, int &j0, int &j1, int &nyg) {
    nyg = trows ? g.nyp_g - 1 : g.nyp_g;
    j0 = g.own0;
    j1 = std::min(g.own1, (trows ? g.nyt : g.nyp));
  };
  for (int slot = 0; slot < L.nsum; ++slot) {
    int j0, j1, nyg;
    owned(L.t_rows(slot), j0, j1, nyg);
    const double *row = h.data() + (size_t)slot * pitch;
    double in = 0.0;
    for (int j = j0; j < j1; ++j) {
      const int jg = g.jg0 + j;
      if (jg == 0) S.v[3 * slot + 1] = row[j];
      else if (jg == nyg - 1) S.v[3 * slot + 2] = row[j];
      else in = in + row[j];
    }
    S.v[3 * slot] = in;
  }
  for (int k = 0; k < nl; ++k) {
    S.v[3 * L.nsum + 2 * k] = corner[2 * k];
    S.v[3 * L.nsum + 2 * k + 1] = corner[2 * k + 1];
  }
  // extrema and jet candidates over owned rows
  int j0, j1, nyg;
  owned(true, j0, j1, nyg);
  double lo = 1.0e30, hi = -1.0e30;
  for (int j = j0; j < j1; ++j) {
    lo = std::min(lo, h[(size_t)(L.oT + 4) * pitch + j]);
    hi = std::max(hi, h[(size_t)(L.oT + 5) * pitch + j]);
  }
  gat[0] = lo; gat[1] = hi;
  for (int k = 0; k < nl; ++k) {
    int p0, p1, nypg;
    owned(false, p0, p1, nypg);
    lo = 1.0e30; hi = -1.0e30;
    for (int j = p0; j < p1; ++j) {
      lo = std::min(lo, h[(size_t)(L.oQ + 4 * k + 2) * pitch + j]);
      hi = std::max(hi, h[(size_t)(L.oQ + 4 * k + 3) * pitch + j]);
    }
    gat[2 + 4 * k] = lo; gat[2 + 4 * k + 1] = hi;
    // largest |zonal mean u| among the owned T rows: first strict maximum, as :696-704
    double best = 0.0; int brow = 0;
    const double *uj = h.data() + (size_t)(L.oU + 7 * k + 6) * pitch;
    for (int j = j0; j < j1; ++j) {
      const double val = std::fabs(uj[j]) / (double)g.nxt;
      if (val > best) { best = val; brow = g.jg0 + j + 1; }
    }
    gat[2 + 4 * k + 2] = best; gat[2 + 4 * k + 3] = (double)brow;
  }
  // couroc: extrema over the owned T rows, mixed layer first
  for (int k = 0; k <= nl; ++k) {
    double e[5] = {1.0e30, -1.0e30, 1.0e30, -1.0e30, -1.0e30};
    const double *cr = h.data() + (size_t)(L.oC + 5 * k) * pitch;
    for (int j = j0; j < j1; ++j) {
      e[0] = std::min(e[0], cr[j]);
      e[1] = std::max(e[1], cr[pitch + j]);
      e[2] = std::min(e[2], cr[2 * pitch + j]);
      e[3] = std::max(e[3], cr[3 * pitch + j]);
      e[4] = std::max(e[4], cr[4 * pitch + j]);
    }
    for (int q = 0; q < 5; ++q) gat[2 + 4 * nl + 5 * k + q] = e[q];
  }
}

// the reference's scalar algebra on the combined sums (src/monitor_diag.F:498-821)
static void mon_finish(qgcm_model *m, const MonLayout &L, const MonShare &S, qgcm_monitor_ocean *r) {
  std::memset(r, 0, sizeof(*r));
  const Grid &g = m->go;
  const qgcm_config &c = m->cfg;
  const int nl = g.nl, nranks = m->nranks;
  const double ocnorm = g.norm, rhooc = c.rhooc, fnot = m->fnot;
  auto gi = [&](int slot, double facsn) { return S.v[3 * slot] + facsn * (S.v[3 * slot + 1] + S.v[3 * slot + 2]); };   // genint :1205
  auto gat = [&](int rk, int i) { return S.v[(size_t)S.nsum + (size_t)rk * S.ngat + i]; };
  r->wetmoc = gi(L.oT + 0, 1.0) * ocnorm;
  r->watmoc = gi(L.oT + 1, 1.0) * ocnorm;
  r->wepmoc = gi(L.oP + 0, 0.5) * ocnorm;
  r->wapmoc = gi(L.oP + 1, 0.5) * ocnorm;
  r->entmoc = gi(L.oP + 2, 0.5) * ocnorm;
  r->enamoc = gi(L.oP + 3, 0.5) * ocnorm;
  for (int k = 0; k < nl - 1; ++k) {
    r->etamoc[k] = gi(L.oE + 4 * k, 0.5) * ocnorm;
    r->et2moc[k] = gi(L.oE + 4 * k + 1, 0.5) * ocnorm;
    r->ddtpeoc[k] = rhooc * m->lo.gp[k] * gi(L.oE + 4 * k + 2, 0.5);
    if (k == 0) r->pkenoc = rhooc * m->lo.gp[0] * gi(L.oE + 3, 0.5) * ocnorm;
  }
  {
    const double utaux = gi(L.oU + 5, 1.0), vtauy = gi(L.oV + 5, 0.5);
    r->utauoc = rhooc * (vtauy + utaux) * ocnorm;
  }
  for (int k = 0; k < nl; ++k) {
    const double hk = m->lo.h[k];
    const int u = L.oU + 7 * k, v = L.oV + 6 * k;
    const double u2diss = gi(u, 1.0), u4diss = gi(u + 1, 1.0), uke = gi(u + 2, 1.0), ukedot = gi(u + 3, 1.0);
    const double v2diss = gi(v, 0.5), v4diss = gi(v + 1, 0.5), vke = gi(v + 2, 0.5), vkedot = gi(v + 3, 0.5);
    r->pavgoc[k] = gi(L.oQ + 4 * k, 0.5) * ocnorm;
    r->qavgoc[k] = gi(L.oQ + 4 * k + 1, 0.5) * ocnorm;
    r->ah2doc[k] = -rhooc * m->lo.ah2[k] * hk * (u2diss + v2diss) * ocnorm;
    r->ah4doc[k] = rhooc * m->lo.ah4[k] * hk * (u4diss + v4diss) * ocnorm;
    r->kealoc[k] = 0.5 * rhooc * hk * (uke + vke) * ocnorm;
    r->ddtkeoc[k] = rhooc * hk * (ukedot + vkedot) * ocnorm;
    double pomin = 1.0e30, pomax = -1.0e30;
    r->ocjpos[k] = 0;
    r->ocjval[k] = 0.0;
    for (int rk = 0; rk < nranks; ++rk) {       // ranks hold increasing rows: the first strict maximum wins
      pomin = std::min(pomin, gat(rk, 2 + 4 * k));
      pomax = std::max(pomax, gat(rk, 2 + 4 * k + 1));
      if (gat(rk, 2 + 4 * k + 2) > r->ocjval[k]) {
        r->ocjval[k] = gat(rk, 2 + 4 * k + 2);
        r->ocjpos[k] = (int32_t)gat(rk, 2 + 4 * k + 3);
      }
    }
    const double ps = S.v[3 * L.nsum + 2 * k], pn = S.v[3 * L.nsum + 2 * k + 1];   // po(1,1,k), po(1,nypo,k)
    const double poref = (fnot > 0.0) ? ps : pn;
    double psiext = std::min(pomin / fnot, pomax / fnot);
    r->osfmin[k] = 1.0e-6 * hk * (psiext - poref / fnot);
    psiext = std::max(pomin / fnot, pomax / fnot);
    r->osfmax[k] = 1.0e-6 * hk * (psiext - poref / fnot);
    r->occirc[k] = 1.0e-6 * hk * (ps - pn) / fnot;
  }
  {
    const double u2 = gi(L.oU + 7 * (nl - 1) + 4, 1.0), v2 = gi(L.oV + 6 * (nl - 1) + 4, 0.5);
    r->btdgoc = 0.5 * rhooc * c.delek * std::fabs(fnot) * (u2 + v2) * ocnorm;
  }
  r->sstmin = 1.0e30;
  r->sstmax = -1.0e30;
  for (int rk = 0; rk < nranks; ++rk) {
    r->sstmin = std::min(r->sstmin, gat(rk, 0));
    r->sstmax = std::max(r->sstmax, gat(rk, 1));
  }
  r->hfmloc = rhooc * c.cpoc * gi(L.oT + 2, 1.0) * ocnorm;
  r->tmlmoc = gi(L.oT + 3, 1.0) * ocnorm;
  r->occtot = 0.0;
  for (int k = 0; k < nl; ++k) r->occtot = r->occtot + r->occirc[k];
  // couroc (:1711-1715, :1915-1919)
  const double cfac = g.hdxm1 * m->dto;
  for (int k = 0; k <= nl; ++k) {
    double e[5] = {1.0e30, -1.0e30, 1.0e30, -1.0e30, -1.0e30};
    for (int rk = 0; rk < nranks; ++rk) {
      const int b = 2 + 4 * nl + 5 * k;
      e[0] = std::min(e[0], gat(rk, b)); e[1] = std::max(e[1], gat(rk, b + 1));
      e[2] = std::min(e[2], gat(rk, b + 2)); e[3] = std::max(e[3], gat(rk, b + 3));
      e[4] = std::max(e[4], gat(rk, b + 4));
    }
    if (k == 0) {
      r->umminoc = e[0]; r->ummaxoc = e[1]; r->vmminoc = e[2]; r->vmmaxoc = e[3];
      r->cnmloc = cfac * std::sqrt(e[4]);
    } else {
      r->ugminoc[k - 1] = e[0]; r->ugmaxoc[k - 1] = e[1]; r->vgminoc[k - 1] = e[2]; r->vgmaxoc[k - 1] = e[3];
      r->cnqgoc[k - 1] = cfac * std::sqrt(e[4]);
    }
  }
}

// On a y-slab partition every rank sums the rows it owns and the shares are added across the
// ranks (in chunks of the all-reduce payload; this runs once per model day).
void launch_monnc_ocean(qgcm_model *m, qgcm_monitor_ocean *rep) {
  std::memset(rep, 0, sizeof(*rep));
  if (!m->has_ocean) return;
  const Ranks ms = ranks_of(m);
  const MonLayout L(m->go.nl);
  std::vector<MonShare> S(ms.size());
  for (size_t r = 0; r < ms.size(); ++r) {
    std::vector<double> h;
    double corner[2 * NLMAX];
    mon_rows(ms[r], L, h, corner);
    mon_share(ms[r], L, h, corner, S[r]);
  }
  if (m->nranks > 1) {
    const size_t n = S[0].v.size();
    for (size_t off = 0; off < n; off += PEER_VEC) {
      const size_t len = std::min((size_t)PEER_VEC, n - off);
      std::vector<std::vector<double>> part(ms.size());
      for (size_t r = 0; r < ms.size(); ++r) part[r].assign(S[r].v.begin() + off, S[r].v.begin() + off + len);
      comm_allreduce_host(ms, part);
      for (size_t r = 0; r < ms.size(); ++r) std::copy(part[r].begin(), part[r].end(), S[r].v.begin() + off);
    }
  }
  for (size_t r = 0; r < ms.size(); ++r)
    if (ms[r] == m) mon_finish(m, L, S[r], rep);
}

// genint's outer sum over the row sums of a whole grid (:1180-1207)
static double rows_int(const double *row, int ny, double facsn) {
  double answer = 0.0;
  for (int j = 1; j < ny - 1; ++j) answer = answer + row[j];
  return answer + facsn * (row[0] + row[ny - 1]);
}

// ---- atmosphere section (src/monitor_diag.F:186-478) and courat (:1215-1445); one GPU ----
void launch_monnc_atmos(qgcm_model *m, qgcm_monitor_atmos *r) {
  std::memset(r, 0, sizeof(*r));
  if (!m->has_atmos) return;
  const Grid &g = m->ga;
  const qgcm_config &c = m->cfg;
  const int nl = g.nl, nxp = g.nxp, nyp = g.nyp, nxt = g.nxt, nyt = g.nyt, ld = g.ld;
  const size_t pitch = nyp;
  // slots: T 8 | P 4 | dtopat 4 | eta 4 per interface | pq 4 per layer | u 7 per layer | v 6 per layer | cour 5 per layer + mixed layer
  const int oT = 0, oP = 8, oD = 12, oE = 16, oQ = oE + 4 * (nl - 1), oU = oQ + 4 * nl, oV = oU + 7 * nl, oC = oV + 6 * nl;
  const int nslot = oC + 5 * (nl + 1);
  const size_t need = (size_t)nslot * pitch;
  if (m->mona_elems < need) {
    m->d_mona = (double *)dalloc(m, sizeof(double) * need);
    m->mona_elems = need;
  }
  if (!m->d_monaf) m->d_monaf = (double *)dalloc(m, sizeof(double) * 4 * g.lsz);
  double *rows = m->d_mona;
  double *ugm = m->d_monaf, *vgm = ugm + g.lsz, *d2 = vgm + g.lsz, *d4 = d2 + g.lsz;
  const double *pa = m->F("pa"), *pam = m->F("pam"), *qa = m->F("qa");
  const double rdxf0 = g.rdxf0, dta = m->dta;
  const int nxaooc = m->go.nxt / c.ndxr, nyaooc = (m->go.nyp_g - 1) / c.ndxr;
  QG_LAUNCH(m, "k_mon_ta", nyt, 256, 0, k_mon_ta, m->F("wekta"), m->F("ast"), m->F("hmixa"), nxt, ld, c.nx1 - 1, c.nx1 - 1 + nxaooc,
            c.ny1 - 1, c.ny1 - 1 + nyaooc, rows + (size_t)oT * pitch, pitch);
  QG_LAUNCH(m, "k_mon_p", nyp, 256, 0, k_mon_p, m->F("wekpa"), m->F("entat"), nxp, ld, rows + (size_t)oP * pitch, pitch);
  QG_LAUNCH(m, "k_mon_p", nyp, 256, 0, k_mon_p, m->F("dtopat"), m->F("dtopat"), nxp, ld, rows + (size_t)oD * pitch, pitch);
  for (int k = 0; k < nl - 1; ++k) {
    const double rgp = 1.0 / m->la.gp[k];
    // eta = rgp (pa_k - pa_k+1): the ocean kernel's (p_k+1 - p_k) with the sign carried by rgp (:259)
    QG_LAUNCH(m, "k_mon_eta", nyp, 256, 0, k_mon_eta, pa + (size_t)k * g.lsz, pa + (size_t)(k + 1) * g.lsz, pam + (size_t)k * g.lsz,
              pam + (size_t)(k + 1) * g.lsz, m->F("entat"), -rgp, rgp / dta, nxp, ld, rows + (size_t)(oE + 4 * k) * pitch, pitch);
  }
  const dim3 full((nxp + 255) / 256, nyp);
  for (int k = 0; k < nl; ++k) {
    const double *pk = pa + (size_t)k * g.lsz, *pmk = pam + (size_t)k * g.lsz;
    QG_LAUNCH(m, "k_mon_pq", nyp, 256, 0, k_mon_pq, pk, qa + (size_t)k * g.lsz, nxp, ld, rows + (size_t)(oQ + 4 * k) * pitch, pitch);
    QG_LAUNCH(m, "k_mon_geo", full, 256, 0, k_mon_geo, pmk, ugm, vgm, g, rdxf0);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxp + 255) / 256, nyt), 256, 0, k_mon_lap, ugm, d2, nxp, nyt, ld, 1, g.dxm2);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxp + 255) / 256, nyt), 256, 0, k_mon_lap, d2, d4, nxp, nyt, ld, 1, g.dxm2);
    QG_LAUNCH(m, "k_mon_u", nyt, 256, 0, k_mon_u, pk, pmk, ugm, d2, d4, m->F("tauxa"), rdxf0, rdxf0 / dta, nxp, ld, 1,
              rows + (size_t)(oU + 7 * k) * pitch, pitch);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxt + 255) / 256, nyp), 256, 0, k_mon_lap, vgm, d2, nxt, nyp, ld, 1, g.dxm2);
    QG_LAUNCH(m, "k_mon_lap", dim3((nxt + 255) / 256, nyp), 256, 0, k_mon_lap, d2, d4, nxt, nyp, ld, 1, g.dxm2);
    QG_LAUNCH(m, "k_mon_v", nyp, 256, 0, k_mon_v, pk, pmk, vgm, d2, d4, m->F("tauya"), rdxf0, rdxf0 / dta, nxt, ld, 1,
              rows + (size_t)(oV + 6 * k) * pitch, pitch);
  }
  QG_LAUNCH(m, "k_mon_cour", nyt, 256, 0, k_mon_cour, pa, m->F("tauxa"), m->F("tauya"), (const double *)m->F("uekat"),
            (const double *)m->F("vekat"), g, 0, 0, 0, rdxf0, 0.0, rows + (size_t)oC * pitch, pitch);
  for (int k = 0; k < nl; ++k)
    QG_LAUNCH(m, "k_mon_cour", nyt, 256, 0, k_mon_cour, pa + (size_t)k * g.lsz, m->F("tauxa"), m->F("tauya"), (const double *)nullptr,
              (const double *)nullptr, g, 0, 0, 0, rdxf0, 0.0, rows + (size_t)(oC + 5 * (k + 1)) * pitch, pitch);
  std::vector<double> h(need);
  QG_CUDA(cudaMemcpyAsync(h.data(), rows, sizeof(double) * need, cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
  auto gi = [&](int slot, int ny, double facsn) { return rows_int(h.data() + (size_t)slot * pitch, ny, facsn); };
  const double atnorm = g.norm, rhoat = c.rhoat;
  r->wetmat = gi(oT + 0, nyt, 1.0) * atnorm;
  r->watmat = gi(oT + 1, nyt, 1.0) * atnorm;
  r->wepmat = gi(oP + 0, nyp, 0.5) * atnorm;
  r->wapmat = gi(oP + 1, nyp, 0.5) * atnorm;
  r->entmat[0] = gi(oP + 2, nyp, 0.5) * atnorm;
  r->enamat[0] = gi(oP + 3, nyp, 0.5) * atnorm;
  for (int k = 0; k < nl - 1; ++k) {
    r->etamat[k] = gi(oE + 4 * k, nyp, 0.5) * atnorm;
    r->et2mat[k] = gi(oE + 4 * k + 1, nyp, 0.5) * atnorm;
    r->ddtpeat[k] = rhoat * m->la.gp[k] * gi(oE + 4 * k + 2, nyp, 0.5);
    r->pkenat[k] = (k == 0) ? rhoat * m->la.gp[0] * gi(oE + 3, nyp, 0.5) * atnorm : 0.0;
  }
  {
    const double utaux = gi(oU + 5, nyt, 1.0), vtauy = gi(oV + 5, nyp, 0.5);
    r->utauat = rhoat * (vtauy + utaux) * atnorm;
  }
  for (int k = 0; k < nl; ++k) {
    const double hk = m->la.h[k];
    const int u = oU + 7 * k, v = oV + 6 * k;
    const double u4diss = gi(u + 1, nyt, 1.0), uke = gi(u + 2, nyt, 1.0), ukedot = gi(u + 3, nyt, 1.0);
    const double v4diss = gi(v + 1, nyp, 0.5), vke = gi(v + 2, nyp, 0.5), vkedot = gi(v + 3, nyp, 0.5);
    r->pavgat[k] = gi(oQ + 4 * k, nyp, 0.5) * atnorm;
    r->qavgat[k] = gi(oQ + 4 * k + 1, nyp, 0.5) * atnorm;
    r->ah4dat[k] = rhoat * m->la.ah4[k] * hk * (u4diss + v4diss) * atnorm;
    r->kealat[k] = 0.5 * rhoat * hk * (uke + vke) * atnorm;
    r->ddtkeat[k] = rhoat * hk * (ukedot + vkedot) * atnorm;
    const double *uj = h.data() + (size_t)(u + 6) * pitch;
    r->atstpos[k] = 0;
    r->atstval[k] = 0.0;
    for (int j = 0; j < nyt; ++j) {
      const double val = std::fabs(uj[j]) / (double)nxt;
      if (val > r->atstval[k]) { r->atstpos[k] = j + 1; r->atstval[k] = val; }
    }
  }
  r->tmlmat = gi(oT + 3, nyt, 1.0) * atnorm;
  r->hmlmat = gi(oT + 4, nyt, 1.0) * atnorm;
  r->hcmlat = rhoat * c.cpat * gi(oT + 2, nyt, 1.0) * atnorm;
  r->astmin = 1.0e30;
  r->astmax = -1.0e30;
  double tmaooc = 0.0;
  for (int j = 0; j < nyt; ++j) {
    r->astmin = std::min(r->astmin, h[(size_t)(oT + 6) * pitch + j]);
    r->astmax = std::max(r->astmax, h[(size_t)(oT + 7) * pitch + j]);
    tmaooc = tmaooc + h[(size_t)(oT + 5) * pitch + j];
  }
  r->tmaooc = tmaooc / (double)(nxaooc * nyaooc);
  // xintp of dtopat (src/intsubs.f:78-133) has genint's weights 0.5, 0.5
  r->davgat = gi(oD, nyp, 0.5) * atnorm;
  double olrtop = c.Bup[nl - 1] * (r->hmlmat - c.hmat) + c.Cup[nl - 1] * r->davgat + c.Dup[nl - 1] * r->tmlmat;
  for (int i = 0; i < nl - 1; ++i) olrtop = olrtop + c.Aup[(nl - 1) + nl * i] * r->etamat[i];
  r->olrtop = olrtop;
  const double cfac = g.hdxm1 * dta;
  for (int k = 0; k <= nl; ++k) {
    double e[5] = {1.0e30, -1.0e30, 1.0e30, -1.0e30, -1.0e30};
    const double *cr = h.data() + (size_t)(oC + 5 * k) * pitch;
    for (int j = 0; j < nyt; ++j) {
      e[0] = std::min(e[0], cr[j]); e[1] = std::max(e[1], cr[pitch + j]);
      e[2] = std::min(e[2], cr[2 * pitch + j]); e[3] = std::max(e[3], cr[3 * pitch + j]);
      e[4] = std::max(e[4], cr[4 * pitch + j]);
    }
    if (k == 0) {
      r->umminat = e[0]; r->ummaxat = e[1]; r->vmminat = e[2]; r->vmmaxat = e[3];
      r->cnmlat = cfac * std::sqrt(e[4]);
    } else {
      r->ugminat[k - 1] = e[0]; r->ugmaxat[k - 1] = e[1]; r->vgminat[k - 1] = e[2]; r->vgmaxat[k - 1] = e[3];
      r->cnqgat[k - 1] = cfac * std::sqrt(e[4]);
    }
  }
}

}  // namespace qg

// Perimeter, pointwise and reduction kernels: ocqbdy / atqzbd (src/vorsubs.F:245-480),
// qcomp / merqcy (src/vorsubs.F:49-236), the oceanic Ekman tail of xforc
// (src/xfosubs.F:568-683), time-level averaging (src/q-gcm.F:1328-1407) and constr
// (src/conhoms.F:44-314).
#include "qgcm_internal.h"

namespace qg {

struct BdyArgs {
  Grid g;
  int atmos, nl;
  double f0, beta, bcfac;   // bcfac = bcco*dxm2/(0.5 bcco+1)/f0
  double amat[NLMAX * NLMAX];
  const double *p, *ddyn, *yrel;
  double *q;
};

// value of q on a solid boundary point from the mixed condition; (di,dj) points inward
__device__ __forceinline__ double qbdy_point(const BdyArgs &a, int k, int i, int j, int di, int dj) {
  const Grid &g = a.g;
  const int nl = a.nl;
  const size_t idx = (size_t)j * g.ld + i, inw = (size_t)(j + dj) * g.ld + (i + di);
  const double *p = a.p;
  const double pk = p[k * g.lsz + idx];
  const double betay = a.beta * a.yrel[j];
  double coupl;
  if (k == 0) {
    coupl = a.f0 * a.amat[0] * pk + a.f0 * a.amat[0 + nl * 1] * p[1 * g.lsz + idx];
  } else if (k == nl - 1) {
    double pc = pk;
    // reference quirk, src/vorsubs.F:470: southern row of the top atmospheric layer
    // multiplies f0*A(nla,nla) by pa(i,2,nla), not pa(i,1,nla)
    if (a.atmos && j == 0) pc = p[k * g.lsz + inw];
    coupl = a.f0 * a.amat[k + nl * (k - 1)] * p[(k - 1) * g.lsz + idx] + a.f0 * a.amat[k + nl * k] * pc;
  } else {
    coupl = a.f0 * a.amat[k + nl * (k - 1)] * p[(k - 1) * g.lsz + idx] + a.f0 * a.amat[k + nl * k] * pk +
            a.f0 * a.amat[k + nl * (k + 1)] * p[(k + 1) * g.lsz + idx];
  }
  double v = a.bcfac * (p[k * g.lsz + inw] - pk) - coupl + betay;
  const int kbot = a.atmos ? 0 : nl - 1;
  if (k == kbot) v = v + a.ddyn[idx];
  return v;
}

// grid (ceil(max(nxp,nyp)/256), 4 sides, nl)
__global__ void __launch_bounds__(256) k_qbdy(BdyArgs a) {
  const Grid &g = a.g;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int side = blockIdx.y, k = blockIdx.z;
  if (side == 0) {          // south row, all i
    if (t < g.nxp) a.q[k * g.lsz + t] = qbdy_point(a, k, t, 0, 0, 1);
  } else if (side == 1) {   // north row
    if (t < g.nxp) a.q[k * g.lsz + (size_t)(g.nyp - 1) * g.ld + t] = qbdy_point(a, k, t, g.nyp - 1, 0, -1);
  } else if (!g.cyclic && !a.atmos) {
    const int j = t + 1;    // meridional walls, j = 2..nyp-1
    if (j < g.nyp - 1) {
      if (side == 2) a.q[k * g.lsz + (size_t)j * g.ld] = qbdy_point(a, k, 0, j, 1, 0);
      else a.q[k * g.lsz + (size_t)j * g.ld + g.nxp - 1] = qbdy_point(a, k, g.nxp - 1, j, -1, 0);
    }
  }
}

static void fill_bdy(qgcm_model *m, bool atmos, BdyArgs &a, double *q, const double *p) {
  const Grid &g = atmos ? m->ga : m->go;
  const LayerConsts &lc = atmos ? m->la : m->lo;
  const double bcco = atmos ? m->cfg.bccoat : m->cfg.bccooc;
  a.g = g; a.atmos = atmos; a.nl = g.nl; a.f0 = m->fnot; a.beta = m->beta;
  a.bcfac = bcco * g.dxm2 / (0.5 * bcco + 1.0) / m->fnot;
  for (int i = 0; i < NLMAX * NLMAX; ++i) a.amat[i] = lc.amat[i];
  a.p = p; a.q = q;
  a.ddyn = m->F(atmos ? "ddynat" : "ddynoc");
  a.yrel = atmos ? m->yparel : m->yporel;
}

void launch_ocqbdy(qgcm_model *m, double *q, const double *p) {
  BdyArgs a;
  fill_bdy(m, false, a, q, p);
  const int n = max(a.g.nxp, a.g.nyp);
  QG_LAUNCH(m, "k_qbdy", dim3((n + 255) / 256, a.g.cyclic ? 2 : 4, a.g.nl), 256, 0, k_qbdy, a);
  QG_CUDA(cudaGetLastError());
}

void launch_atqzbd(qgcm_model *m, double *q, const double *p) {
  BdyArgs a;
  fill_bdy(m, true, a, q, p);
  QG_LAUNCH(m, "k_qbdy", dim3((a.g.nxp + 255) / 256, 2, a.g.nl), 256, 0, k_qbdy, a);
  QG_CUDA(cudaGetLastError());
}

// q from p at interior points (qcomp) and, for periodic grids, the W/E columns (merqcy)
__global__ void __launch_bounds__(256) k_qcomp(BdyArgs a) {
  const Grid &g = a.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y + 1, k = blockIdx.z;
  if (i >= g.nxp) return;
  const bool edge = (i == 0 || i == g.nxp - 1);
  if (edge && !g.cyclic) return;
  const int nl = a.nl, per = g.nxp - 1;
  const int ic = edge ? 0 : i;
  const int iw = (ic == 0) ? per - 1 : ic - 1, ie = ic + 1;
  const double *p = a.p + k * g.lsz;
  const size_t r = (size_t)j * g.ld;
  const double dx2fac = g.dxm2 / a.f0;
  const double betay = a.beta * a.yrel[j];
  const double lap = p[r - g.ld + ic] + p[r + iw] + p[r + ie] + p[r + g.ld + ic] - 4.0 * p[r + ic];
  double coupl;
  if (k == 0) coupl = a.amat[0] * p[r + ic] + a.amat[nl] * a.p[g.lsz + r + ic];
  else if (k == nl - 1) coupl = a.amat[k + nl * (k - 1)] * a.p[(k - 1) * g.lsz + r + ic] + a.amat[k + nl * k] * p[r + ic];
  else coupl = a.amat[k + nl * (k - 1)] * a.p[(k - 1) * g.lsz + r + ic] + a.amat[k + nl * k] * p[r + ic] +
               a.amat[k + nl * (k + 1)] * a.p[(k + 1) * g.lsz + r + ic];
  double v = dx2fac * lap + betay - a.f0 * coupl;
  const int kbot = a.atmos ? 0 : nl - 1;
  if (k == kbot) v = v + a.ddyn[r + ic];
  a.q[k * g.lsz + r + i] = v;
}

void launch_qcomp(qgcm_model *m, bool ocean, double *q, const double *p) {
  BdyArgs a;
  fill_bdy(m, !ocean, a, q, p);
  QG_LAUNCH(m, "k_qcomp", dim3((a.g.nxp + 255) / 256, a.g.nyp - 2, a.g.nl), 256, 0, k_qcomp, a);
  QG_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------
// oceanic Ekman velocities from the stress (src/xfosubs.F:568-683)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wekto(Grid g, double hxofac, const double *tx, const double *ty, double *wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= g.nxt) return;
  const size_t r = (size_t)j * g.ld + i, rn = r + g.ld;
  wt[r] = hxofac * (ty[rn + 1] + ty[r + 1] - (ty[rn] + ty[r]) + tx[r + 1] + tx[r] - (tx[rn + 1] + tx[rn]));
}
__global__ void __launch_bounds__(256) k_wekpo(Grid g, const double *wt, double *wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= g.nxp) return;
  const int nxt = g.nxt, nyt = g.nyt, ld = g.ld, cyc = g.cyclic;
  const bool colW = i == 0, colE = i == g.nxp - 1, rowS = j == 0, rowN = j == g.nyp - 1;
  int im = i - 1, ic = i;
  if (cyc && (colW || colE)) { im = nxt - 1; ic = 0; }
#define W(ii, jj) wt[(size_t)(jj) * ld + (ii)]
  double v;
  if (!rowS && !rowN) {
    if (!cyc && colW) v = 0.5 * (W(0, j - 1) + W(0, j));
    else if (!cyc && colE) v = 0.5 * (W(nxt - 1, j - 1) + W(nxt - 1, j));
    else v = 0.25 * (W(im, j - 1) + W(im, j) + W(ic, j - 1) + W(ic, j));
  } else {
    const int jt = rowS ? 0 : nyt - 1;
    if (!cyc && colW) v = W(0, jt);
    else if (!cyc && colE) v = W(nxt - 1, jt);
    else v = 0.5 * (W(im, jt) + W(ic, jt));
  }
#undef W
  wp[(size_t)j * ld + i] = v;
}
// channel: txisoc, txinoc (src/xfosubs.F:672-683)
__global__ void __launch_bounds__(256) k_txis(Grid g, double dx, const double *tx, double *out_s, double *out_n) {
  __shared__ double red[2][8];
  double s = 0.0, n = 0.0;
  for (int i = threadIdx.x; i < g.nxp; i += 256) {
    const double w = (i == 0 || i == g.nxp - 1) ? 0.5 : 1.0;
    s += w * (tx[i] + tx[g.ld + i]);
    n += w * (tx[(size_t)(g.nyp - 2) * g.ld + i] + tx[(size_t)(g.nyp - 1) * g.ld + i]);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); n += __shfl_down_sync(0xffffffffu, n, o); }
  if (lane == 0) { red[0][w] = s; red[1][w] = n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
    *out_s = 0.5 * dx * a;
    *out_n = 0.5 * dx * b;
  }
}

void launch_xforc_ocean_ekman(qgcm_model *m) {
  const Grid &g = m->go;
  const double hxofac = 0.5 * g.rdxf0;
  QG_LAUNCH(m, "k_wekto", dim3((g.nxt + 255) / 256, g.nyt), 256, 0, k_wekto, g, hxofac, m->F("tauxo"), m->F("tauyo"), m->F("wekto"));
  QG_LAUNCH(m, "k_wekpo", dim3((g.nxp + 255) / 256, g.nyp), 256, 0, k_wekpo, g, m->F("wekto"), m->F("wekpo"));
  if (g.cyclic) {
    QG_LAUNCH(m, "k_txis", 1, 256, 0, k_txis, g, g.dx, m->F("tauxo"), &m->d_scal->txisoc, &m->d_scal->txinoc);
  }
  QG_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------
// time-level averaging (src/q-gcm.F:1328-1407)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_avg2(double *x, const double *xm, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = 0.5 * (x[i] + xm[i]);
}
__global__ void k_avg_scalars(qgcm_scalars *s, int atmos, int cyclic, int nl) {
  if (threadIdx.x != 0) return;
  if (!atmos) {
    for (int k = 0; k < nl - 1; ++k) s->dpioc[k] = 0.5 * (s->dpioc[k] + s->dpiocp[k]);
    if (cyclic)
      for (int k = 0; k < nl; ++k) {
        s->ocncs[k] = 0.5 * (s->ocncs[k] + s->ocncsp[k]);
        s->ocncn[k] = 0.5 * (s->ocncn[k] + s->ocncnp[k]);
      }
  } else {
    for (int k = 0; k < nl - 1; ++k) s->dpiat[k] = 0.5 * (s->dpiat[k] + s->dpiatp[k]);
    for (int k = 0; k < nl; ++k) {
      s->atmcs[k] = 0.5 * (s->atmcs[k] + s->atmcsp[k]);
      s->atmcn[k] = 0.5 * (s->atmcn[k] + s->atmcnp[k]);
    }
  }
}
static void avg(qgcm_model *m, const char *a, const char *b, size_t n) {
  QG_LAUNCH(m, "k_avg2", (unsigned)((n + 255) / 256), 256, 0, k_avg2, m->F(a), m->F(b), n);
}
void launch_tlavg_ocean(qgcm_model *m) {
  // y-slabs: both time levels carry valid halos, so the average needs no exchange
  const Grid &g = m->go;
  avg(m, "qo", "qom", g.lsz * g.nl);
  avg(m, "po", "pom", g.lsz * g.nl);
  avg(m, "sst", "sstm", (size_t)g.ld * g.nyt);
  QG_LAUNCH(m, "k_avg_scalars", 1, 32, 0, k_avg_scalars, m->d_scal, 0, g.cyclic, g.nl);
  QG_CUDA(cudaGetLastError());
}
void launch_tlavg_atmos(qgcm_model *m) {
  const Grid &g = m->ga;
  avg(m, "qa", "qam", g.lsz * g.nl);
  avg(m, "pa", "pam", g.lsz * g.nl);
  avg(m, "ast", "astm", (size_t)g.ld * g.nyt);
  avg(m, "hmixa", "hmixam", (size_t)g.ld * g.nyt);
  QG_LAUNCH(m, "k_avg_scalars", 1, 32, 0, k_avg_scalars, m->d_scal, 1, 1, g.nl);
  QG_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------
// constr (src/conhoms.F:44-314): xintp of interface pressure differences, and for
// channels the boundary line integrals.  Row sums on the device, the (tiny) rest on host.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_diff_rowsum(Grid g, const double *pa, const double *pb, double *rowsum) {
  __shared__ double red[8];
  const int j = blockIdx.x;
  double part = 0.0;
  for (int i = threadIdx.x; i < g.nxp; i += 256) {
    const double v = pa[(size_t)j * g.ld + i] - pb[(size_t)j * g.ld + i];
    part += (i == 0 || i == g.nxp - 1) ? 0.5 * v : v;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  if (lane == 0) red[w] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    rowsum[j] = t;
  }
}

// xintp(a - b) over the rows this rank owns (the whole grid on one GPU)
static double xintp_diff(qgcm_model *m, const Grid &g, const double *a, const double *b) {
  std::vector<double> rs(g.nyp);
  QG_LAUNCH(m, "k_diff_rowsum", g.nyp, 256, 0, k_diff_rowsum, g, a, b, m->d_red);
  QG_CUDA(cudaMemcpyAsync(rs.data(), m->d_red, sizeof(double) * g.nyp, cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
  const int lo = g.own0 + (g.wall_s() ? 1 : 0), hi = g.own1 - (g.wall_n() ? 1 : 0);
  double sump = 0.0;
  for (int j = lo; j < hi; ++j) sump += rs[j];
  if (g.wall_s()) sump += 0.5 * rs[0];
  if (g.wall_n()) sump += 0.5 * rs[g.nyp - 1];
  return sump;
}

// y-slabs: this rank's share of dpiocp(k), dpioc(k) (src/conhoms.F:44-314), then the totals
void constr_ocean_share(qgcm_model *m, std::vector<double> &v) {
  const Grid &g = m->go;
  const double *po = m->F("po"), *pom = m->F("pom");
  v.assign(2 * (g.nl - 1), 0.0);
  for (int k = 0; k < g.nl - 1; ++k) {
    v[2 * k] = xintp_diff(m, g, pom + (k + 1) * g.lsz, pom + k * g.lsz) * g.dx * g.dx;
    v[2 * k + 1] = xintp_diff(m, g, po + (k + 1) * g.lsz, po + k * g.lsz) * g.dx * g.dx;
  }
}
void constr_ocean_store(qgcm_model *m, const std::vector<double> &v) {
  qgcm_scalars s;
  QG_CUDA(cudaMemcpyAsync(&s, m->d_scal, sizeof(s), cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
  for (int k = 0; k < m->go.nl - 1; ++k) { s.dpiocp[k] = v[2 * k]; s.dpioc[k] = v[2 * k + 1]; }
  QG_CUDA(cudaMemcpy(m->d_scal, &s, sizeof(s), cudaMemcpyHostToDevice));
}

static void constr_lines(qgcm_model *m, const Grid &g, const LayerConsts &lc, const double *p, const double *pm,
                         double *cs, double *cn, double *csp, double *cnp) {
  const int nxp = g.nxp, nyp = g.nyp, nl = g.nl;
  // rows 1,2,nyp-1,nyp of both time levels: 4 rows x nl layers, copied to the host
  std::vector<double> rows((size_t)2 * nl * 4 * nxp);
  const int rj[4] = {0, 1, nyp - 2, nyp - 1};
  for (int t = 0; t < 2; ++t)
    for (int k = 0; k < nl; ++k)
      for (int r = 0; r < 4; ++r)
        QG_CUDA(cudaMemcpyAsync(&rows[(((size_t)t * nl + k) * 4 + r) * nxp], (t ? pm : p) + k * g.lsz + (size_t)rj[r] * g.ld,
                                sizeof(double) * nxp, cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
  const double dx = g.dx, dy = g.dx, f0 = m->fnot;
  for (int t = 0; t < 2; ++t) {
    double pins[NLMAX], pinn[NLMAX];
    double *os = t ? csp : cs, *on = t ? cnp : cn;
    for (int k = 0; k < nl; ++k) {
      const double *r0 = &rows[(((size_t)t * nl + k) * 4 + 0) * nxp], *r1 = r0 + nxp, *r2 = r1 + nxp, *r3 = r2 + nxp;
      double a = 0.5 * r0[0], b = 0.5 * r3[0], c = 0.5 * (r1[0] - r0[0]), d = 0.5 * (r3[0] - r2[0]);
      for (int i = 1; i < nxp - 1; ++i) {
        a += r0[i]; b += r3[i]; c += (r1[i] - r0[i]); d += (r3[i] - r2[i]);
      }
      a += 0.5 * r0[nxp - 1]; b += 0.5 * r3[nxp - 1];
      c += 0.5 * (r1[nxp - 1] - r0[nxp - 1]); d += 0.5 * (r3[nxp - 1] - r2[nxp - 1]);
      os[k] = c * (dx / dy); on[k] = d * (dx / dy);
      pins[k] = dx * a; pinn[k] = dx * b;
    }
    double ts[NLMAX], tn[NLMAX];
    for (int k = 0; k < nl; ++k) {
      double aps = 0.0, apn = 0.0;
      for (int j = 0; j < nl; ++j) { aps += lc.amat[k + nl * j] * pins[j]; apn += lc.amat[k + nl * j] * pinn[j]; }
      ts[k] = -os[k] + 0.5 * dy * f0 * f0 * aps;
      tn[k] = on[k] + 0.5 * dy * f0 * f0 * apn;
    }
    for (int k = 0; k < nl; ++k) { os[k] = ts[k]; on[k] = tn[k]; }
  }
}

void launch_constr(qgcm_model *m) {
  if (m->nranks > 1) { slab_constr(ranks_of(m)); return; }
  qgcm_scalars s;
  QG_CUDA(cudaMemcpyAsync(&s, m->d_scal, sizeof(s), cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
  if (m->has_ocean) {
    const Grid &g = m->go;
    const double *po = m->F("po"), *pom = m->F("pom");
    for (int k = 0; k < g.nl - 1; ++k) {
      s.dpiocp[k] = xintp_diff(m, g, pom + (k + 1) * g.lsz, pom + k * g.lsz) * g.dx * g.dx;
      s.dpioc[k] = xintp_diff(m, g, po + (k + 1) * g.lsz, po + k * g.lsz) * g.dx * g.dx;
    }
    if (g.cyclic) constr_lines(m, g, m->lo, po, pom, s.ocncs, s.ocncn, s.ocncsp, s.ocncnp);
  }
  if (m->has_atmos) {
    const Grid &g = m->ga;
    const double *pa = m->F("pa"), *pam = m->F("pam");
    for (int k = 0; k < g.nl - 1; ++k) {
      s.dpiatp[k] = xintp_diff(m, g, pam + k * g.lsz, pam + (k + 1) * g.lsz) * g.dx * g.dx;
      s.dpiat[k] = xintp_diff(m, g, pa + k * g.lsz, pa + (k + 1) * g.lsz) * g.dx * g.dx;
    }
    constr_lines(m, g, m->la, pa, pam, s.atmcs, s.atmcn, s.atmcsp, s.atmcnp);
  }
  QG_CUDA(cudaMemcpy(m->d_scal, &s, sizeof(s), cudaMemcpyHostToDevice));
}

}  // namespace qg

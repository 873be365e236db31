// Device-side qocdiag_out (src/qocdiag.F:303-683): the ocean vorticity tendency
// and its Jacobian, del-4th, del-6th and forcing/drag terms at the sub-sampled output points,
// so that the -Dqoc_diag decks (double_gyre_ocean_only) do not download po, pom, qo, qom,
// wekpo, entoc between oml and qgostep at every output interval (src/q-gcm.F:1237-1239).
// del-sqd and del-4th of pom are full-field passes into the modal work array (free between
// steps); the five terms are then evaluated only where the reference samples them.
#include "qgcm_internal.h"

namespace qg {

// out = del-sqd(in) with the mixed condition on solid walls / the periodic wrap
// (src/qocdiag.F:399-477; the same arithmetic as src/qgosubs.F:86-148)
__global__ void __launch_bounds__(256) k_qd_lap(const double *__restrict__ in, double *__restrict__ out, Grid g, double bcfac) {
  const int i = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
  if (i >= g.nxp) return;
  const int ld = g.ld, nxp = g.nxp, nyp = g.nyp;
  const double *r = in + (size_t)j * ld;
  double v;
  if (j == 0) {
    v = bcfac * (r[ld + i] - r[i]);
  } else if (j == nyp - 1) {
    v = bcfac * (r[i - ld] - r[i]);
  } else if (i == 0 || i == nxp - 1) {
    if (g.cyclic)      // column nxp repeats column 1
      v = (r[-ld] + r[nxp - 2] + r[1] + r[ld] - 4.0 * r[0]) * g.dxm2;
    else
      v = (i == 0) ? bcfac * (r[1] - r[0]) : bcfac * (r[nxp - 2] - r[nxp - 1]);
  } else {
    v = (r[i - ld] + r[i - 1] + r[i + 1] + r[i + ld] - 4.0 * r[i]) * g.dxm2;
  }
  out[(size_t)j * ld + i] = v;
}

struct QdArgs {
  Grid g;
  int k, nl, nsk, iw, jw, js0;        // layer (0-based); output points per direction; first sub-sampled row of this rank
  double adfac, ah2fac, ah4fac, foh, bdrfac, rdto;
  int forced;                         // 1: foh (wekpo - entoc), 2: foh entoc, 0: none
  int bottom;
  const double *po, *qo, *qom, *d2, *d4, *wek, *ent;
  double *out;                        // [5][nl][jw][iw]
};

__global__ void __launch_bounds__(256) k_qd_terms(QdArgs a) {
  const Grid &g = a.g;
  const int is = blockIdx.x * 256 + threadIdx.x, js = blockIdx.y;
  if (is >= a.iw) return;
  const int ld = g.ld, nxp = g.nxp;
  const int ig = is * a.nsk;
  const int jg = (a.js0 + js) * a.nsk;          // global p row
  const int j = jg - g.jg0;                     // local row
  double dq, jac = 0.0, t2 = 0.0, t4 = 0.0, ent = 0.0;
  const bool wall_ns = (jg == 0 || jg == g.nyp_g - 1);
  const bool edge = (ig == 0 || ig == nxp - 1);
  if (wall_ns || (edge && !g.cyclic)) {
    // time difference of qo on solid walls (:526-531, :594-600, :607-612)
    const size_t o = (size_t)j * ld + ig;
    dq = a.rdto * (a.qo[o] - a.qom[o]);
  } else {
    const int i = edge ? 0 : ig;                // periodic: the eastern column copies the western one
    const int im = (i == 0) ? nxp - 2 : i - 1, ip = i + 1;
    const double *q = a.qo + (size_t)j * ld, *p = a.po + (size_t)j * ld;
    const double *d2 = a.d2 + (size_t)j * ld, *d4 = a.d4 + (size_t)j * ld;
    const double d6p = g.dxm2 * (d4[i - ld] + d4[im] + d4[ip] + d4[i + ld] - 4.0 * d4[i]);
    t2 = a.ah2fac * d4[i];
    t4 = -a.ah4fac * d6p;
    jac = a.adfac * ((q[ip] - q[im]) * (p[i + ld] - p[i - ld]) + (q[i - ld] - q[i + ld]) * (p[ip] - p[im]) +
                     q[ip] * (p[ip + ld] - p[ip - ld]) - q[im] * (p[im + ld] - p[im - ld]) -
                     q[i + ld] * (p[ip + ld] - p[im + ld]) + q[i - ld] * (p[ip - ld] - p[im - ld]) +
                     p[i + ld] * (q[ip + ld] - q[im + ld]) - p[i - ld] * (q[ip - ld] - q[im - ld]) -
                     p[ip] * (q[ip + ld] - q[ip - ld]) + p[im] * (q[im + ld] - q[im - ld]));
    const size_t o = (size_t)j * ld + i;
    if (a.forced == 1) ent = a.foh * (a.wek[o] - a.ent[o]);
    else if (a.forced == 2) ent = a.foh * a.ent[o];
    if (a.bottom) ent = ent - a.bdrfac * d2[i];
    dq = jac + t2 + t4 + ent;
  }
  const size_t plane = (size_t)a.jw * a.iw, o = ((size_t)a.k * a.jw + js) * a.iw + is;
  a.out[o] = dq;
  a.out[(size_t)1 * a.nl * plane + o] = jac;
  a.out[(size_t)2 * a.nl * plane + o] = t2;
  a.out[(size_t)3 * a.nl * plane + o] = t4;
  a.out[(size_t)4 * a.nl * plane + o] = ent;
}

void qocdiag_size(qgcm_model *m, int nsk, int64_t *n) {
  if (!m->has_ocean) throw std::runtime_error("qgcm_qocdiag: no ocean in this model");
  if (nsk < 1) throw std::runtime_error("qgcm_qocdiag: nsko must be >= 1");
  *n = (int64_t)5 * sub_count(m->go.nxp, nsk) * sub_count(m->go.nyp_g, nsk) * m->go.nl;
}

// host(ipwk, jpwk, nlo, 5): dqdt, qotjac, qt2dif, qt4dif, qotent.  A y-slab fills the sub-sampled
// rows it owns (3 halo rows cover the del-6th stencil exactly as in the vorticity step).
void launch_qocdiag(qgcm_model *m, int nsk, double *host, int64_t n) {
  int64_t want;
  qocdiag_size(m, nsk, &want);
  if (n != want) throw std::runtime_error("qgcm_qocdiag: element count mismatch");
  const Grid &g = m->go;
  const qgcm_config &c = m->cfg;
  const int iw = sub_count(g.nxp, nsk), jwg = sub_count(g.nyp_g, nsk);
  const int g0 = g.jg0 + g.own0, g1 = g.jg0 + g.own1;
  const int js0 = (g0 + nsk - 1) / nsk, js1 = std::min(jwg, (g1 + nsk - 1) / nsk);
  const int jw = js1 - js0;
  if (jw <= 0) return;
  const size_t need = (size_t)5 * g.nl * jw * iw;
  if (m->pack_elems < need) {
    m->d_pack = (double *)dalloc(m, sizeof(double) * need);
    m->pack_elems = need;
  }
  const double bcfac = c.bccooc * g.dxm2 / (0.5 * c.bccooc + 1.0);
  double *d2 = m->wrk_o, *d4 = m->wrk_o + g.lsz;       // modal work array: free between steps
  const dim3 full((g.nxp + 255) / 256, g.nyp);
  for (int k = 0; k < g.nl; ++k) {
    QG_LAUNCH(m, "k_qd_lap", full, 256, 0, k_qd_lap, m->F("pom") + (size_t)k * g.lsz, d2, g, bcfac);
    QG_LAUNCH(m, "k_qd_lap", full, 256, 0, k_qd_lap, d2, d4, g, bcfac);
    QdArgs a;
    a.g = g; a.k = k; a.nl = g.nl; a.nsk = nsk; a.iw = iw; a.jw = jw; a.js0 = js0;
    a.adfac = 1.0 / (12.0 * g.dx * g.dx * m->fnot);
    a.ah2fac = m->lo.ah2[k] / m->fnot;
    a.ah4fac = m->lo.ah4[k] / m->fnot;
    a.foh = m->fnot / m->lo.h[k];
    a.bdrfac = 0.5 * (m->fnot < 0.0 ? -1.0 : 1.0) * c.delek / m->lo.h[g.nl - 1];
    a.rdto = 1.0 / m->dto;
    a.forced = (k == 0) ? 1 : (k == 1) ? 2 : 0;
    a.bottom = (k == g.nl - 1);
    a.po = m->F("po") + (size_t)k * g.lsz; a.qo = m->F("qo") + (size_t)k * g.lsz; a.qom = m->F("qom") + (size_t)k * g.lsz;
    a.d2 = d2; a.d4 = d4; a.wek = m->F("wekpo"); a.ent = m->F("entoc");
    a.out = m->d_pack;
    QG_LAUNCH(m, "k_qd_terms", dim3((iw + 255) / 256, jw), 256, 0, k_qd_terms, a);
  }
  m->hpo.walls_dirty = true;        // the work array's wall rows no longer hold zeros
  for (int t = 0; t < 5; ++t)
    for (int k = 0; k < g.nl; ++k)
      QG_CUDA(cudaMemcpyAsync(host + (((size_t)t * g.nl + k) * jwg + js0) * iw, m->d_pack + ((size_t)t * g.nl + k) * jw * iw,
                              sizeof(double) * (size_t)iw * jw, cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
}

}  // namespace qg

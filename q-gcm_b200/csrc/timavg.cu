// Device-side running sums and packed output:
//   tavini / tavocn / tavatm  (src/timavge.F:108-273, :425-617, :278-419)
//   avg_ocn_k247              (src/timavge.F:624-660, called every ocean step, src/q-gcm.F:1250-1252)
//   the sub-sampling of ocnc_out / atnc_out (src/nc_subs.F:880-890, :906-917, ...)
// The sums live beside the state in HBM under the reference's array names, so the host reads
// them with qgcm_get_field when tavout needs them (end of run) instead of downloading the
// full state at every accumulation; the sub-sampled read returns exactly the vector the
// reference hands to nf_put_vara_double.  These run at diagnostic cadence (once per model day
// / per output interval): plain coalesced one-pass kernels.
#include <cstring>

#include "qgcm_internal.h"

namespace qg {

struct TavArgs {
  Grid g;
  int atmos;                 // atmosphere: always periodic, u at i = nxp from its own column, no boundary outflow
  int sflux, nflux;          // sb_hflux / nb_hflux
  double ug;                 // geostrophic factor: ycexp/(dxo f0) ocean, 1/(dxa f0) atmosphere
  double rhu, rhv;           // +-0.5/(f0 hm): u = -ug dp/dy + rhu (tauy sum), v = ug dp/dx + rhv (taux sum)
  double tsbdy, tnbdy;
  const double *taux, *tauy, *wekp, *wekt, *fnet, *t, *p, *q;
  double *txav, *tyav, *wpav, *wtav, *fmav, *tav, *uuf, *tuf, *utuf, *vvf, *tvf, *vtvf, *pav, *qav;
};

// one thread per (i, j) of the p grid; every sum is a single add per call, as in the reference
__global__ void __launch_bounds__(256) k_tav(TavArgs a) {
  const Grid &g = a.g;
  const int i = blockIdx.x * 256 + threadIdx.x;       // 0-based p column
  const int j = blockIdx.y;                           // 0-based local p row
  if (i >= g.nxp) return;
  const size_t o = (size_t)j * g.ld + i;
  // p-grid sums (src/timavge.F:460-466, :315-320)
  a.txav[o] += a.taux[o];
  a.tyav[o] += a.tauy[o];
  if (a.wpav) a.wpav[o] += a.wekp[o];
  for (int k = 0; k < g.nl; ++k) {                    // :598-607, :402-411
    a.pav[(size_t)k * g.lsz + o] += a.p[(size_t)k * g.lsz + o];
    a.qav[(size_t)k * g.lsz + o] += a.q[(size_t)k * g.lsz + o];
  }
  // T-grid sums (:473-479, :327-333)
  if (i < g.nxt && j < g.nyt) {
    a.wtav[o] += a.wekt[o];
    a.fmav[o] += a.fnet[o];
    a.tav[o] += a.t[o];
  }
  // zonal advection at (i = 1..nxp, j = 1..nyt) (:487-531, :343-360)
  if (j < g.nyt) {
    double uu, tu, utu;
    const bool west = (i == 0), east = (i == g.nxp - 1);
    if ((west || east) && !g.cyclic) {                // finite box: no normal flux
      uu = 0.0;
      tu = a.t[(size_t)j * g.ld + (west ? 0 : g.nxt - 1)];
      utu = 0.0;
    } else {
      // cyclic ocean copies column 1 into column nxp; the atmosphere evaluates u in place
      const int iu = (east && !a.atmos) ? 0 : i;
      const size_t ou = (size_t)j * g.ld + iu;
      uu = -a.ug * (a.p[ou + g.ld] - a.p[ou]) + a.rhu * (a.tauy[ou + g.ld] + a.tauy[ou]);
      tu = (west || east) ? 0.5 * (a.t[(size_t)j * g.ld] + a.t[(size_t)j * g.ld + g.nxt - 1])
                          : 0.5 * (a.t[o] + a.t[o - 1]);
      utu = __dmul_rn(uu, tu);
    }
    a.uuf[o] += uu;
    a.tuf[o] += tu;
    a.utuf[o] += utu;
  }
  // meridional advection at (i = 1..nxt, j = 1..nyp) (:537-592, :365-399)
  if (i < g.nxt) {
    const bool south = (j == 0), north = (j == g.nyp - 1);
    double vv = 0.0, tv = 0.0, vtv = 0.0;
    bool store = true;
    if (south || north) {
      const bool wall = south ? g.wall_s() : g.wall_n();
      if (!wall) {
        store = false;                               // first/last halo row of a y-slab: owned by the neighbour
      } else {
        const int jt = south ? 0 : g.nyt - 1;       // T row next to the wall
        const double tin = a.t[(size_t)jt * g.ld + i];
        if (south ? a.sflux : a.nflux) {             // Ekman outflow carrying tsbdy / tnbdy
          vv = a.rhv * (a.taux[o + 1] + a.taux[o]);
          tv = 0.5 * (tin + (south ? a.tsbdy : a.tnbdy));
          vtv = __dmul_rn(vv, tv);
        } else {
          tv = tin;
        }
      }
    } else {
      vv = a.ug * (a.p[o + 1] - a.p[o]) + a.rhv * (a.taux[o + 1] + a.taux[o]);
      tv = 0.5 * (a.t[o] + a.t[o - g.ld]);
      vtv = __dmul_rn(vv, tv);
    }
    if (store) {
      a.vvf[o] += vv;
      a.tvf[o] += tv;
      a.vtvf[o] += vtv;
    }
  }
}

// acc += f over nl layers, two columns per lane
__global__ void __launch_bounds__(256) k_accum(double *__restrict__ acc, const double *__restrict__ f, int ld2, int rows, size_t lsz2) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= ld2) return;
  double2 *A = reinterpret_cast<double2 *>(acc) + (size_t)blockIdx.z * lsz2;
  const double2 *F = reinterpret_cast<const double2 *>(f) + (size_t)blockIdx.z * lsz2;
  for (int j = blockIdx.y; j < rows; j += gridDim.y) {
    const size_t o = (size_t)j * ld2 + i;
    double2 s = A[o];
    const double2 v = F[o];
    s.x += v.x;
    s.y += v.y;
    A[o] = s;
  }
}

// out(i, j, k) = f(i*nsk, j*nsk, k): the wrk vector of ocnc_out / atnc_out (src/nc_subs.F:880-890)
__global__ void __launch_bounds__(256) k_subsample(const double *__restrict__ f, double *__restrict__ out, int iw, int jw, int nsk, int ld,
                                                   size_t lsz, int jfirst) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;
  if (i >= iw) return;
  out[((size_t)k * jw + j) * iw + i] = f[(size_t)k * lsz + (size_t)(jfirst + j * nsk) * ld + (size_t)i * nsk];
}

static const char *OC_P[] = {"txocav", "tyocav", "wpocav"};
static const char *OC_T[] = {"wtocav", "fmocav", "sstav"};
static const char *OC_U[] = {"uufo", "tufo", "utufo"};
static const char *OC_V[] = {"vvfo", "tvfo", "vtvfo"};
static const char *AT_P[] = {"txatav", "tyatav"};
static const char *AT_T[] = {"wtatav", "fmatav", "astav"};
static const char *AT_U[] = {"uufa", "tufa", "utufa"};
static const char *AT_V[] = {"vvfa", "tvfa", "vtvfa"};

static void zero_field(qgcm_model *m, const char *name) {
  qgcm_model::Field &f = m->fields.at(name);
  QG_CUDA(cudaMemsetAsync(f.d, 0, sizeof(double) * f.elems, m->stream));
}

// allocate (first use) and zero the sums; which: 1 atmosphere, 2 ocean, 4 po_avg
static void tav_alloc(qgcm_model *m, int which, bool zero) {
  auto need = [&](const char *name, int nx, int ny, int nl, const Grid &g, bool slab) {
    if (!m->fields.count(name)) add_field(m, name, nx, ny, nl, g.ld, g.lsz, slab ? &g : nullptr);   // dalloc zero-fills
    else if (zero) zero_field(m, name);
  };
  if ((which & 2) && m->has_ocean) {
    const Grid &g = m->go;
    for (const char *n : OC_P) need(n, g.nxp, g.nyp, 1, g, true);
    for (const char *n : OC_T) need(n, g.nxt, g.nyt, 1, g, true);
    for (const char *n : OC_U) need(n, g.nxp, g.nyt, 1, g, true);
    for (const char *n : OC_V) need(n, g.nxt, g.nyp, 1, g, true);
    need("pocav", g.nxp, g.nyp, g.nl, g, true);
    need("qocav", g.nxp, g.nyp, g.nl, g, true);
    if (zero) m->nsumoc = 0;
  }
  if ((which & 4) && m->has_ocean) {
    const Grid &g = m->go;
    need("po_avg", g.nxp, g.nyp, g.nl, g, true);
    if (zero) m->nsum_ocavg = 0;
  }
  if ((which & 1) && m->has_atmos) {
    const Grid &g = m->ga;
    for (const char *n : AT_P) need(n, g.nxp, g.nyp, 1, g, false);
    for (const char *n : AT_T) need(n, g.nxt, g.nyt, 1, g, false);
    for (const char *n : AT_U) need(n, g.nxp, g.nyt, 1, g, false);
    for (const char *n : AT_V) need(n, g.nxt, g.nyp, 1, g, false);
    need("patav", g.nxp, g.nyp, g.nl, g, false);
    need("qatav", g.nxp, g.nyp, g.nl, g, false);
    if (zero) m->nsumat = 0;
  }
}

void launch_tavini(qgcm_model *m) { tav_alloc(m, 7, true); }

void launch_tavocn(qgcm_model *m) {
  if (!m->has_ocean) return;
  tav_alloc(m, 2, false);
  const Grid &g = m->go;
  TavArgs a;
  a.g = g;
  a.atmos = 0;
  a.sflux = m->sb_hflux; a.nflux = m->nb_hflux;
  a.ug = m->cfg.ycexp * g.rdxf0;
  const double rh = 0.5 / (m->fnot * m->cfg.hmoc);
  a.rhu = rh; a.rhv = -rh;
  a.tsbdy = m->cfg.tsbdy; a.tnbdy = m->cfg.tnbdy;
  a.taux = m->F("tauxo"); a.tauy = m->F("tauyo"); a.wekp = m->F("wekpo"); a.wekt = m->F("wekto");
  a.fnet = m->F("fnetoc"); a.t = m->F("sst"); a.p = m->F("po"); a.q = m->F("qo");
  a.txav = m->F("txocav"); a.tyav = m->F("tyocav"); a.wpav = m->F("wpocav");
  a.wtav = m->F("wtocav"); a.fmav = m->F("fmocav"); a.tav = m->F("sstav");
  a.uuf = m->F("uufo"); a.tuf = m->F("tufo"); a.utuf = m->F("utufo");
  a.vvf = m->F("vvfo"); a.tvf = m->F("tvfo"); a.vtvf = m->F("vtvfo");
  a.pav = m->F("pocav"); a.qav = m->F("qocav");
  QG_LAUNCH(m, "k_tav", dim3((g.nxp + 255) / 256, g.nyp), 256, 0, k_tav, a);
  m->nsumoc++;
}

void launch_tavatm(qgcm_model *m) {
  if (!m->has_atmos) return;
  tav_alloc(m, 1, false);
  const Grid &g = m->ga;
  TavArgs a;
  a.g = g;
  a.atmos = 1;
  a.sflux = a.nflux = 0;
  a.ug = g.rdxf0;
  const double rh = 0.5 / (m->fnot * m->cfg.hmat);
  a.rhu = -rh; a.rhv = rh;
  a.tsbdy = a.tnbdy = 0.0;
  a.taux = m->F("tauxa"); a.tauy = m->F("tauya"); a.wekp = nullptr; a.wekt = m->F("wekta");
  a.fnet = m->F("fnetat"); a.t = m->F("ast"); a.p = m->F("pa"); a.q = m->F("qa");
  a.txav = m->F("txatav"); a.tyav = m->F("tyatav"); a.wpav = nullptr;
  a.wtav = m->F("wtatav"); a.fmav = m->F("fmatav"); a.tav = m->F("astav");
  a.uuf = m->F("uufa"); a.tuf = m->F("tufa"); a.utuf = m->F("utufa");
  a.vvf = m->F("vvfa"); a.tvf = m->F("tvfa"); a.vtvf = m->F("vtvfa");
  a.pav = m->F("patav"); a.qav = m->F("qatav");
  QG_LAUNCH(m, "k_tav", dim3((g.nxp + 255) / 256, g.nyp), 256, 0, k_tav, a);
  m->nsumat++;
}

void launch_avg_ocn_k247(qgcm_model *m) {
  if (!m->has_ocean) return;
  tav_alloc(m, 4, false);
  const Grid &g = m->go;
  const int ld2 = g.ld / 2;
  QG_LAUNCH(m, "k_accum", dim3((ld2 + 255) / 256, std::min(g.nyp, 592), g.nl), 256, 0, k_accum, m->F("po_avg"), m->F("po"), ld2, g.nyp,
            g.lsz / 2);
  m->nsum_ocavg++;
}

void field_sub_size(qgcm_model *m, const char *name, int nsk, int64_t *n) {
  auto it = m->fields.find(name);
  if (it == m->fields.end() || !it->second.ld) throw std::runtime_error(std::string("qgcm_get_field_sub: no gridded field '") + name + "'");
  if (nsk < 1) throw std::runtime_error("qgcm_get_field_sub: nsk must be >= 1");
  const qgcm_model::Field &f = it->second;
  *n = (int64_t)sub_count(f.nx, nsk) * sub_count(f.nyg, nsk) * f.nl;
}

// host(iw, jw, nl) <- every nsk-th point of the named field.  A y-slab fills the sub-sampled
// rows it owns and leaves the rest of the host vector untouched (as qgcm_get_field does).
void get_field_sub(qgcm_model *m, const char *name, int nsk, double *host, int64_t n) {
  int64_t want;
  field_sub_size(m, name, nsk, &want);
  if (n != want) throw std::runtime_error(std::string("qgcm_get_field_sub '") + name + "': element count mismatch");
  const qgcm_model::Field &f = m->fields.at(name);
  const int iw = sub_count(f.nx, nsk), jwg = sub_count(f.nyg, nsk);
  // sub-sampled global rows jg = js*nsk that fall into the owned local rows [o0, o1)
  const int g0 = f.joff + f.o0, g1 = f.joff + f.o1;
  const int js0 = (g0 + nsk - 1) / nsk, js1 = std::min(jwg, (g1 + nsk - 1) / nsk);
  const int jw = js1 - js0;
  if (jw <= 0) return;
  const size_t need = (size_t)iw * jw * f.nl;
  if (m->pack_elems < need) {
    m->d_pack = (double *)dalloc(m, sizeof(double) * need);   // the old, smaller buffer stays in allocs until destroy
    m->pack_elems = need;
  }
  QG_LAUNCH(m, "k_subsample", dim3((iw + 255) / 256, jw, f.nl), 256, 0, k_subsample, f.d, m->d_pack, iw, jw, nsk, f.ld, f.lsz,
            js0 * nsk - f.joff);
  for (int k = 0; k < f.nl; ++k)
    QG_CUDA(cudaMemcpyAsync(host + ((size_t)k * jwg + js0) * iw, m->d_pack + (size_t)k * jw * iw, sizeof(double) * (size_t)iw * jw,
                            cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
}

}  // namespace qg

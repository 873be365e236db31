// Atmosphere mixed layer (aml + amladf, src/amlsubs.F:47-563) and the coupled forcing
// xforc (src/xfosubs.F:52-858) with its helpers auvbcu/bcuini/wts2bb (bicubic
// coarse->fine regridding of the wind, :997-1728), bilint (:891-993) and fsprim (:862-887).
//
// The atmosphere grids are small (385 x 97 at most) and stay L2 resident, so aml is one
// thread per T cell reading its stencils straight from global memory.  xforc works on the
// ocean-resolution atmosphere grid (nxta*ndxr+1) x (nyta*ndxr+1) -- larger than the ocean
// itself: one fused kernel interpolates the wind (16-term bicubic dot products), subtracts
// the ocean velocity (tau_udiff), applies the quadratic drag law and writes the stress once
// (fine grid + the ocean window scaled by rhoat/rhooc); everything downstream (sampling,
// side integrals, box averages of the Ekman pumping, line integrals) reads that stress
// without materialising wektaor, u1ator or v1ator.
#include <cmath>
#include <cstring>

#include "qgcm_internal.h"

namespace qg {

// ======================================================================================
// aml
// ======================================================================================
struct AmlArgs {
  Grid g;
  int nl;
  double hmat, hmamin, hmainv, hdrcdt, diabcr, entfac, xcexp, xbfac, dface, cface, tat1, rrcpat, tdt;
  double d2tfac, d4tfac, hmdfac, rdxf0, hdxm1;
  double afacdp[NLMAX];
  const double *pa, *pam, *ast, *astm, *hm, *hmm, *uek, *vek, *fnet, *wekt, *xc1, *dtop;
  double *astnew, *hmnew, *xfa, *entat;
  double *part;    // [2][nblocks]
  int nblocks;
  double *rowsum;  // [nyp]
  qgcm_scalars *sc;
};

// 5-point sum of amladf with the no-flux rows of the temperature equation
// (src/amlsubs.F:300-382): rows 1 and nyta drop the missing neighbour
__device__ __forceinline__ double aml_d2(const double *f, int i, int j, int nxt, int nyt, int ld) {
  const int im = (i == 0) ? nxt - 1 : i - 1, ip = (i == nxt - 1) ? 0 : i + 1;
  const double *r = f + (size_t)j * ld;
  if (j == 0) return r[im] + r[ip] + r[ld + i] - 3.0 * r[i];
  if (j == nyt - 1) return r[i - ld] + r[im] + r[ip] - 3.0 * r[i];
  return r[i - ld] + r[im] + r[ip] + r[ld + i] - 4.0 * r[i];
}

__global__ void __launch_bounds__(256) k_aml_step(AmlArgs a) {
  __shared__ double red[2][8];
  const Grid &g = a.g;
  const int nxt = g.nxt, nyt = g.nyt, ld = g.ld;
  const int i = blockIdx.x * 64 + (threadIdx.x & 63), j = blockIdx.y * 4 + (threadIdx.x >> 6);
  double cfr = 0.0, cen = 0.0;
  if (i < nxt && j < nyt) {
    const int im = (i == 0) ? nxt - 1 : i - 1, ip = (i == nxt - 1) ? 0 : i + 1;
    const size_t c = (size_t)j * ld + i;
    const double *pa = a.pa;
#define PA1(ii, jj) pa[(size_t)(jj) * ld + (ii)]
#define AST(ii, jj) a.ast[(size_t)(jj) * ld + (ii)]
#define HM(ii, jj) a.hm[(size_t)(jj) * ld + (ii)]
#define HMM(ii, jj) a.hmm[(size_t)(jj) * ld + (ii)]
#define UE(ii, jj) a.uek[(size_t)(jj) * ld + (ii)]
#define VE(ii, jj) a.vek[(size_t)(jj) * ld + (ii)]
    const double hmat = a.hmat;
    // ---- amladf: C-grid flux-form advection by geostrophic + Ekman flow, src/amlsubs.F:291-460 ----
    const double um = -a.rdxf0 * (PA1(i, j + 1) - PA1(i, j)) + UE(i, j);
    const double up = -a.rdxf0 * (PA1(i + 1, j + 1) - PA1(i + 1, j)) + UE(i + 1, j);
    const double tm = AST(i, j) + AST(im, j), tp = AST(i, j) + AST(ip, j);
    const double hm = HM(i, j) + HM(im, j), hp = HM(i, j) + HM(ip, j);
    const double xadvt = a.hdxm1 * (up * tp - um * tm);
    const double xadvh = a.hdxm1 * (up * hp - um * hm);
    double yadvt, yadvh, d2h;
    if (j == 0) {
      const double vm = VE(i, 0);
      const double vp = a.rdxf0 * (PA1(i + 1, 1) - PA1(i, 1)) + VE(i, 1);
      yadvt = a.hdxm1 * vp * (AST(i, 1) + AST(i, 0));
      yadvh = a.hdxm1 * (vp * (HM(i, 1) + HM(i, 0)) - vm * (HM(i, 0) + hmat));
      d2h = hmat + HMM(im, 0) + HMM(ip, 0) + HMM(i, 1) - 4.0 * HMM(i, 0);
    } else if (j == nyt - 1) {
      const double vm = a.rdxf0 * (PA1(i + 1, j) - PA1(i, j)) + VE(i, j);
      const double vp = VE(i, j + 1);
      yadvt = a.hdxm1 * (-vm * (AST(i, j) + AST(i, j - 1)));
      yadvh = a.hdxm1 * (vp * (hmat + HM(i, j)) - vm * (HM(i, j) + HM(i, j - 1)));
      d2h = HMM(i, j - 1) + HMM(im, j) + HMM(ip, j) + hmat - 4.0 * HMM(i, j);
    } else {
      const double vm = a.rdxf0 * (PA1(i + 1, j) - PA1(i, j)) + VE(i, j);
      const double vp = a.rdxf0 * (PA1(i + 1, j + 1) - PA1(i, j + 1)) + VE(i, j + 1);
      yadvt = a.hdxm1 * (vp * (AST(i, j + 1) + AST(i, j)) - vm * (AST(i, j) + AST(i, j - 1)));
      yadvh = a.hdxm1 * (vp * (HM(i, j + 1) + HM(i, j)) - vm * (HM(i, j) + HM(i, j - 1)));
      d2h = HMM(i, j - 1) + HMM(im, j) + HMM(ip, j) + HMM(i, j + 1) - 4.0 * HMM(i, j);
    }
    // del2 / del4 of the lagged temperature, src/amlsubs.F:470-560
    const double d2c = aml_d2(a.astm, i, j, nxt, nyt, ld);
    const double d2w = aml_d2(a.astm, im, j, nxt, nyt, ld), d2e = aml_d2(a.astm, ip, j, nxt, nyt, ld);
    double d4;
    if (j == 0) d4 = d2w + d2e + aml_d2(a.astm, i, 1, nxt, nyt, ld) - 3.0 * d2c;
    else if (j == nyt - 1) d4 = aml_d2(a.astm, i, j - 1, nxt, nyt, ld) + d2w + d2e - 3.0 * d2c;
    else d4 = aml_d2(a.astm, i, j - 1, nxt, nyt, ld) + d2w + d2e + aml_d2(a.astm, i, j + 1, nxt, nyt, ld) - 4.0 * d2c;
    const double tmrhs = -(xadvt + yadvt) + a.d2tfac * d2c - a.d4tfac * d4;
    const double hmrhs = -(xadvh + yadvh) + a.hmdfac * d2h;
    // ---- aml: predict hmixa and ast, entrainment, convection, src/amlsubs.F:104-166 ----
    const double astm = a.astm[c], hmm = a.hmm[c];
    double hnew, dtfix;
    if (astm <= a.diabcr) {
      const double dhdiab = a.hdrcdt * (hmm - hmat) / (a.tat1 - astm);
      hnew = hmm + a.tdt * hmrhs - dhdiab;
      const double dhfix = fmax(a.hmamin - hnew, 0.0);
      hnew = hnew + dhfix;
      dtfix = dhfix * (a.tat1 - astm) / hmm;
    } else {
      hnew = hmat;
      dtfix = 0.0;
    }
    const double trhtot = tmrhs + a.rrcpat * a.fnet[c] / hmm - a.hmainv * a.wekt[c] * astm;
    double astnew = astm + a.tdt * trhtot + dtfix;
    const double xfaent = a.xbfac * (hmm - hmat) + a.dface * (a.xcexp * astm + a.xc1[c]);
    const double dtanew = a.tat1 - astnew;
    const double conena = a.entfac * a.hm[c] * fmin(0.0, dtanew);
    a.xfa[c] = xfaent - a.xcexp * conena;
    astnew = astnew + fmin(0.0, dtanew);
    cfr = (dtanew >= 0.0) ? 0.0 : 1.0;     // 0.5 - sign(0.5, dtanew)
    cen = -conena;
    a.astnew[c] = astnew;
    a.hmnew[c] = hnew;
#undef PA1
#undef AST
#undef HM
#undef HMM
#undef UE
#undef VE
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cfr += __shfl_down_sync(0xffffffffu, cfr, o);
    cen += __shfl_down_sync(0xffffffffu, cen, o);
  }
  if (lane == 0) { red[0][w] = cfr; red[1][w] = cen; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int q = 0; q < 8; ++q) { t0 += red[0][q]; t1 += red[1][q]; }
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    a.part[b] = t0;
    a.part[a.nblocks + b] = t1;
  }
}

// entat: 4-point average of xfa onto p points (periodic in x, half-cell rows at the walls),
// plus the eta and topography terms that live on p points (src/amlsubs.F:173-212); one
// block per p row, which also leaves the xintp row sum of that row
__global__ void __launch_bounds__(256) k_aml_entat(AmlArgs a) {
  __shared__ double red[8];
  const Grid &g = a.g;
  const int j = blockIdx.x, nxp = g.nxp, nyp = g.nyp, nxt = g.nxt, nyt = g.nyt, ld = g.ld;
  double part = 0.0;
  for (int i = threadIdx.x; i < nxp; i += 256) {
    int im = i - 1, ic = i;
    if (i == 0 || i == nxp - 1) { im = nxt - 1; ic = 0; }
#define X(ii, jj) a.xfa[(size_t)(jj) * ld + (ii)]
    double v;
    if (j == 0) v = 0.5 * (X(im, 0) + X(ic, 0));
    else if (j == nyp - 1) v = 0.5 * (X(im, nyt - 1) + X(ic, nyt - 1));
    else v = 0.25 * (X(im, j - 1) + X(ic, j - 1) + X(im, j) + X(ic, j));
#undef X
    double adpsum = 0.0;
    const size_t c = (size_t)j * ld + i;
    for (int l = 0; l < a.nl - 1; ++l) adpsum = adpsum + a.afacdp[l] * (a.pam[l * g.lsz + c] - a.pam[(l + 1) * g.lsz + c]);
    v = v + adpsum + a.cface * a.dtop[c];
    a.entat[c] = v;
    part += (i == 0 || i == nxp - 1) ? 0.5 * v : v;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  if (lane == 0) red[w] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += red[q];
    a.rowsum[j] = t;
  }
}

// cfraat, centat, xan(1), enisat(1), eninat(1) (src/amlsubs.F:214-236)
__global__ void __launch_bounds__(256) k_aml_finish(AmlArgs a) {
  __shared__ double red[8];
  const Grid &g = a.g;
  const double cfr = block256_range_sum(a.part, 0, a.nblocks, red);
  const double cen = block256_range_sum(a.part + a.nblocks, 0, a.nblocks, red);
  const double sump = block256_range_sum(a.rowsum, 1, g.nyp - 1, red);
  if (threadIdx.x != 0) return;
  a.sc->cfraat = cfr * g.norm;
  a.sc->centat = cen * g.dx * g.dx;
  a.sc->xan[0] = (sump + 0.5 * (a.rowsum[0] + a.rowsum[g.nyp - 1])) * g.dx * g.dx;
  a.sc->enisat[0] = g.dx * a.rowsum[0];
  a.sc->eninat[0] = g.dx * a.rowsum[g.nyp - 1];
}

void launch_aml(qgcm_model *m) {
  if (!m->has_atmos) throw std::runtime_error("qgcm_aml: this model has no atmosphere (ocean_only)");
  const Grid &g = m->ga;
  const qgcm_config &c = m->cfg;
  AmlArgs a;
  a.g = g; a.nl = g.nl;
  a.hmat = c.hmat; a.hmamin = c.hmamin; a.hmainv = 1.0 / c.hmat;
  a.hdrcdt = c.hmadmp * m->rrcpat * g.tdt;
  a.tat1 = c.tat[0];
  a.diabcr = c.tat[0] - 2.0 * a.hdrcdt;
  a.entfac = 1.0 / (g.tdt * (c.tat[1] - c.tat[0]));
  a.xcexp = c.xcexp; a.xbfac = c.xcexp * c.bface; a.dface = c.dface; a.cface = c.cface;
  a.rrcpat = m->rrcpat; a.tdt = g.tdt;
  a.d2tfac = c.at2d * g.dxm2; a.d4tfac = c.at4d * g.dxm2 * g.dxm2; a.hmdfac = c.ahmd * g.dxm2;
  a.rdxf0 = g.rdxf0; a.hdxm1 = g.hdxm1;
  for (int l = 0; l < NLMAX; ++l) a.afacdp[l] = (l < g.nl - 1) ? c.aface[l] / c.gpat[l] : 0.0;
  a.pa = m->F("pa"); a.pam = m->F("pam"); a.ast = m->F("ast"); a.astm = m->F("astm");
  a.hm = m->F("hmixa"); a.hmm = m->F("hmixam"); a.uek = m->F("uekat"); a.vek = m->F("vekat");
  a.fnet = m->F("fnetat"); a.wekt = m->F("wekta"); a.xc1 = m->F("xc1ast"); a.dtop = m->F("dtopat");
  a.astnew = m->astnew; a.hmnew = m->hmnew; a.xfa = m->xfa; a.entat = m->F("entat");
  dim3 grid((g.nxt + 63) / 64, (g.nyt + 3) / 4);
  a.nblocks = grid.x * grid.y;
  a.part = m->d_red_a;
  a.rowsum = m->d_red_a + 2 * (size_t)a.nblocks;
  a.sc = m->d_scal;
  if (m->red_elems < 2 * (size_t)a.nblocks + g.nyp) throw std::runtime_error("aml: reduction scratch too small");
  QG_LAUNCH(m, "k_aml_step", grid, 256, 0, k_aml_step, a);
  QG_LAUNCH(m, "k_aml_entat", g.nyp, 256, 0, k_aml_entat, a);
  QG_LAUNCH(m, "k_aml_finish", 1, 256, 0, k_aml_finish, a);
  QG_CUDA(cudaGetLastError());
  // astm <- ast, ast <- new; hmixam <- hmixa, hmixa <- new (src/amlsubs.F:163-166): rotations
  double *old = m->fields.at("astm").d;
  m->fields.at("astm").d = m->fields.at("ast").d;
  m->fields.at("ast").d = m->astnew;
  m->astnew = old;
  old = m->fields.at("hmixam").d;
  m->fields.at("hmixam").d = m->fields.at("hmixa").d;
  m->fields.at("hmixa").d = m->hmnew;
  m->hmnew = old;
}

// ======================================================================================
// xforc
// ======================================================================================
struct XfArgsK {
  Grid ga, go;
  int ndxr, nxf, nyf, ldf;             // fine grid
  int iocoff, jocoff;                  // ocean p-point offsets in the fine grid (0-based: fine = ocean + off)
  int tau_udiff, cyclic_oc;
  double hxafac, hxofac, zbfcat, zbfcoc, raoro;
  double cdrfaa, cdrfab, qu2faa, qu2fab;
  double uvekfc, hmrdxa, dxo;
  const double *pam, *pom;
  double *u1, *v1, *taux, *tauy;
  const double *stb;
  double *tauxa, *tauya, *uek, *vek, *wekta, *wekpa, *tauxo, *tauyo;
  qgcm_scalars *sc;
};

// geostrophic wind of layer 1 at atmosphere p points (src/xfosubs.F:186-214)
__global__ void __launch_bounds__(256) k_xf_wind(XfArgsK a) {
  const Grid &g = a.ga;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= g.nxp) return;
  const int ld = g.ld, nxp = g.nxp, nyp = g.nyp;
  const double *p = a.pam;
  double u, v;
  if (j == 0) {
    u = -a.zbfcat * (p[ld + i] - p[i]);
    v = 0.0;
  } else if (j == nyp - 1) {
    u = -a.zbfcat * (p[(size_t)j * ld + i] - p[(size_t)(j - 1) * ld + i]);
    v = 0.0;
  } else {
    const int ic = (i == nxp - 1) ? 0 : i;       // eastern column copies the western one
    const int iw = (ic == 0) ? nxp - 2 : ic - 1, ie = ic + 1;
    u = -a.hxafac * (p[(size_t)(j + 1) * ld + ic] - p[(size_t)(j - 1) * ld + ic]);
    v = a.hxafac * (p[(size_t)j * ld + ie] - p[(size_t)j * ld + iw]);
  }
  a.u1[(size_t)j * ld + i] = u;
  a.v1[(size_t)j * ld + i] = v;
}

// velocity difference over the ocean, quadratic drag law and the stores of one fine p point
// (src/xfosubs.F:250-354, :554-559)
__device__ __forceinline__ void xf_stress_point(const XfArgsK &a, int i, int j, double usum, double vsum) {
  // velocity difference over the ocean (tau_udiff), src/xfosubs.F:250-300
  const Grid &go = a.go;
  const int io = i - a.iocoff, jo = j - a.jocoff;
  const bool over_ocean = io >= 0 && io < go.nxp && jo >= 0 && jo < go.nyp;
  double cdrfac = a.cdrfaa, qu2fac = a.qu2faa;
  if (a.tau_udiff && over_ocean) {
    cdrfac = a.cdrfab;
    qu2fac = a.qu2fab;
    const double *p = a.pom;
    const int ld = go.ld, nxp = go.nxp, nyp = go.nyp;
    double u1oc, v1oc;
    if (jo == 0) {
      u1oc = -a.zbfcoc * (p[ld + io] - p[io]);
      v1oc = 0.0;
    } else if (jo == nyp - 1) {
      u1oc = -a.zbfcoc * (p[(size_t)jo * ld + io] - p[(size_t)(jo - 1) * ld + io]);
      v1oc = 0.0;
    } else if (io == 0 || io == nxp - 1) {
      if (a.cyclic_oc) {
        u1oc = -a.hxofac * (p[(size_t)(jo + 1) * ld + io] - p[(size_t)(jo - 1) * ld + io]);
        v1oc = a.hxofac * (p[(size_t)jo * ld + 1] - p[(size_t)jo * ld + nxp - 2]);
      } else {
        u1oc = 0.0;
        v1oc = (io == 0) ? a.zbfcoc * (p[(size_t)jo * ld + 1] - p[(size_t)jo * ld])
                         : a.zbfcoc * (p[(size_t)jo * ld + nxp - 1] - p[(size_t)jo * ld + nxp - 2]);
      }
    } else {
      u1oc = -a.hxofac * (p[(size_t)(jo + 1) * ld + io] - p[(size_t)(jo - 1) * ld + io]);
      v1oc = a.hxofac * (p[(size_t)jo * ld + io + 1] - p[(size_t)jo * ld + io - 1]);
    }
    usum = usum - u1oc;
    vsum = vsum - v1oc;
  }
  // quadratic drag law, src/xfosubs.F:319-354
  const double scasqd = -0.5 + 0.5 * sqrt(1.0 + qu2fac * (usum * usum + vsum * vsum));
  const double scashr = sqrt(scasqd);
  const double cdochi = cdrfac * scashr / (1.0 + scasqd);
  const double tx = cdochi * (usum - scashr * vsum), ty = cdochi * (vsum + scashr * usum);
  a.taux[(size_t)j * a.ldf + i] = tx;
  a.tauy[(size_t)j * a.ldf + i] = ty;
  if (over_ocean) {   // src/xfosubs.F:554-559
    a.tauxo[(size_t)jo * go.ld + io] = a.raoro * tx;
    a.tauyo[(size_t)jo * go.ld + io] = a.raoro * ty;
  }
}

// bicubic wind on the fine grid, ocean-velocity correction, quadratic drag; one thread per
// fine p point.  grid (ceil(nxf/128), nyf)
// `edge` != 0: only the fine rows of the southernmost and northernmost coarse cells (the separable
// kernel below does the rest): blockIdx.y < ndxr -> southern rows, otherwise the northern ones
__global__ void __launch_bounds__(128) k_xf_stress(XfArgsK a, int edge) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  if (edge && j >= a.ndxr) j = (a.ga.nyt - 1) * a.ndxr + (j - a.ndxr);
  if (i >= a.nxf) return;
  const Grid &ga = a.ga;
  const int n = a.ndxr, nxta = ga.nxt, nyta = ga.nyt, lda = ga.ld;
  // coarse cell and position inside it; the last fine column repeats the first
  // (src/xfosubs.F:1211-1214) and the last fine row belongs to the northern cells with jj = ndxr
  const int ifi = (i == a.nxf - 1) ? 0 : i;
  const int ic = ifi / n, ii = ifi - ic * n;
  int jc = j / n, jj = j - jc * n;
  if (j == a.nyf - 1) { jc = nyta - 1; jj = n; }
  const bool south = (jc == 0), north = (jc == nyta - 1);
  const int icm1 = (ic == 0) ? nxta - 1 : ic - 1, icp2 = (ic + 2) % nxta;
  const int ix[4] = {icm1, ic, ic + 1, icp2};
  // weights are stored [variant][k][fine point] so that neighbouring lanes read neighbouring doubles
  const int npt = (n + 1) * (n + 1);
  const double *wu = a.stb + (size_t)(south ? 1 : (north ? 3 : 0)) * npt * 16 + (ii + (n + 1) * jj);
  const double *wv = a.stb + (size_t)(south ? 2 : (north ? 4 : 0)) * npt * 16 + (ii + (n + 1) * jj);
  double usum = 0.0, vsum = 0.0;
#pragma unroll
  for (int row = 0; row < 4; ++row) {
    const int jd = row - 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double ud, vd;
      if (south && jd == -1) {
        ud = 0.0;
        vd = a.u1[ix[q]];                                    // u on the boundary pads v (vy = -ux)
      } else if (north && jd == 2) {
        ud = 0.0;
        vd = a.u1[(size_t)nyta * lda + ix[q]];
      } else {
        ud = a.u1[(size_t)(jc + jd) * lda + ix[q]];
        vd = a.v1[(size_t)(jc + jd) * lda + ix[q]];
      }
      const double wuq = __ldg(wu + (size_t)(4 * row + q) * npt);
      const double wvq = (south || north) ? __ldg(wv + (size_t)(4 * row + q) * npt) : wuq;   // one table away from the walls
      usum = usum + ud * wuq;
      vsum = vsum + vd * wvq;
    }
  }
  xf_stress_point(a, i, j, usum, vsum);
}

// Away from the zonal boundaries the bicubic patch with centred-difference derivatives is the
// separable Catmull-Rom spline: stb(k = 4 row + q; ii, jj) = ay_row(jj) bx_q(ii).  One thread then
// takes a whole column of ndxr fine points of a coarse cell: the 32 coarse values are loaded and
// combined in x once (16 multiply-adds per component) and every fine point costs 4 more per
// component, instead of 16 with 16 weight loads.  The 1-D factors are read from the reference's own
// table (bx_q(ii) = stb(4+q; ii, 0), ay_row(jj) = stb(4 row + 1; 0, jj)), so no new constants enter;
// sums are associated differently from the 16-term loop of src/xfosubs.F:1100-1180 (1e-16 level).
// grid (ceil(nxf/128), nyta-2): coarse rows 1 .. nyta-2.
__global__ void __launch_bounds__(128) k_xf_stress_sep(XfArgsK a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int jc = blockIdx.y + 1;
  if (i >= a.nxf) return;
  const Grid &ga = a.ga;
  const int n = a.ndxr, nxta = ga.nxt, lda = ga.ld;
  const int ifi = (i == a.nxf - 1) ? 0 : i;
  const int ic = ifi / n, ii = ifi - ic * n;
  const int icm1 = (ic == 0) ? nxta - 1 : ic - 1, icp2 = (ic + 2) % nxta;
  const int ix[4] = {icm1, ic, ic + 1, icp2};
  const int npt = (n + 1) * (n + 1);
  double bx[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) bx[q] = __ldg(a.stb + (size_t)(4 + q) * npt + ii);
  double ux[4], vx[4];
#pragma unroll
  for (int row = 0; row < 4; ++row) {
    double su = 0.0, sv = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      su = su + a.u1[(size_t)(jc + row - 1) * lda + ix[q]] * bx[q];
      sv = sv + a.v1[(size_t)(jc + row - 1) * lda + ix[q]] * bx[q];
    }
    ux[row] = su; vx[row] = sv;
  }
  for (int jj = 0; jj < n; ++jj) {
    double usum = 0.0, vsum = 0.0;
#pragma unroll
    for (int row = 0; row < 4; ++row) {
      const double ay = __ldg(a.stb + (size_t)(4 * row + 1) * npt + (size_t)(n + 1) * jj);
      usum = usum + ux[row] * ay;
      vsum = vsum + vx[row] * ay;
    }
    xf_stress_point(a, i, jc * n + jj, usum, vsum);
  }
}

// tauxa/tauya samples and the side integrals vekat, uekat (src/xfosubs.F:362-407); one thread
// per atmosphere p point
__global__ void __launch_bounds__(256) k_xf_sample(XfArgsK a) {
  const Grid &g = a.ga;
  const int ia = blockIdx.x * blockDim.x + threadIdx.x, ja = blockIdx.y;
  if (ia >= g.nxp) return;
  const int n = a.ndxr, ldf = a.ldf, ld = g.ld;
  const int ioff = ia * n, joff = ja * n;
  a.tauxa[(size_t)ja * ld + ia] = a.taux[(size_t)joff * ldf + ioff];
  a.tauya[(size_t)ja * ld + ia] = a.tauy[(size_t)joff * ldf + ioff];
  if (ia < g.nxt) {   // vekat(nxta, nypa): taux along the southern side of cell (ia, ja)
    const double *t = a.taux + (size_t)joff * ldf + ioff;
    double tausum = 0.5 * t[0];
    for (int i = 1; i <= n - 1; ++i) tausum = tausum + t[i];
    tausum = tausum + 0.5 * t[n];
    a.vek[(size_t)ja * ld + ia] = a.uvekfc * tausum;
  }
  if (ja < g.nyt) {   // uekat(nxpa, nyta): tauy along the western side; column nxpa repeats column 1
    const int is = (ia == g.nxp - 1) ? 0 : ioff;
    const double *t = a.tauy + (size_t)joff * ldf + is;
    double tausum = 0.5 * t[0];
    for (int j = 1; j <= n - 1; ++j) tausum = tausum + t[(size_t)j * ldf];
    tausum = tausum + 0.5 * t[(size_t)n * ldf];
    a.uek[(size_t)ja * ld + ia] = -a.uvekfc * tausum;
  }
}

// wekta = -hmat*(d/dx uekat + d/dy vekat) (src/xfosubs.F:411-416)
__global__ void __launch_bounds__(256) k_xf_wekta(XfArgsK a) {
  const Grid &g = a.ga;
  const int ia = blockIdx.x * blockDim.x + threadIdx.x, ja = blockIdx.y;
  if (ia >= g.nxt) return;
  const size_t c = (size_t)ja * g.ld + ia;
  a.wekta[c] = -a.hmrdxa * (a.uek[c + 1] - a.uek[c] + a.vek[c + g.ld] - a.vek[c]);
}

// wekpa: weighted box average of the fine-grid Ekman pumping around each atmosphere p point
// (src/xfosubs.F:425-471); wektaor is evaluated on the fly.  One warp per p point, lanes
// stride the box in x, rows in order; fixed shuffle tree => deterministic
__global__ void __launch_bounds__(256) k_xf_wekpa(XfArgsK a) {
  const Grid &g = a.ga;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= g.nxp * g.nyp) return;
  const int ja = warp / g.nxp, ia = warp - ja * g.nxp;
  const int n = a.ndxr, ldf = a.ldf;
  const int nxtf = a.nxf - 1, nytf = a.nyf - 1;
  const bool odd = n & 1;
  const int nij = n + (n & 1);
  const int jbeg = ja * n - (n - 1) / 2, ibeg = ia * n - (n - 1) / 2;     // 1-based T subscripts of the box start
  const int jlo = max(1, jbeg), jhi = min(jbeg + nij - 1, nytf);
  double wsum = 0.0, wtasum = 0.0;
  // the box is walked as one flat list (row-major), so that all 32 lanes work whatever ndxr is
  // (with ndxr = 16 a lane-per-column walk would leave half the warp idle)
  const int nrows = jhi - jlo + 1;
  for (int e = lane; e < nrows * nij; e += 32) {
    const int jr = e / nij, di = e - jr * nij;
    const int j = jlo + jr, dj = j - jbeg;
    const double wtj = odd ? ((dj == 0 || dj == n) ? 0.5 : 1.0) : ((dj == n) ? 0.0 : 1.0);
    const double wti = odd ? ((di == 0 || di == n) ? 0.5 : 1.0) : ((di == n) ? 0.0 : 1.0);
    const int it = (ibeg + di - 1 + nxtf) % nxtf;     // 0-based fine T column
    const size_t r = (size_t)(j - 1) * ldf + it, rn = r + ldf;
    const double w = a.hxofac * (a.tauy[r + 1] + a.tauy[rn + 1] - (a.tauy[r] + a.tauy[rn]) + a.taux[r] + a.taux[r + 1] -
                                 (a.taux[rn] + a.taux[rn + 1]));
    wsum += wti * wtj;
    wtasum += wti * wtj * w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wsum += __shfl_down_sync(0xffffffffu, wsum, o);
    wtasum += __shfl_down_sync(0xffffffffu, wtasum, o);
  }
  if (lane == 0) a.wekpa[(size_t)ja * g.ld + ia] = wtasum / wsum;
}

// txisat, txinat: stress line integrals half a coarse cell inside the walls
// (src/xfosubs.F:493-517); one block
__global__ void __launch_bounds__(256) k_xf_txis(XfArgsK a) {
  __shared__ double red[2][8];
  const int n = a.ndxr, nxf = a.nxf, ldf = a.ldf;
  const bool odd = n & 1;
  const int jsou = n / 2, jnor = a.nyf - 1 - n / 2;     // 0-based rows
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < nxf; i += 256) {
    const double w = (i == 0 || i == nxf - 1) ? 0.5 : 1.0;
    if (odd) {
      s += w * (a.taux[(size_t)jsou * ldf + i] + a.taux[(size_t)(jsou + 1) * ldf + i]);
      q += w * (a.taux[(size_t)jnor * ldf + i] + a.taux[(size_t)(jnor - 1) * ldf + i]);
    } else {
      s += w * a.taux[(size_t)jsou * ldf + i];
      q += w * a.taux[(size_t)jnor * ldf + i];
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); q += __shfl_down_sync(0xffffffffu, q, o); }
  if (lane == 0) { red[0][w] = s; red[1][w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < 8; ++k) { t0 += red[0][k]; t1 += red[1][k]; }
    const double f = odd ? 0.5 * a.dxo : a.dxo;
    a.sc->txisat = f * t0;
    a.sc->txinat = f * t1;
  }
}

// ---- diabatic forcing (src/xfosubs.F:711-853) ----
struct FnArgs {
  Grid ga, go;
  int ndxr, nx1, ny1, nxaooc, nyaooc;   // nx1, ny1 0-based
  double Dmup, Dmdown, D0up, xlamda, ocfrac, fmafac, fmatop, hmafac, hmat;
  const double *astm, *sstm, *pam, *dtop, *hmm;
  const int *iam, *iap, *jam, *jap;
  const double *wpx, *wmx, *wpy, *wmy, *fsp_o, *fsp_a;
  double *fnetat, *fnetoc;
  double *part;      // [4][npart]: land astm, slhf, ocnrad, atmrad(into ocean)
  int npart;
  qgcm_scalars *sc;
};

// land value of fnetat and the land sum of astm; one block per atmosphere T row
__global__ void __launch_bounds__(128) k_xf_fnet_land(FnArgs a) {
  __shared__ double red[4];
  const Grid &g = a.ga;
  const int ja = blockIdx.x;
  const bool jin = ja >= a.ny1 && ja < a.ny1 + a.nyaooc;
  double s = 0.0;
  for (int ia = threadIdx.x; ia < g.nxt; ia += 128) {
    const size_t c = (size_t)ja * g.ld + ia;
    const bool ocean = jin && ia >= a.nx1 && ia < a.nx1 + a.nxaooc;
    const double t = a.astm[c];
    a.fnetat[c] = ocean ? 0.0 : (-a.fsp_a[ja] - a.Dmup * t);
    if (!ocean) s += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) a.part[ja] = red[0] + red[1] + red[2] + red[3];
}

// one block per atmosphere cell over the ocean: bilinear astm at the ndxr x ndxr ocean T
// cells below it, fnetoc there, and the cell's share of the ocean-atmosphere exchange
__global__ void __launch_bounds__(256) k_xf_fnet_ocean(FnArgs a) {
  __shared__ double red[4][8];
  const Grid &ga = a.ga, &go = a.go;
  const int ca = blockIdx.x, cb = blockIdx.y;         // cell within the ocean window
  const int n = a.ndxr;
  double s_at = 0.0, s_sl = 0.0, s_or = 0.0, s_ar = 0.0;
  for (int e = threadIdx.x; e < n * n; e += 256) {
    const int dj = e / n, di = e - dj * n;
    const int io = ca * n + di, jo = cb * n + dj;
    const int jm = a.jam[jo], jp = a.jap[jo];
    const double wmy = a.wmy[jo], wpy = a.wpy[jo], wmx = a.wmx[io], wpx = a.wpx[io];
    const int im = a.iam[io], ip = a.iap[io];
    const double asto = wmx * wmy * a.astm[(size_t)jm * ga.ld + im] + wpx * wmy * a.astm[(size_t)jm * ga.ld + ip] +
                        wmx * wpy * a.astm[(size_t)jp * ga.ld + im] + wpx * wpy * a.astm[(size_t)jp * ga.ld + ip];
    const size_t c = (size_t)jo * go.ld + io;
    const double sst = a.sstm[c];
    const double ocnrad = a.D0up * sst;
    const double slhf = a.xlamda * (sst - asto);
    const double atmrad = a.Dmdown * asto;
    a.fnetoc[c] = -a.fsp_o[jo] - atmrad - ocnrad - slhf;
    s_ar += atmrad;
    s_at += a.ocfrac * (ocnrad + (a.Dmdown - a.Dmup) * asto + slhf);
    s_sl += slhf;
    s_or += ocnrad;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_at += __shfl_down_sync(0xffffffffu, s_at, o);
    s_sl += __shfl_down_sync(0xffffffffu, s_sl, o);
    s_or += __shfl_down_sync(0xffffffffu, s_or, o);
    s_ar += __shfl_down_sync(0xffffffffu, s_ar, o);
  }
  if (lane == 0) { red[0][w] = s_at; red[1][w] = s_sl; red[2][w] = s_or; red[3][w] = s_ar; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0, 0, 0, 0};
    for (int q = 0; q < 4; ++q)
      for (int k = 0; k < 8; ++k) t[q] += red[q][k];
    const int b = cb * gridDim.x + ca;
    a.fnetat[(size_t)(a.ny1 + cb) * ga.ld + a.nx1 + ca] = t[0];   // land pass left 0 here
    a.part[a.npart + b] = t[1];
    a.part[2 * a.npart + b] = t[2];
    a.part[3 * a.npart + b] = t[3];
  }
}

// eta, topography and mixed-layer thickness terms of fnetat (src/xfosubs.F:833-844), and the
// four monitor averages (:765, :850-852)
__global__ void __launch_bounds__(256) k_xf_fnet_finish(FnArgs a, int nocean_blocks) {
  __shared__ double red[8];
  const Grid &g = a.ga;
  for (int ja = blockIdx.x; ja < g.nyt; ja += gridDim.x)
    for (int ia = threadIdx.x; ia < g.nxt; ia += 256) {
      const size_t c = (size_t)ja * g.ld + ia, cn = c + g.ld;
      const double *p1 = a.pam, *p2 = a.pam + g.lsz;
      a.fnetat[c] = a.fnetat[c] -
                    a.fmafac * (p1[c] - p2[c] + p1[c + 1] - p2[c + 1] + p1[cn] - p2[cn] + p1[cn + 1] - p2[cn + 1]) -
                    a.fmatop * (a.dtop[c] + a.dtop[c + 1] + a.dtop[cn] + a.dtop[cn + 1]) + a.hmafac * (a.hmm[c] - a.hmat);
    }
  if (blockIdx.x != 0) return;
  const double land = block256_range_sum(a.part, 0, g.nyt, red);
  const double slhf = block256_range_sum(a.part + a.npart, 0, nocean_blocks, red);
  const double orad = block256_range_sum(a.part + 2 * a.npart, 0, nocean_blocks, red);
  const double arad = block256_range_sum(a.part + 3 * a.npart, 0, nocean_blocks, red);
  if (threadIdx.x == 0) {
    const int natlan = g.nxt * g.nyt - a.nxaooc * a.nyaooc;
    a.sc->arlaav = natlan == 0 ? 0.0 : a.Dmup * land / (double)natlan;
    a.sc->slhfav = slhf * a.go.norm;
    a.sc->oradav = orad * a.go.norm;
    a.sc->arocav = arad * a.go.norm;
  }
}

// ---- host: bicubic weight tables (bcuini + wts2bb, src/xfosubs.F:1238-1728) ----
namespace {
struct W4 {   // weights indexed (id, jd, ip, jp), id/jd in -1..2
  double v[4][4][2][2];
  W4() { std::memset(v, 0, sizeof(v)); }
  double &at(int id, int jd, int ip, int jp) { return v[id + 1][jd + 1][ip][jp]; }
};

// rows of the inverse bicubic basis (the 16x16 integer matrix of src/xfosubs.F:1650-1667,
// stored here by row: coefficient i = sum_j SINV[i][j] * {f, dx fx, dy fy, dxdy fxy}_j)
const signed char SINV[16][16] = {
    {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},     {0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {-3, 3, 0, 0, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},  {2, -2, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0},     {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, -2, -1, 0, 0},  {0, 0, 0, 0, 0, 0, 0, 0, 2, -2, 0, 0, 1, 1, 0, 0},
    {-3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0, 0, 0, 0, 0},  {0, 0, 0, 0, -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0},
    {9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1}, {-6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1},
    {2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0},    {0, 0, 0, 0, 2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0},
    {-6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1}, {4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1}};

void bicubic_variant(int variant, double bcdy, int n, double *out /* [(n+1)^2][16] */) {
  W4 fcn, fnx, fny, fxy;
  for (int jp = 0; jp <= 1; ++jp)
    for (int ip = 0; ip <= 1; ++ip) {
      fcn.at(ip, jp, ip, jp) = 1.0;
      fnx.at(ip + 1, jp, ip, jp) = 0.5;
      fnx.at(ip - 1, jp, ip, jp) = -0.5;
      const bool wall = ((variant == 1 || variant == 2) && jp == 0) || ((variant == 3 || variant == 4) && jp == 1);
      const int out_j = (jp == 0) ? jp - 1 : jp + 1;        // the padded data row outside the wall
      const double sg = (jp == 0) ? 1.0 : -1.0;
      if (wall && (variant == 1 || variant == 3)) {          // u: mixed pressure condition
        fny.at(ip, jp, ip, jp) = sg * bcdy * fcn.at(ip, jp, ip, jp);
        fxy.at(ip + 1, jp, ip, jp) = sg * bcdy * fnx.at(ip + 1, jp, ip, jp);
        fxy.at(ip - 1, jp, ip, jp) = sg * bcdy * fnx.at(ip - 1, jp, ip, jp);
      } else if (wall) {                                      // v: vy = -ux from the padded u row
        fny.at(ip + 1, out_j, ip, jp) = -fnx.at(ip + 1, jp, ip, jp);
        fny.at(ip - 1, out_j, ip, jp) = -fnx.at(ip - 1, jp, ip, jp);
        fxy.at(ip + 1, out_j, ip, jp) = -1.0;
        fxy.at(ip, out_j, ip, jp) = 2.0;
        fxy.at(ip - 1, out_j, ip, jp) = -1.0;
      } else {                                                // centred differences
        fny.at(ip, jp + 1, ip, jp) = 0.5;
        fny.at(ip, jp - 1, ip, jp) = -0.5;
        fxy.at(ip + 1, jp + 1, ip, jp) = 0.25;
        fxy.at(ip - 1, jp + 1, ip, jp) = -0.25;
        fxy.at(ip + 1, jp - 1, ip, jp) = -0.25;
        fxy.at(ip - 1, jp - 1, ip, jp) = 0.25;
      }
    }
  // u2f(kp + 4*d, kd): data -> {f, dx fx, dy fy, dxdy fxy} at the four vertices
  double u2f[16][16], bmat[16][16];
  int kp = 0;
  for (int jp = 0; jp <= 1; ++jp)
    for (int ip = 0; ip <= 1; ++ip, ++kp) {
      int kd = 0;
      for (int jd = -1; jd <= 2; ++jd)
        for (int id = -1; id <= 2; ++id, ++kd) {
          u2f[kp][kd] = fcn.at(id, jd, ip, jp);
          u2f[kp + 4][kd] = fnx.at(id, jd, ip, jp);
          u2f[kp + 8][kd] = fny.at(id, jd, ip, jp);
          u2f[kp + 12][kd] = fxy.at(id, jd, ip, jp);
        }
    }
  for (int kd = 0; kd < 16; ++kd)
    for (int i = 0; i < 16; ++i) {
      double s = 0.0;
      for (int j = 0; j < 16; ++j) s = s + (double)SINV[i][j] * u2f[j][kd];
      bmat[i][kd] = s;
    }
  // weights at the fine points: sum_m bmat(m,k) * ss^i tt^j; only ii, jj < n are defined,
  // the rest stays zero as in the reference's static storage (src/xfosubs.F:1542 vs :1189)
  for (int jj = 0; jj < n; ++jj)
    for (int ii = 0; ii < n; ++ii) {
      const double ss = (double)ii / (double)n, tt = (double)jj / (double)n;
      double st[16];
      int mm = 0;
      for (int j = 0; j <= 3; ++j)
        for (int i = 0; i <= 3; ++i) st[mm++] = std::pow(ss, i) * std::pow(tt, j);
      for (int k = 0; k < 16; ++k) {
        double s = 0.0;
        for (int q = 0; q < 16; ++q) s = s + bmat[q][k] * st[q];
        out[(size_t)k * (n + 1) * (n + 1) + (ii + (n + 1) * jj)] = s;
      }
    }
}
}  // namespace

static void xf_plan(qgcm_model *m) {
  XfPlan &x = m->xf;
  const qgcm_config &c = m->cfg;
  const Grid &ga = m->ga, &go = m->go;
  const int n = c.ndxr;
  x.nxf = ga.nxt * n + 1; x.nyf = ga.nyt * n + 1; x.ldf = ((x.nxf + 15) / 16) * 16;
  x.u1 = (double *)dalloc(m, sizeof(double) * ga.lsz);
  x.v1 = (double *)dalloc(m, sizeof(double) * ga.lsz);
  x.taux = (double *)dalloc(m, sizeof(double) * (size_t)x.ldf * x.nyf);
  x.tauy = (double *)dalloc(m, sizeof(double) * (size_t)x.ldf * x.nyf);
  const size_t per = (size_t)(n + 1) * (n + 1) * 16;
  std::vector<double> stb(5 * per, 0.0);
  for (int v = 0; v < 5; ++v) bicubic_variant(v, c.bccoat / ga.dx, n, stb.data() + v * per);
  x.stb = (double *)dalloc(m, sizeof(double) * stb.size());
  QG_CUDA(cudaMemcpy(x.stb, stb.data(), sizeof(double) * stb.size(), cudaMemcpyHostToDevice));
  // bilint subscripts and weights (src/xfosubs.F:921-968), computed once on the host
  const double dxa = ga.dx, dxo = go.dx, dxainv = 1.0 / dxa;
  const int nxto = go.nxt, nyto = go.nyt, nxta = ga.nxt, nyta = ga.nyt;
  std::vector<int> iam(nxto), iap(nxto), jam(nyto), jap(nyto);
  std::vector<double> wpx(nxto), wmx(nxto), wpy(nyto), wmy(nyto), fso(nyto), fsa(nyta);
  const double xa1 = 0.5 * dxa, yla = nyta * dxa, PI = 3.14159265358979324;
  for (int io = 1; io <= nxto; ++io) {
    const double xo = ((io - 1) * dxo + (c.nx1 - 1) * dxa) + 0.5 * dxo;
    int im = (int)(1.0 + dxainv * (xo - xa1));
    const int ip = im + 1;
    const double xam = (im >= 1) ? ((im - 1) * dxa + 0.5 * dxa) : (xa1 - dxa);
    wpx[io - 1] = dxainv * (xo - xam);
    wmx[io - 1] = 1.0 - wpx[io - 1];
    iam[io - 1] = (im + nxta - 1) % nxta;      // 0-based
    iap[io - 1] = (ip + nxta - 1) % nxta;
  }
  for (int jo = 1; jo <= nyto; ++jo) {
    const double ypo = (c.ny1 - 1) * dxa + (jo - 1) * dxo;
    const double yo = ypo + 0.5 * dxo;
    int jm = (int)(1.0 + dxainv * (yo - 0.5 * dxa));
    int jp = jm + 1;
    jm = std::max(jm, 1);
    jp = std::min(jp, nyta);
    const double yam = (jm - 1) * dxa + 0.5 * dxa;
    wpy[jo - 1] = dxainv * (yo - yam);
    wmy[jo - 1] = 1.0 - wpy[jo - 1];
    jam[jo - 1] = jm - 1;
    jap[jo - 1] = jp - 1;
    fso[jo - 1] = c.fspco * 0.5 * std::sin(PI * (yo - 0.5 * yla) / yla);
  }
  for (int ja = 1; ja <= nyta; ++ja) {
    const double yta = (ja - 1) * dxa + 0.5 * dxa;
    fsa[ja - 1] = c.fspco * 0.5 * std::sin(PI * (yta - 0.5 * yla) / yla);
  }
  auto upi = [&](const std::vector<int> &v) {
    int *d = (int *)dalloc(m, sizeof(int) * v.size());
    QG_CUDA(cudaMemcpy(d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
    return d;
  };
  auto upd = [&](const std::vector<double> &v) {
    double *d = (double *)dalloc(m, sizeof(double) * v.size());
    QG_CUDA(cudaMemcpy(d, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice));
    return d;
  };
  x.iam = upi(iam); x.iap = upi(iap); x.jam = upi(jam); x.jap = upi(jap);
  x.wpx = upd(wpx); x.wmx = upd(wmx); x.wpy = upd(wpy); x.wmy = upd(wmy);
  x.fsp_o = upd(fso); x.fsp_a = upd(fsa);
  x.npart = std::max(nyta, (nxto / n) * (nyto / n)) + 8;
  x.part = (double *)dalloc(m, sizeof(double) * 4 * x.npart);
  x.ready = true;
}

void launch_xforc(qgcm_model *m) {
  if (m->ocean_only) {
    // ocean_only builds execute only the oceanic Ekman tail in the loop (src/xfosubs.F:568-709);
    // tauxo, tauyo, fnetoc are time-invariant inputs there
    launch_xforc_ocean_ekman(m);
    return;
  }
  if (m->atmos_only) throw std::runtime_error("qgcm_xforc: atmos_only decks are not supported");
  if (!m->xf.ready) xf_plan(m);
  XfPlan &x = m->xf;
  const qgcm_config &c = m->cfg;
  const Grid &ga = m->ga, &go = m->go;
  const int n = c.ndxr;
  XfArgsK a;
  a.ga = ga; a.go = go; a.ndxr = n; a.nxf = x.nxf; a.nyf = x.nyf; a.ldf = x.ldf;
  a.iocoff = (c.nx1 - 1) * n; a.jocoff = (c.ny1 - 1) * n;
  a.tau_udiff = m->tau_udiff; a.cyclic_oc = m->cyclic;
  a.hxafac = 0.5 * ga.rdxf0; a.hxofac = 0.5 * go.rdxf0;
  a.zbfcat = ga.rdxf0 / (0.5 * c.bccoat + 1.0);
  a.zbfcoc = go.rdxf0 / (0.5 * c.bccooc + 1.0);
  a.raoro = m->raoro;
  const double cdhfaa = (c.cdat / m->fnot) / c.hmat;
  const double cdhfab = (c.cdat / m->fnot) * (1.0 / c.hmat + m->raoro / c.hmoc);
  a.cdrfaa = c.cdat / std::fabs(cdhfaa); a.cdrfab = c.cdat / std::fabs(cdhfab);
  a.qu2faa = 4.0 * cdhfaa * cdhfaa; a.qu2fab = 4.0 * cdhfab * cdhfab;
  a.uvekfc = 1.0 / (c.hmat * m->fnot * (double)n);
  a.hmrdxa = c.hmat / ga.dx;
  a.dxo = go.dx;
  a.pam = m->F("pam"); a.pom = m->F("pom");
  a.u1 = x.u1; a.v1 = x.v1; a.taux = x.taux; a.tauy = x.tauy; a.stb = x.stb;
  a.tauxa = m->F("tauxa"); a.tauya = m->F("tauya"); a.uek = m->F("uekat"); a.vek = m->F("vekat");
  a.wekta = m->F("wekta"); a.wekpa = m->F("wekpa"); a.tauxo = m->F("tauxo"); a.tauyo = m->F("tauyo");
  a.sc = m->d_scal;
  QG_LAUNCH(m, "k_xf_wind", dim3((ga.nxp + 255) / 256, ga.nyp), 256, 0, k_xf_wind, a);
  if (ga.nyt >= 3 && env_int("QGCM_XF_SEP", 1)) {
    // interior coarse rows: separable kernel; the first and last coarse rows (2 ndxr + 1 fine rows) keep the
    // general one, whose weights carry the boundary conditions
    QG_LAUNCH(m, "k_xf_stress", dim3((x.nxf + 127) / 128, ga.nyt - 2), 128, 0, k_xf_stress_sep, a);
    QG_LAUNCH(m, "k_xf_stress_edge", dim3((x.nxf + 127) / 128, 2 * n + 1), 128, 0, k_xf_stress, a, 1);
  } else {
    QG_LAUNCH(m, "k_xf_stress", dim3((x.nxf + 127) / 128, x.nyf), 128, 0, k_xf_stress, a, 0);
  }
  QG_LAUNCH(m, "k_xf_sample", dim3((ga.nxp + 255) / 256, ga.nyp), 256, 0, k_xf_sample, a);
  QG_LAUNCH(m, "k_xf_wekta", dim3((ga.nxt + 255) / 256, ga.nyt), 256, 0, k_xf_wekta, a);
  QG_LAUNCH(m, "k_xf_wekpa", (ga.nxp * ga.nyp * 32 + 255) / 256, 256, 0, k_xf_wekpa, a);
  QG_LAUNCH(m, "k_xf_txis", 1, 256, 0, k_xf_txis, a);
  launch_xforc_ocean_ekman(m);
  FnArgs f;
  f.ga = ga; f.go = go; f.ndxr = n; f.nx1 = c.nx1 - 1; f.ny1 = c.ny1 - 1;
  f.nxaooc = go.nxt / n; f.nyaooc = go.nyt / n;
  f.Dmup = c.Dmup; f.Dmdown = c.Dmdown; f.D0up = c.D0up; f.xlamda = c.xlamda;
  f.ocfrac = go.dx * go.dx / (ga.dx * ga.dx);
  f.fmafac = c.Adown[0] * 0.25 / c.gpat[0];
  f.fmatop = 0.25 * (c.Cmup + c.C1down);
  f.hmafac = -c.hmadmp - c.Bmup - c.B1down;
  f.hmat = c.hmat;
  f.astm = m->F("astm"); f.sstm = m->F("sstm"); f.pam = m->F("pam"); f.dtop = m->F("dtopat"); f.hmm = m->F("hmixam");
  f.iam = x.iam; f.iap = x.iap; f.jam = x.jam; f.jap = x.jap;
  f.wpx = x.wpx; f.wmx = x.wmx; f.wpy = x.wpy; f.wmy = x.wmy; f.fsp_o = x.fsp_o; f.fsp_a = x.fsp_a;
  f.fnetat = m->F("fnetat"); f.fnetoc = m->F("fnetoc");
  f.part = x.part; f.npart = x.npart; f.sc = m->d_scal;
  QG_LAUNCH(m, "k_xf_fnet_land", ga.nyt, 128, 0, k_xf_fnet_land, f);
  QG_LAUNCH(m, "k_xf_fnet_ocean", dim3(f.nxaooc, f.nyaooc), 256, 0, k_xf_fnet_ocean, f);
  QG_LAUNCH(m, "k_xf_fnet_finish", std::min(ga.nyt, 64), 256, 0, k_xf_fnet_finish, f, f.nxaooc * f.nyaooc);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qg

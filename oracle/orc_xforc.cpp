// TEST INFRASTRUCTURE ONLY -- CPU oracle (pinned against the translated reference, see orc_model.h).
// Coupled forcing: restatement of src/xfosubs.F:52-858 (xforc), :862-887 (fsprim),
// :891-993 (bilint), :997-1234 (auvbcu), :1238-1621 (bcuini), :1625-1728 (wts2bb).
// Loop order and expression association follow the Fortran; 1-based accessor macros.
#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "orc_model.h"

namespace orc {

#define IX2(i, j, nx) ((size_t)((i)-1) + (size_t)(nx) * (size_t)((j)-1))
#define IX3(i, j, k, nx, ny) ((size_t)((i)-1) + (size_t)(nx) * ((size_t)((j)-1) + (size_t)(ny) * (size_t)((k)-1)))

// ---------------------------------------------------------------- src/xfosubs.F:1625-1728
// stinv: inverse of the bicubic basis matrix, DATA statement at src/xfosubs.F:1650-1667
// (column-major fill: each row below is one column of stinv)
static const double STINV_COLS[16][16] = {
    {1, 0, -3, 2, 0, 0, 0, 0, -3, 0, 9, -6, 2, 0, -6, 4},
    {0, 0, 3, -2, 0, 0, 0, 0, 0, 0, -9, 6, 0, 0, 6, -4},
    {0, 0, 0, 0, 0, 0, 0, 0, 3, 0, -9, 6, -2, 0, 6, -4},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 9, -6, 0, 0, -6, 4},
    {0, 1, -2, 1, 0, 0, 0, 0, 0, -3, 6, -3, 0, 2, -4, 2},
    {0, 0, -1, 1, 0, 0, 0, 0, 0, 0, 3, -3, 0, 0, -2, 2},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 3, -6, 3, 0, -2, 4, -2},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, 2, -2},
    {0, 0, 0, 0, 1, 0, -3, 2, -2, 0, 6, -4, 1, 0, -3, 2},
    {0, 0, 0, 0, 0, 0, 3, -2, 0, 0, -6, 4, 0, 0, 3, -2},
    {0, 0, 0, 0, 0, 0, 0, 0, -1, 0, 3, -2, 1, 0, -3, 2},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -3, 2, 0, 0, 3, -2},
    {0, 0, 0, 0, 0, 1, -2, 1, 0, -2, 4, -2, 0, 1, -2, 1},
    {0, 0, 0, 0, 0, 0, -1, 1, 0, 0, 2, -2, 0, 0, -1, 1},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 2, -1, 0, 1, -2, 1},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 0, 0, -1, 1}};

// weight arrays wf**(-1:2,-1:2,0:1,0:1)
struct Wf {
  double v[4][4][2][2];   // [id+1][jd+1][ip][jp]
  double &operator()(int id, int jd, int ip, int jp) { return v[id + 1][jd + 1][ip][jp]; }
  double operator()(int id, int jd, int ip, int jp) const { return v[id + 1][jd + 1][ip][jp]; }
  void zero() {
    for (auto &a : v) for (auto &b : a) for (auto &c : b) for (auto &d : c) d = 0.0;
  }
};

// bbb(16,16), column-major: bbb[(i-1) + 16*(kd-1)]
static void wts2bb(const Wf &wfcn, const Wf &wfnx, const Wf &wfny, const Wf &wfxy, double *bbb) {
  double u2f[16 * 16];
#define U2F(i, j) u2f[((i)-1) + 16 * ((j)-1)]
  int kp = 0;
  for (int jp = 0; jp <= 1; ++jp)
    for (int ip = 0; ip <= 1; ++ip) {
      kp = kp + 1;
      int kd = 0;
      for (int jd = -1; jd <= 2; ++jd)
        for (int id = -1; id <= 2; ++id) {
          kd = kd + 1;
          U2F(kp, kd) = wfcn(id, jd, ip, jp);
          U2F(kp + 4, kd) = wfnx(id, jd, ip, jp);
          U2F(kp + 8, kd) = wfny(id, jd, ip, jp);
          U2F(kp + 12, kd) = wfxy(id, jd, ip, jp);
        }
    }
  for (int kd = 1; kd <= 16; ++kd)
    for (int i = 1; i <= 16; ++i) {
      double wfsum = 0.0;
      for (int j = 1; j <= 16; ++j) wfsum = wfsum + STINV_COLS[j - 1][i - 1] * U2F(j, kd);
      bbb[(i - 1) + 16 * (kd - 1)] = wfsum;
    }
#undef U2F
}

// regular centred differences at vertex (ip,jp), src/xfosubs.F:1310-1318
static void wf_regular_y(Wf &wfny, Wf &wfxy, int ip, int jp) {
  wfny(ip, jp + 1, ip, jp) = 0.5;
  wfny(ip, jp - 1, ip, jp) = -0.5;
  wfxy(ip + 1, jp + 1, ip, jp) = 0.25;
  wfxy(ip - 1, jp + 1, ip, jp) = -0.25;
  wfxy(ip + 1, jp - 1, ip, jp) = -0.25;
  wfxy(ip - 1, jp - 1, ip, jp) = 0.25;
}

// ---------------------------------------------------------------- src/xfosubs.F:1238-1621
// stb**(16, 0:ndxr, 0:ndxr) stored as [(k-1) + 16*(ii + (ndxr+1)*jj)].  Only jj < ndxr and
// ii < ndxr are filled; the rest stays zero like the reference's static storage
// (auvbcu reads jj = ndxr of stbun/stbvn).
static void bcuini(Model &m) {
  const int n = m.ndxr;
  const size_t sz = (size_t)16 * (n + 1) * (n + 1);
  vec stfn(sz, 0.0);
#define ST(a, k, ii, jj) a[((k)-1) + (size_t)16 * ((ii) + (size_t)(n + 1) * (jj))]
  for (int jj = 0; jj <= n; ++jj) {
    const double tt = (double)jj / (double)n;
    for (int ii = 0; ii <= n; ++ii) {
      const double ss = (double)ii / (double)n;
      int mm = 0;
      for (int j = 0; j <= 3; ++j)
        for (int i = 0; i <= 3; ++i) {
          mm = mm + 1;
          ST(stfn, mm, ii, jj) = std::pow(ss, i) * std::pow(tt, j);
        }
    }
  }
  const double bcdy = m.c.bccoat / m.dya;
  Wf wfcn, wfnx, wfny, wfxy;
  double bmat[256];
  vec *dst[5] = {&m.stbbb, &m.stbus, &m.stbvs, &m.stbun, &m.stbvn};
  for (int variant = 0; variant < 5; ++variant) {
    wfcn.zero(); wfnx.zero(); wfny.zero(); wfxy.zero();
    for (int jp = 0; jp <= 1; ++jp)
      for (int ip = 0; ip <= 1; ++ip) {
        wfcn(ip, jp, ip, jp) = 1.0;
        wfnx(ip + 1, jp, ip, jp) = 0.5;
        wfnx(ip - 1, jp, ip, jp) = -0.5;
        const bool south_edge = (variant == 1 || variant == 2) && jp == 0;
        const bool north_edge = (variant == 3 || variant == 4) && jp == 1;
        if (variant == 1 && south_edge) {          // u, southern boundary: mixed pressure BC
          wfny(ip, jp, ip, jp) = bcdy * wfcn(ip, jp, ip, jp);
          wfxy(ip + 1, jp, ip, jp) = bcdy * wfnx(ip + 1, jp, ip, jp);
          wfxy(ip - 1, jp, ip, jp) = bcdy * wfnx(ip - 1, jp, ip, jp);
        } else if (variant == 2 && south_edge) {   // v, southern boundary: u values at jd = -1
          wfny(ip + 1, jp - 1, ip, jp) = -wfnx(ip + 1, jp, ip, jp);
          wfny(ip - 1, jp - 1, ip, jp) = -wfnx(ip - 1, jp, ip, jp);
          wfxy(ip + 1, jp - 1, ip, jp) = -1.0;
          wfxy(ip, jp - 1, ip, jp) = 2.0;
          wfxy(ip - 1, jp - 1, ip, jp) = -1.0;
        } else if (variant == 3 && north_edge) {   // u, northern boundary
          wfny(ip, jp, ip, jp) = -bcdy * wfcn(ip, jp, ip, jp);
          wfxy(ip + 1, jp, ip, jp) = -bcdy * wfnx(ip + 1, jp, ip, jp);
          wfxy(ip - 1, jp, ip, jp) = -bcdy * wfnx(ip - 1, jp, ip, jp);
        } else if (variant == 4 && north_edge) {   // v, northern boundary: u values at jd = 2
          wfny(ip + 1, jp + 1, ip, jp) = -wfnx(ip + 1, jp, ip, jp);
          wfny(ip - 1, jp + 1, ip, jp) = -wfnx(ip - 1, jp, ip, jp);
          wfxy(ip + 1, jp + 1, ip, jp) = -1.0;
          wfxy(ip, jp + 1, ip, jp) = 2.0;
          wfxy(ip - 1, jp + 1, ip, jp) = -1.0;
        } else {
          wf_regular_y(wfny, wfxy, ip, jp);
        }
      }
    wts2bb(wfcn, wfnx, wfny, wfxy, bmat);
    vec &out = *dst[variant];
    out.assign(sz, 0.0);
    for (int jj = 0; jj <= n - 1; ++jj)
      for (int ii = 0; ii <= n - 1; ++ii)
        for (int k = 1; k <= 16; ++k) {
          double stbsum = 0.0;
          for (int mm = 1; mm <= 16; ++mm) stbsum = stbsum + bmat[(mm - 1) + 16 * (k - 1)] * ST(stfn, mm, ii, jj);
          ST(out, k, ii, jj) = stbsum;
        }
  }
  m.bcu_ready = true;
#undef ST
}

// ---------------------------------------------------------------- src/xfosubs.F:997-1234
static void auvbcu(Model &m, const double *u1astd, vec &u1afin, const double *v1astd, vec &v1afin) {
  if (!m.bcu_ready) bcuini(m);
  const int n = m.ndxr, nxta = m.nxta, nyta = m.nyta, nxpa = m.nxpa, nypa = m.nypa;
  const int nxpaor = m.nxpaor, nypaor = m.nypaor;
#define US(i, j) u1astd[IX2(i, j, nxpa)]
#define VS(i, j) v1astd[IX2(i, j, nxpa)]
#define UF(i, j) u1afin[IX2(i, j, nxpaor)]
#define VF(i, j) v1afin[IX2(i, j, nxpaor)]
#define ST(a, k, ii, jj) a[((k)-1) + (size_t)16 * ((ii) + (size_t)(n + 1) * (jj))]
  for (int jc = 1; jc <= nyta; ++jc) {
    const int jfoff = 1 + (jc - 1) * n;
    const bool south = jc == 1, north = jc == nyta;
    for (int ic = 1; ic <= nxta; ++ic) {
      const int ifoff = 1 + (ic - 1) * n;
      const int icm1 = 1 + (ic - 2 + nxta) % nxta;
      const int icp2 = 1 + (ic + 1) % nxta;
      const int ix[4] = {icm1, ic, ic + 1, icp2};
      double udat[17], vdat[17];
      for (int row = 0; row < 4; ++row) {
        const int jd = row - 1;
        for (int q = 0; q < 4; ++q) {
          const int k = 4 * row + q + 1;
          if (south && jd == -1) {
            udat[k] = 0.0;
            vdat[k] = US(ix[q], 1);
          } else if (north && jd == 2) {
            udat[k] = 0.0;
            vdat[k] = US(ix[q], nypa);
          } else {
            udat[k] = US(ix[q], jc + jd);
            vdat[k] = VS(ix[q], jc + jd);
          }
        }
      }
      const vec &wu = south ? m.stbus : (north ? m.stbun : m.stbbb);
      const vec &wv = south ? m.stbvs : (north ? m.stbvn : m.stbbb);
      const int jjmax = north ? n : n - 1;
      for (int jj = 0; jj <= jjmax; ++jj)
        for (int ii = 0; ii <= n - 1; ++ii) {
          double usum = 0.0, vsum = 0.0;
          for (int k = 1; k <= 16; ++k) {
            usum = usum + udat[k] * ST(wu, k, ii, jj);
            vsum = vsum + vdat[k] * ST(wv, k, ii, jj);
          }
          UF(ifoff + ii, jfoff + jj) = usum;
          VF(ifoff + ii, jfoff + jj) = vsum;
        }
    }
  }
  for (int jj = 1; jj <= nypaor; ++jj) {
    UF(nxpaor, jj) = UF(1, jj);
    VF(nxpaor, jj) = VF(1, jj);
  }
#undef US
#undef VS
#undef UF
#undef VF
#undef ST
}

// ---------------------------------------------------------------- src/xfosubs.F:862-887
static inline double fsprim(const Model &m, double yrel) {
  const double PI = 3.14159265358979324;
  return m.c.fspco * 0.5 * std::sin(PI * yrel / m.yla);
}

// ---------------------------------------------------------------- src/xfosubs.F:891-993
static void bilint(const Model &m, const double *xa, const double *ya, int nxat, int nyat, const double *atmos,
                   const double *xo, const double *yo, int nxoc, int nyoc, double *ocean, double fmult) {
  const double dxainv = 1.0 / m.dxa, dyainv = 1.0 / m.dya;
  std::vector<int> iam(nxoc), iap(nxoc);
  vec wpx(nxoc), wmx(nxoc);
  for (int io = 1; io <= nxoc; ++io) {
    int im = (int)(1.0 + dxainv * (xo[io - 1] - xa[0]));
    int ip = im + 1;
    double xam;
    if (im >= 1) xam = xa[im - 1];
    else xam = xa[0] - m.dxa;
    wpx[io - 1] = dxainv * (xo[io - 1] - xam);
    wmx[io - 1] = 1.0 - wpx[io - 1];
    iam[io - 1] = 1 + (im + nxat - 1) % nxat;
    iap[io - 1] = 1 + (ip + nxat - 1) % nxat;
  }
  for (int jo = 1; jo <= nyoc; ++jo) {
    int jam = (int)(1.0 + dyainv * (yo[jo - 1] - ya[0]));
    int jap = jam + 1;
    jam = std::max(jam, 1);
    jap = std::min(jap, nyat);
    const double wpy = dyainv * (yo[jo - 1] - ya[jam - 1]);
    const double wmy = 1.0 - wpy;
    for (int io = 1; io <= nxoc; ++io)
      ocean[IX2(io, jo, nxoc)] = fmult * (wmx[io - 1] * wmy * atmos[IX2(iam[io - 1], jam, nxat)] +
                                          wpx[io - 1] * wmy * atmos[IX2(iap[io - 1], jam, nxat)] +
                                          wmx[io - 1] * wpy * atmos[IX2(iam[io - 1], jap, nxat)] +
                                          wpx[io - 1] * wpy * atmos[IX2(iap[io - 1], jap, nxat)]);
  }
}

// ---------------------------------------------------------------- src/xfosubs.F:52-858
void Model::xforc() {
  if (ocean_only) {
    // ocean_only decks: tauxo/tauyo/fnetoc are time-invariant inputs and only the oceanic
    // Ekman tail executes in the loop (src/xfosubs.F:568-709)
    xforc_ocean_ekman();
    return;
  }
  const double hxafac = 0.5 * rdxaf0, hxofac = 0.5 * rdxof0;
  const double hmat = c.hmat, hmoc = c.hmoc, cdat = c.cdat;
  const int iocoff = (nx1 - 1) * ndxr, jocoff = (ny1 - 1) * ndxr;
  const int ipobeg = iocoff + 1, ipoend = ipobeg + nxto, jpobeg = jocoff + 1, jpoend = jpobeg + nyto;
  const int jsou = 1 + ndxr / 2, jnor = nypaor - ndxr / 2;
  const double uvekfc = 1.0 / (hmat * fnot * (double)ndxr);
  const double hmrdxa = hmat / dxa;
  const double cdhfaa = (cdat / fnot) / hmat;
  const double cdhfab = (cdat / fnot) * (1.0 / hmat + raoro / hmoc);
  const double cdrfaa = cdat / std::fabs(cdhfaa);
  const double cdrfab = cdat / std::fabs(cdhfab);
  const double qu2faa = 4.0 * cdhfaa * cdhfaa;
  const double qu2fab = 4.0 * cdhfab * cdhfab;
  const bool ndxodd = (ndxr % 2) == 1;
  const int nijwid = ndxr + ndxr % 2;
  vec wt(ndxr + 1, 1.0);
  if (ndxodd) {
    wt[0] = 0.5;
    wt[ndxr] = 0.5;
  } else {
    wt[ndxr] = 0.0;
  }
#define PAM(i, j, k) pam[IX3(i, j, k, nxpa, nypa)]
#define POM(i, j, k) pom[IX3(i, j, k, nxpo, nypo)]
  // geostrophic velocity of atmosphere layer 1 at p points, src/xfosubs.F:186-214
  vec u1at((size_t)nxpa * nypa), v1at((size_t)nxpa * nypa);
#define U1(i, j) u1at[IX2(i, j, nxpa)]
#define V1(i, j) v1at[IX2(i, j, nxpa)]
  const double zbfcat = rdxaf0 / (0.5 * c.bccoat + 1.0);
  for (int i = 1; i <= nxpa; ++i) {
    U1(i, 1) = -zbfcat * (PAM(i, 2, 1) - PAM(i, 1, 1));
    V1(i, 1) = 0.0;
    U1(i, nypa) = -zbfcat * (PAM(i, nypa, 1) - PAM(i, nypa - 1, 1));
    V1(i, nypa) = 0.0;
  }
  for (int j = 2; j <= nypa - 1; ++j) {
    U1(1, j) = -hxafac * (PAM(1, j + 1, 1) - PAM(1, j - 1, 1));
    V1(1, j) = hxafac * (PAM(2, j, 1) - PAM(nxpa - 1, j, 1));
    for (int i = 2; i <= nxpa - 1; ++i) {
      U1(i, j) = -hxafac * (PAM(i, j + 1, 1) - PAM(i, j - 1, 1));
      V1(i, j) = hxafac * (PAM(i + 1, j, 1) - PAM(i - 1, j, 1));
    }
    U1(nxpa, j) = U1(1, j);
    V1(nxpa, j) = V1(1, j);
  }
  const size_t nfine = (size_t)nxpaor * nypaor;
  vec &u1ator = work(wk_u1ator, nfine), &v1ator = work(wk_v1ator, nfine);
  vec &tauxaor = work(wk_tauxaor, nfine), &tauyaor = work(wk_tauyaor, nfine);
#pragma omp parallel for schedule(static)
  for (size_t n = 0; n < nfine; ++n) u1ator[n] = v1ator[n] = 0.0;
  auvbcu(*this, u1at.data(), u1ator, v1at.data(), v1ator);
#define UF(i, j) u1ator[IX2(i, j, nxpaor)]
#define VF(i, j) v1ator[IX2(i, j, nxpaor)]
#define TXF(i, j) tauxaor[IX2(i, j, nxpaor)]
#define TYF(i, j) tauyaor[IX2(i, j, nxpaor)]
  if (!atmos_only && tau_udiff) {
    // subtract the ocean layer-1 geostrophic velocity, src/xfosubs.F:250-300
    const double zbfcoc = rdxof0 / (0.5 * c.bccooc + 1.0);
    for (int i = 1; i <= nxpo; ++i) {
      double u1oc = -zbfcoc * (POM(i, 2, 1) - POM(i, 1, 1));
      UF(iocoff + i, jocoff + 1) = UF(iocoff + i, jocoff + 1) - u1oc;
      u1oc = -zbfcoc * (POM(i, nypo, 1) - POM(i, nypo - 1, 1));
      UF(iocoff + i, jocoff + nypo) = UF(iocoff + i, jocoff + nypo) - u1oc;
    }
    for (int j = 2; j <= nypo - 1; ++j) {
      double u1oc, v1oc;
      if (cyclic) {
        u1oc = -hxofac * (POM(1, j + 1, 1) - POM(1, j - 1, 1));
        v1oc = hxofac * (POM(2, j, 1) - POM(nxpo - 1, j, 1));
      } else {
        u1oc = 0.0;
        v1oc = zbfcoc * (POM(2, j, 1) - POM(1, j, 1));
      }
      UF(iocoff + 1, jocoff + j) = UF(iocoff + 1, jocoff + j) - u1oc;
      VF(iocoff + 1, jocoff + j) = VF(iocoff + 1, jocoff + j) - v1oc;
      for (int i = 2; i <= nxpo - 1; ++i) {
        u1oc = -hxofac * (POM(i, j + 1, 1) - POM(i, j - 1, 1));
        v1oc = hxofac * (POM(i + 1, j, 1) - POM(i - 1, j, 1));
        UF(iocoff + i, jocoff + j) = UF(iocoff + i, jocoff + j) - u1oc;
        VF(iocoff + i, jocoff + j) = VF(iocoff + i, jocoff + j) - v1oc;
      }
      if (cyclic) {
        u1oc = -hxofac * (POM(nxpo, j + 1, 1) - POM(nxpo, j - 1, 1));
        v1oc = hxofac * (POM(2, j, 1) - POM(nxpo - 1, j, 1));
      } else {
        u1oc = 0.0;
        v1oc = zbfcoc * (POM(nxpo, j, 1) - POM(nxpo - 1, j, 1));
      }
      UF(iocoff + nxpo, jocoff + j) = UF(iocoff + nxpo, jocoff + j) - u1oc;
      VF(iocoff + nxpo, jocoff + j) = VF(iocoff + nxpo, jocoff + j) - v1oc;
    }
  }
  // quadratic drag law, src/xfosubs.F:319-354
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= nypaor; ++j) {
    for (int i = 1; i <= nxpaor; ++i) {
      double cdrfac = cdrfaa, qu2fac = qu2faa;
      if (tau_udiff && j >= jpobeg && j <= jpoend && i >= ipobeg && i <= ipoend) {
        cdrfac = cdrfab;
        qu2fac = qu2fab;
      }
      const double delu1 = UF(i, j), delv1 = VF(i, j);
      const double scasqd = -0.5 + 0.5 * std::sqrt(1.0 + qu2fac * (delu1 * delu1 + delv1 * delv1));
      const double scashr = std::sqrt(scasqd);
      const double cdochi = cdrfac * scashr / (1.0 + scasqd);
      TXF(i, j) = cdochi * (delu1 - scashr * delv1);
      TYF(i, j) = cdochi * (delv1 + scashr * delu1);
    }
  }
  // sample to the standard atmosphere grid, src/xfosubs.F:362-367
  for (int ja = 1; ja <= nypa; ++ja)
    for (int ia = 1; ia <= nxpa; ++ia) {
      tauxa[IX2(ia, ja, nxpa)] = TXF(1 + (ia - 1) * ndxr, 1 + (ja - 1) * ndxr);
      tauya[IX2(ia, ja, nxpa)] = TYF(1 + (ia - 1) * ndxr, 1 + (ja - 1) * ndxr);
    }
  // horizontal Ekman velocities, src/xfosubs.F:377-416
#define UE(i, j) uekat[IX2(i, j, nxpa)]
#define VE(i, j) vekat[IX2(i, j, nxta)]
  for (int ja = 1; ja <= nypa; ++ja) {
    const int joff = 1 + (ja - 1) * ndxr;
    for (int ia = 1; ia <= nxta; ++ia) {
      const int ioff = 1 + (ia - 1) * ndxr;
      double tausum = 0.5 * TXF(ioff, joff);
      for (int i = 1; i <= ndxr - 1; ++i) tausum = tausum + TXF(ioff + i, joff);
      tausum = tausum + 0.5 * TXF(ioff + ndxr, joff);
      VE(ia, ja) = uvekfc * tausum;
    }
  }
  for (int ja = 1; ja <= nyta; ++ja) {
    const int joff = 1 + (ja - 1) * ndxr;
    for (int ia = 1; ia <= nxta; ++ia) {
      const int ioff = 1 + (ia - 1) * ndxr;
      double tausum = 0.5 * TYF(ioff, joff);
      for (int j = 1; j <= ndxr - 1; ++j) tausum = tausum + TYF(ioff, joff + j);
      tausum = tausum + 0.5 * TYF(ioff, joff + ndxr);
      UE(ia, ja) = -uvekfc * tausum;
    }
    UE(nxpa, ja) = UE(1, ja);
    for (int ia = 1; ia <= nxta; ++ia)
      wekta[IX2(ia, ja, nxta)] = -hmrdxa * (UE(ia + 1, ja) - UE(ia, ja) + VE(ia, ja + 1) - VE(ia, ja));
  }
  // Ekman pumping at ocean resolution on atmosphere T points, src/xfosubs.F:425-432
  vec &wektaor = work(wk_wektaor, (size_t)nxtaor * nytaor);
#define WTF(i, j) wektaor[IX2(i, j, nxtaor)]
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= nytaor; ++j)
    for (int i = 1; i <= nxtaor; ++i)
      WTF(i, j) = hxofac * (TYF(i + 1, j) + TYF(i + 1, j + 1) - (TYF(i, j) + TYF(i, j + 1)) + TXF(i, j) + TXF(i + 1, j) -
                            (TXF(i, j + 1) + TXF(i + 1, j + 1)));
  // box average to the coarse p grid, src/xfosubs.F:446-471
  for (int ja = 1; ja <= nypa; ++ja) {
    const int jbeg = (ja - 1) * ndxr - (ndxr - 1) / 2;
    const int jlo = std::max(1, jbeg);
    const int jhi = std::min(jbeg + nijwid - 1, nytaor);
    for (int ia = 1; ia <= nxpa; ++ia) {
      const int ibeg = (ia - 1) * ndxr - (ndxr - 1) / 2;
      double wsum = 0.0, wtasum = 0.0;
      for (int j = jlo; j <= jhi; ++j) {
        const double wtj = wt[j - jbeg];
        for (int i = ibeg; i <= ibeg + nijwid - 1; ++i) {
          const int it = 1 + (i - 1 + nxtaor) % nxtaor;
          wsum = wsum + wt[i - ibeg] * wtj;
          wtasum = wtasum + wt[i - ibeg] * wtj * WTF(it, j);
        }
      }
      wekpa[IX2(ia, ja, nxpa)] = wtasum / wsum;
    }
  }
  // momentum-constraint line integrals, src/xfosubs.F:493-517
  {
    double txsums, txsumn;
    if (ndxodd) {
      txsums = 0.5 * (TXF(1, jsou) + TXF(1, jsou + 1));
      txsumn = 0.5 * (TXF(1, jnor) + TXF(1, jnor - 1));
      for (int i = 2; i <= nxpaor - 1; ++i) {
        txsums = txsums + (TXF(i, jsou) + TXF(i, jsou + 1));
        txsumn = txsumn + (TXF(i, jnor) + TXF(i, jnor - 1));
      }
      txsums = txsums + 0.5 * (TXF(nxpaor, jsou) + TXF(nxpaor, jsou + 1));
      txsumn = txsumn + 0.5 * (TXF(nxpaor, jnor) + TXF(nxpaor, jnor - 1));
      s.txisat = 0.5 * dxo * txsums;
      s.txinat = 0.5 * dxo * txsumn;
    } else {
      txsums = 0.5 * TXF(1, jsou);
      txsumn = 0.5 * TXF(1, jnor);
      for (int i = 2; i <= nxpaor - 1; ++i) {
        txsums = txsums + TXF(i, jsou);
        txsumn = txsumn + TXF(i, jnor);
      }
      txsums = txsums + 0.5 * TXF(nxpaor, jsou);
      txsumn = txsumn + 0.5 * TXF(nxpaor, jnor);
      s.txisat = dxo * txsums;
      s.txinat = dxo * txsumn;
    }
  }
  if (!atmos_only) {
    // oceanic stresses, src/xfosubs.F:554-559, then the Ekman tail :568-709
    for (int jo = 1; jo <= nypo; ++jo)
      for (int io = 1; io <= nxpo; ++io) {
        tauxo[IX2(io, jo, nxpo)] = raoro * TXF(iocoff + io, jocoff + jo);
        tauyo[IX2(io, jo, nxpo)] = raoro * TYF(iocoff + io, jocoff + jo);
      }
    xforc_ocean_ekman();
  }
  // ---- diabatic forcing, src/xfosubs.F:711-853 ----
  vec xta(nxta), xto(nxto);
  vec &asto = work(wk_asto, (size_t)nxto * nyto);
  for (int i = 1; i <= nxta; ++i) xta[i - 1] = (i - 1) * dxa + 0.5 * dxa;
  for (int i = 1; i <= nxto; ++i) xto[i - 1] = ((i - 1) * dxo + (nx1 - 1) * dxa) + 0.5 * dxo;
  bilint(*this, xta.data(), yta.data(), nxta, nyta, astm.data(), xto.data(), yto.data(), nxto, nyto, asto.data(), 1.0);
  double arlasm = 0.0;
  for (int ja = 1; ja <= nyta; ++ja) {
    const double fsp = fsprim(*this, ytarel[ja - 1]);
    for (int ia = 1; ia <= nxta; ++ia) {
      fnetat[IX2(ia, ja, nxta)] = -fsp - c.Dmup * astm[IX2(ia, ja, nxta)];
      arlasm = arlasm + astm[IX2(ia, ja, nxta)];
    }
  }
  const int nxaooc = nxto / ndxr, nyaooc = nyto / ndxr;
  int natocn = 0;
  for (int ja = ny1; ja <= ny1 + nyaooc - 1; ++ja)
    for (int ia = nx1; ia <= nx1 + nxaooc - 1; ++ia) {
      fnetat[IX2(ia, ja, nxta)] = 0.0;
      arlasm = arlasm - astm[IX2(ia, ja, nxta)];
      natocn = natocn + 1;
    }
  const int natlan = nxta * nyta - natocn;
  s.arlaav = natlan == 0 ? 0.0 : c.Dmup * arlasm / (double)natlan;
  const double ocfrac = dxo * dyo / (dxa * dya);
  const double fmafac = c.Adown[0] * 0.25 / c.gpat[0];
  const double fmatop = 0.25 * (c.Cmup + c.C1down);
  const double hmafac = -c.hmadmp - c.Bmup - c.B1down;
  double slhfsm = 0.0, oradsm = 0.0, arocsm = 0.0;
  for (int jo = 1; jo <= nyto; ++jo) {
    const int ja = ny1 + (jo - 1) / ndxr;
    const double fsp = fsprim(*this, ytorel[jo - 1]);
    for (int io = 1; io <= nxto; ++io) {
      const int ia = nx1 + (io - 1) / ndxr;
      const double sstv = sstm[IX2(io, jo, nxto)], astv = asto[IX2(io, jo, nxto)];
      const double ocnrad = c.D0up * sstv;
      const double slhf = c.xlamda * (sstv - astv);
      double atmrad;
      if (!atmos_only) {
        atmrad = c.Dmdown * astv;
        fnetoc[IX2(io, jo, nxto)] = -fsp - atmrad - ocnrad - slhf;
        arocsm = arocsm + atmrad;
      }
      atmrad = (c.Dmdown - c.Dmup) * astv;
      fnetat[IX2(ia, ja, nxta)] = fnetat[IX2(ia, ja, nxta)] + ocfrac * (ocnrad + atmrad + slhf);
      slhfsm = slhfsm + slhf;
      oradsm = oradsm + ocnrad;
    }
  }
  for (int j = 1; j <= nyta; ++j)
    for (int i = 1; i <= nxta; ++i)
      fnetat[IX2(i, j, nxta)] =
          fnetat[IX2(i, j, nxta)] -
          fmafac * (PAM(i, j, 1) - PAM(i, j, 2) + PAM(i + 1, j, 1) - PAM(i + 1, j, 2) + PAM(i, j + 1, 1) - PAM(i, j + 1, 2) +
                    PAM(i + 1, j + 1, 1) - PAM(i + 1, j + 1, 2)) -
          fmatop * (dtopat[IX2(i, j, nxpa)] + dtopat[IX2(i + 1, j, nxpa)] + dtopat[IX2(i, j + 1, nxpa)] +
                    dtopat[IX2(i + 1, j + 1, nxpa)]) +
          hmafac * (hmixam[IX2(i, j, nxta)] - hmat);
  s.slhfav = slhfsm * ocnorm;
  s.oradav = oradsm * ocnorm;
  s.arocav = arocsm * ocnorm;
}

}  // namespace orc

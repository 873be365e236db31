// Leapfrog vorticity step, shared by ocean (qgostep + ocadif, src/qgosubs.F:45-446) and
// atmosphere (qgastep + atadif, src/qgasubs.F:45-317).
//
// One fused kernel per layer-tile: del2(pom) -> del4 -> del6, the Arakawa 9-point
// Jacobian J(q,p), forcing, bottom drag and the leapfrog update, with the three
// intermediate Laplacians staged in shared memory (halo 3 of pom, halo 1 of po/qo)
// instead of round-tripping through HBM as the reference's del2p/d4p/dqdt arrays do.
// q(new) is written over qom in place (qom is only read at the centre point), and the
// host rotates the qo/qom pointers afterwards.
#include "qgcm_internal.h"

namespace qg {

constexpr int TX = 64, TY = 14;          // output tile (46 KB of static shared memory)
constexpr int PW = TX + 6, PH = TY + 6;  // pom tile (halo 3)
constexpr int DW2 = TX + 4, DH2 = TY + 4;  // del2 tile (halo 2)
constexpr int DW4 = TX + 2, DH4 = TY + 2;  // del4 / po / qo tiles (halo 1)

struct QgArgs {
  Grid g;
  int atmos;            // 0 ocean, 1 atmosphere (sign conventions of the forcing differ)
  double adfac;         // 1/(12 dx dy f0)
  double bcfac;         // mixed-BC factor bcco*dxm2/(0.5 bcco + 1)
  double f0;
  double fohfac[NLMAX];
  double ah2fac[NLMAX], ah4fac[NLMAX];   // ah2/f0, ah4/f0
  double bdrfac;        // ocean bottom drag
  const double *pm, *p, *q;   // lagged p, current p, current q   [nl][nyp][ld]
  double *qm;                 // lagged q in, new q out (in place)
  const double *wek, *ent;    // Ekman velocity and entrainment at p points
};

// canonical column for loads: periodic grids read column nxp as column 1
__device__ __forceinline__ int wrapx(int i, int nxp, int cyclic) {
  if (!cyclic) return i;
  const int per = nxp - 1;
  if (i < 0) i += per;
  if (i >= per) i -= per;
  return i;
}

__global__ void __launch_bounds__(256) k_qgstep(QgArgs a) {
  __shared__ double s_pm[PH][PW];
  __shared__ double s_d2[DH2][DW2];
  __shared__ double s_d4[DH4][DW4];
  __shared__ double s_p[DH4][DW4];
  __shared__ double s_q[DH4][DW4];
  const Grid &g = a.g;
  const int k = blockIdx.z;
  const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;   // 0-based origin of the output tile
  const size_t lo = (size_t)k * g.lsz;
  const double *pm = a.pm + lo, *p = a.p + lo, *q = a.q + lo;
  const int nxp = g.nxp, nyp = g.nyp, ld = g.ld, cyc = g.cyclic;
  const int tid = threadIdx.x;

  // ---- stage pom (halo 3), po and qo (halo 1); out-of-domain entries are never used
  for (int e = tid; e < PH * PW; e += 256) {
    const int ly = e / PW, lx = e - ly * PW;
    const int gj = j0 + ly - 3;
    int gi = i0 + lx - 3;
    double v = 0.0;
    if (gj >= 0 && gj < nyp) {
      gi = wrapx(gi, nxp, cyc);
      if (gi >= 0 && gi < nxp) v = pm[(size_t)gj * ld + gi];
    }
    s_pm[ly][lx] = v;
  }
  for (int e = tid; e < DH4 * DW4; e += 256) {
    const int ly = e / DW4, lx = e - ly * DW4;
    const int gj = j0 + ly - 1;
    int gi = i0 + lx - 1;
    double vp = 0.0, vq = 0.0;
    if (gj >= 0 && gj < nyp) {
      gi = wrapx(gi, nxp, cyc);
      if (gi >= 0 && gi < nxp) {
        vp = p[(size_t)gj * ld + gi];
        vq = q[(size_t)gj * ld + gi];
      }
    }
    s_p[ly][lx] = vp;
    s_q[ly][lx] = vq;
  }
  __syncthreads();

  // ---- del2 of pom on the halo-2 region (qgosubs.F:86-130 / qgasubs.F:74-100)
  for (int e = tid; e < DH2 * DW2; e += 256) {
    const int ly = e / DW2, lx = e - ly * DW2;
    const int gj = j0 + ly - 2, gi = i0 + lx - 2;
    const int py = ly + 1, px = lx + 1;   // position in s_pm
    double v = 0.0;
    if (gj >= 0 && gj < nyp && (cyc || (gi >= 0 && gi < nxp))) {
      if (gj == 0)
        v = a.bcfac * (s_pm[py + 1][px] - s_pm[py][px]);
      else if (gj == nyp - 1)
        v = a.bcfac * (s_pm[py - 1][px] - s_pm[py][px]);
      else if (!cyc && gi == 0)
        v = a.bcfac * (s_pm[py][px + 1] - s_pm[py][px]);
      else if (!cyc && gi == nxp - 1)
        v = a.bcfac * (s_pm[py][px - 1] - s_pm[py][px]);
      else
        v = (s_pm[py - 1][px] + s_pm[py][px - 1] + s_pm[py][px + 1] + s_pm[py + 1][px] - 4.0 * s_pm[py][px]) * g.dxm2;
    }
    s_d2[ly][lx] = v;
  }
  __syncthreads();
  // ---- del4 on the halo-1 region (qgosubs.F:310-341 / qgasubs.F:218-237)
  for (int e = tid; e < DH4 * DW4; e += 256) {
    const int ly = e / DW4, lx = e - ly * DW4;
    const int gj = j0 + ly - 1, gi = i0 + lx - 1;
    const int py = ly + 1, px = lx + 1;   // position in s_d2
    double v = 0.0;
    if (gj >= 0 && gj < nyp && (cyc || (gi >= 0 && gi < nxp))) {
      if (gj == 0)
        v = a.bcfac * (s_d2[py + 1][px] - s_d2[py][px]);
      else if (gj == nyp - 1)
        v = a.bcfac * (s_d2[py - 1][px] - s_d2[py][px]);
      else if (!cyc && gi == 0)
        v = a.bcfac * (s_d2[py][px + 1] - s_d2[py][px]);
      else if (!cyc && gi == nxp - 1)
        v = a.bcfac * (s_d2[py][px - 1] - s_d2[py][px]);
      else
        v = g.dxm2 * (s_d2[py - 1][px] + s_d2[py][px - 1] + s_d2[py][px + 1] + s_d2[py + 1][px] - 4.0 * s_d2[py][px]);
    }
    s_d4[ly][lx] = v;
  }
  __syncthreads();

  // ---- dq/dt, forcing, leapfrog (qgosubs.F:345-402, :184-219 / qgasubs.F:245-283, :115-146)
  const int nl = g.nl;
  for (int e = tid; e < TY * TX; e += 256) {
    const int ly = e / TX, lx = e - ly * TX;
    const int gj = j0 + ly, gi = i0 + lx;
    if (gj >= nyp || gi >= nxp) continue;
    const size_t idx = (size_t)gj * ld + gi;
    double *qm = a.qm + lo;
    if (gj == 0 || gj == nyp - 1) {
      // zonal boundary rows are not stepped: after the pointer rotation both time
      // levels hold the current boundary value (qgosubs.F:214-219)
      qm[idx] = q[idx];
      continue;
    }
    const int y = ly + 1, x = lx + 1;   // position in the halo-1 tiles
    double dqdt;
    if (!cyc && (gi == 0 || gi == nxp - 1)) {
      dqdt = 0.0;   // qgosubs.F:371, :397
    } else {
      const double d6p = g.dxm2 * (s_d4[y - 1][x] + s_d4[y][x - 1] + s_d4[y][x + 1] + s_d4[y + 1][x] - 4.0 * s_d4[y][x]);
#define Q(dx_, dy_) s_q[y + (dy_)][x + (dx_)]
#define P(dx_, dy_) s_p[y + (dy_)][x + (dx_)]
      const double jac = (Q(1, 0) - Q(-1, 0)) * (P(0, 1) - P(0, -1)) + (Q(0, -1) - Q(0, 1)) * (P(1, 0) - P(-1, 0)) +
                         Q(1, 0) * (P(1, 1) - P(1, -1)) - Q(-1, 0) * (P(-1, 1) - P(-1, -1)) -
                         Q(0, 1) * (P(1, 1) - P(-1, 1)) + Q(0, -1) * (P(1, -1) - P(-1, -1)) +
                         P(0, 1) * (Q(1, 1) - Q(-1, 1)) - P(0, -1) * (Q(1, -1) - Q(-1, -1)) -
                         P(1, 0) * (Q(1, 1) - Q(1, -1)) + P(-1, 0) * (Q(-1, 1) - Q(-1, -1));
#undef Q
#undef P
      if (a.atmos) {
        dqdt = a.adfac * jac - a.ah4fac[k] * d6p;
      } else {
        const double diffus = a.ah2fac[k] * s_d4[y][x] - a.ah4fac[k] * d6p;
        dqdt = a.adfac * jac + diffus;
      }
    }
    // layer-specific forcing; columns read through the canonical map so periodic
    // copies stay bit-identical
    const int ci = wrapx(gi, nxp, cyc);
    const size_t cidx = (size_t)gj * ld + ci;
    double qdot = dqdt;
    if (a.atmos) {
      if (k == 0) qdot = dqdt + a.fohfac[0] * (a.ent[cidx] - a.wek[cidx]);
      if (k == 1) qdot = dqdt - a.fohfac[1] * a.ent[cidx];
    } else {
      if (k == 0) qdot = dqdt + a.fohfac[0] * (a.wek[cidx] - a.ent[cidx]);
      if (k == 1) qdot = dqdt + a.fohfac[1] * a.ent[cidx];
      if (k == nl - 1) qdot = qdot - a.bdrfac * s_d2[y + 1][x + 1];
    }
    // qm is updated in place: read only this thread's own element (another block
    // owns column 1, so the periodic copy must not be read through the canonical map)
    qm[idx] = qm[idx] + g.tdt * qdot;
  }
}

// Boundary-strip sums feeding the momentum constraints of periodic channels:
// Jacobian strips (qgosubs.F:284-296, :409-423), third/fifth-derivative strips
// (:429-443; qgasubs.F:303-313) and the bottom-drag strip (qgosubs.F:155-162).
// One block per (layer, side); fixed-order block reduction.
struct StripArgs {
  Grid g;
  int atmos;
  double adfac, bcfac, f0, dxdy, delekfac;   // delekfac = 0.5 sign(f0) delek
  double ah2[NLMAX], ah4[NLMAX];
  const double *pm, *p, *q;
  qgcm_scalars *sc;
};

__device__ __forceinline__ double blk_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256) k_strips(StripArgs a) {
  __shared__ double red[32];
  const Grid &g = a.g;
  const int k = blockIdx.x, north = blockIdx.y;
  const size_t lo = (size_t)k * g.lsz;
  const double *pm = a.pm + lo, *p = a.p + lo, *q = a.q + lo;
  const int nxp = g.nxp, nyp = g.nyp, ld = g.ld, per = nxp - 1;
  // rows (0-based): south uses 0,1,2,3 ; north nyp-1, nyp-2, ...
  const int jb = north ? nyp - 1 : 0;       // boundary row
  const int dj = north ? -1 : 1;            // inward direction
  auto PM = [&](int i, int r) { return pm[(size_t)(jb + dj * r) * ld + ((i % per + per) % per)]; };
  auto d2 = [&](int i, int r) -> double {   // del2p at inward distance r (0 = boundary)
    if (r == 0) return a.bcfac * (PM(i, 1) - PM(i, 0));
    // interior formula order: (i,j-1)+(i-1,j)+(i+1,j)+(i,j+1)
    const double s = north ? PM(i, r + 1) : PM(i, r - 1);
    const double n = north ? PM(i, r - 1) : PM(i, r + 1);
    return (s + PM(i - 1, r) + PM(i + 1, r) + n - 4.0 * PM(i, r)) * g.dxm2;
  };
  auto d4 = [&](int i, int r) -> double {
    if (r == 0) return a.bcfac * (d2(i, 1) - d2(i, 0));
    const double s = north ? d2(i, r + 1) : d2(i, r - 1);
    const double n = north ? d2(i, r - 1) : d2(i, r + 1);
    return g.dxm2 * (s + d2(i - 1, r) + d2(i + 1, r) + n - 4.0 * d2(i, r));
  };
  double aj5 = 0.0, aj9 = 0.0, a3 = 0.0, a5 = 0.0, bd = 0.0;
  for (int i = threadIdx.x; i < nxp; i += blockDim.x) {
    // Jacobian strips: weights 0.5 at i=1 and i=nxp (qgosubs.F:284-296)
    const double wgt = (i == 0 || i == nxp - 1) ? 0.5 : 1.0;
    const int ic = i % per;
    const double dp = p[(size_t)(jb + dj) * ld + (ic + 1) % per] - p[(size_t)(jb + dj) * ld + (ic + per - 1) % per];
    const double q0 = q[(size_t)jb * ld + ic], q1 = q[(size_t)(jb + dj) * ld + ic];
    const double sgn = north ? -1.0 : 1.0;
    aj5 += sgn * wgt * q0 * dp;
    aj9 += sgn * wgt * q1 * dp;
    // derivative strips: south d(.)(2)-d(.)(1); north d(.)(nyp)-d(.)(nyp-1)
    const double s3 = north ? (d2(i, 0) - d2(i, 1)) : (d2(i, 1) - d2(i, 0));
    const double s5 = north ? (d4(i, 0) - d4(i, 1)) : (d4(i, 1) - d4(i, 0));
    const double sb = north ? (PM(i, 0) - PM(i, 1)) : (PM(i, 1) - PM(i, 0));
    if (a.atmos) {
      a5 += wgt * s5;                 // trapezoid weights (qgasubs.F:303-311)
    } else if (i < nxp - 1) {
      a3 += s3;                       // i = 1..nxpo-1 (qgosubs.F:433-438)
      a5 += s5;
      bd += sb;
    }
  }
  aj5 = blk_sum(aj5, red);
  aj9 = blk_sum(aj9, red);
  a3 = blk_sum(a3, red);
  a5 = blk_sum(a5, red);
  bd = blk_sum(bd, red);
  if (threadIdx.x == 0) {
    const double aj = a.dxdy * (a.f0 * a.adfac * (aj5 + 2.0 * aj9));
    qgcm_scalars *s = a.sc;
    if (a.atmos) {
      if (north) { s->ajinat[k] = aj; s->ap5nat[k] = a.ah4[k] * a5; }
      else       { s->ajisat[k] = aj; s->ap5sat[k] = a.ah4[k] * a5; }
    } else {
      if (north) { s->ajinoc[k] = aj; s->ap3noc[k] = a.ah2[k] * a3; s->ap5noc[k] = a.ah4[k] * a5; }
      else       { s->ajisoc[k] = aj; s->ap3soc[k] = a.ah2[k] * a3; s->ap5soc[k] = a.ah4[k] * a5; }
      if (k == g.nl - 1) {
        if (north) s->bdrinn = a.delekfac * bd; else s->bdrins = a.delekfac * bd;
      }
    }
  }
}

static void fill_common(qgcm_model *m, bool atmos, QgArgs &a, StripArgs &s) {
  const Grid &g = atmos ? m->ga : m->go;
  const LayerConsts &lc = atmos ? m->la : m->lo;
  const double bcco = atmos ? m->cfg.bccoat : m->cfg.bccooc;
  a.g = g;
  a.atmos = atmos;
  a.f0 = m->fnot;
  a.adfac = 1.0 / (12.0 * g.dx * g.dx * m->fnot);
  a.bcfac = bcco * g.dxm2 / (0.5 * bcco + 1.0);
  for (int k = 0; k < g.nl; ++k) {
    a.fohfac[k] = m->fnot / lc.h[k];
    a.ah2fac[k] = lc.ah2[k] / m->fnot;
    a.ah4fac[k] = lc.ah4[k] / m->fnot;
  }
  const double sgn = m->fnot >= 0.0 ? 1.0 : -1.0;
  a.bdrfac = atmos ? 0.0 : 0.5 * sgn * m->cfg.delek / lc.h[g.nl - 1];
  s.g = g;
  s.atmos = atmos;
  s.adfac = a.adfac;
  s.bcfac = a.bcfac;
  s.f0 = m->fnot;
  s.dxdy = g.dx * g.dx;
  s.delekfac = 0.5 * sgn * m->cfg.delek;
  for (int k = 0; k < g.nl; ++k) {
    s.ah2[k] = lc.ah2[k];
    s.ah4[k] = lc.ah4[k];
  }
  s.sc = m->d_scal;
}

static void launch(qgcm_model *m, bool atmos) {
  QgArgs a;
  StripArgs s;
  fill_common(m, atmos, a, s);
  const Grid &g = a.g;
  const char *np = atmos ? "pa" : "po", *npm = atmos ? "pam" : "pom";
  const char *nq = atmos ? "qa" : "qo", *nqm = atmos ? "qam" : "qom";
  a.pm = m->F(npm);
  a.p = m->F(np);
  a.q = m->F(nq);
  a.qm = m->F(nqm);
  a.wek = m->F(atmos ? "wekpa" : "wekpo");
  a.ent = m->F(atmos ? "entat" : "entoc");
  if (g.cyclic) {
    s.pm = a.pm; s.p = a.p; s.q = a.q;
    QG_LAUNCH(m, "k_strips", dim3(g.nl, 2), 256, 0, k_strips, s);
  }
  dim3 grid((g.nxp + TX - 1) / TX, (g.nyp + TY - 1) / TY, g.nl);
  QG_LAUNCH(m, "k_qgstep", grid, 256, 0, k_qgstep, a);
  QG_CUDA(cudaGetLastError());
  m->swapf(nq, nqm);   // new q lives in the old qom buffer; old q becomes qom
}

void launch_qgostep(qgcm_model *m) { launch(m, false); }
void launch_qgastep(qgcm_model *m) { launch(m, true); }

}  // namespace qg

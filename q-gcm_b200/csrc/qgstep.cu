// Leapfrog vorticity step, shared by ocean (qgostep + ocadif, src/qgosubs.F:45-446) and
// atmosphere (qgastep + atadif, src/qgasubs.F:45-317).
//
// One fused kernel: del2(pom) -> del4 -> del6, the Arakawa 9-point Jacobian J(q,p),
// forcing, bottom drag and the leapfrog update, with the three intermediate Laplacians
// kept in registers (halo 3 of pom, halo 1 of po/qo) instead of round-tripping through
// HBM as the reference's del2p/d4p/dqdt arrays do.
// q(new) is written over qom in place (qom is only read at the centre point), and the
// host rotates the qo/qom pointers afterwards.
#include <algorithm>
#include <cstdlib>

#include "qgcm_internal.h"

namespace qg {

constexpr int RCH = 128;    // most rows marched by one warp (6 pipeline fill rows per march)

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
  int mrows;                  // rows marched by one warp of k_qgstep2 (<= RCH)
  int erows;                  // > 0: the first and the last march cover only this many rows (the rows that need
                              // the boundary formulas), so that every other march runs the interior fast path
  // compact launches of the general loop (PART 0):
  //   pmode 1: one warp per block, only the warps that hold a wall or padding column (wx = 0 and
  //            wx >= w_hi), all rows in marches of mrows_edge rows;
  //   pmode 2: the two short boundary marches for the warps in between
  int pmode, w_hi, mrows_edge;
};

__device__ __forceinline__ double shl(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }    // value of lane-1 (west)
__device__ __forceinline__ double shr(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }  // value of lane+1 (east)

constexpr int QG_NF = 6;    // fields per stage: pom, p, q, qm, wek, ent

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------
// Warp-marching stencil pipeline.  A warp covers 64 consecutive columns -- each lane owns the
// even/odd pair (g0, g0+1), 56 outputs + a halo of 4 (the del-6th needs 3; 4 keeps the pairs
// 16-byte aligned) -- and marches north.  Every lane keeps the last three rows of pom, del2,
// del4, p and q of its two columns in registers; east/west neighbours come from one warp
// shuffle pair per field row (no block barriers), and the rows ahead are prefetched Q2_D deep
// with 16-byte cp.async into a per-warp shared-memory ring in which each lane only ever
// touches its own slots.  (The first version kept one column per lane: 60 % of the HBM peak,
// issue bound; this one 69 %.)
// ------------------------------------------------------------------------------------------
constexpr int W2OUT = 56;
constexpr int Q2_D = 4;     // row stages in flight per warp
__device__ __forceinline__ void cp_async16(double2 *smem_dst, const double *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
// 5-point operator with the mixed condition on solid walls (qgosubs.F:86-130, :310-341)
__device__ __forceinline__ double qg_lap(double c, double w, double e, double s, double n, int j, int nyp, bool wallW, bool wallE,
                                         double bcf, double dxm2) {
  if (j == 0) return bcf * (n - c);
  if (j == nyp - 1) return bcf * (s - c);
  if (wallW) return bcf * (e - c);
  if (wallE) return bcf * (w - c);
  return (s + w + e + n - 4.0 * c) * dxm2;
}
// Arakawa 9-point Jacobian J(q,p) at one point; rows A,B,C = j-1, j, j+1 (qgosubs.F:376-388)
__device__ __forceinline__ double qg_jac(double pA, double pAw, double pAe, double pBw, double pBe, double pC, double pCw,
                                         double pCe, double qA, double qAw, double qAe, double qBw, double qBe, double qC,
                                         double qCw, double qCe) {
  return (qBe - qBw) * (pC - pA) + (qA - qC) * (pBe - pBw) + qBe * (pCe - pAe) - qBw * (pCw - pAw) - qC * (pCe - pCw) +
         qA * (pAe - pAw) + pC * (qCe - qCw) - pA * (qAe - qAw) - pBe * (qCe - qAe) + pBw * (qCw - qAw);
}

// ---- interior fast path -------------------------------------------------------------------
// A warp whose 64 columns hold no wall or padding column and whose march needs no boundary row
// (the vast majority at benchmark sizes) runs the same arithmetic without the wall / boundary
// predicates, with its row windows rotated at compile time (the march is unrolled by three, so
// "the last three rows" are three fixed register sets instead of two register moves per value
// and row) and with the pipeline fill split from the steady state (no output predicate).
struct QgWin {
  double pm[3][2], d2[3][2], d4[3][2], p[3][2], q[3][2];
  double pw[3], pe[3], qw[3], qe[3];      // west neighbour of column 0, east neighbour of column 1
};
// the constants of a row step are read from the kernel's parameter block where they are used
// (constant-bank operands), not copied into registers
struct QgRowK {
  const QgArgs &a;
  int k, kind;       // kind 0: top layer (wek, ent), 1: second layer (ent), 2: unforced; 3: unforced + ocean bottom drag
};
template <int PH, bool OUT>
__device__ __forceinline__ void qg_row_interior(QgWin &S, const double2 *slot, const QgRowK &cq, double *qm_row, bool outl) {
  struct { double dxm2, adf, ah2f, ah4f, tdt, foh, bdr; int atmos, kind; } c = {
      cq.a.g.dxm2, cq.a.adfac, cq.a.ah2fac[cq.k], cq.a.ah4fac[cq.k], cq.a.g.tdt, cq.a.fohfac[cq.k < 2 ? cq.k : 0], cq.a.bdrfac, cq.a.atmos, cq.kind};
  constexpr int N = PH, M = (PH + 2) % 3, O = (PH + 1) % 3;      // newest, middle, oldest row of every window
  const double2 vpm = slot[0], vp = slot[32], vq = slot[64];
  S.pm[N][0] = vpm.x; S.pm[N][1] = vpm.y;
  S.p[N][0] = vp.x; S.p[N][1] = vp.y;
  S.q[N][0] = vq.x; S.q[N][1] = vq.y;
  S.pw[N] = shl(vp.y); S.pe[N] = shr(vp.x);
  S.qw[N] = shl(vq.y); S.qe[N] = shr(vq.x);
  {   // del2 at row r-1 (same association as qg_lap)
    const double w0 = shl(S.pm[M][1]), e1 = shr(S.pm[M][0]);
    S.d2[N][0] = (S.pm[O][0] + w0 + S.pm[M][1] + S.pm[N][0] - 4.0 * S.pm[M][0]) * c.dxm2;
    S.d2[N][1] = (S.pm[O][1] + S.pm[M][0] + e1 + S.pm[N][1] - 4.0 * S.pm[M][1]) * c.dxm2;
  }
  {   // del4 at row r-2
    const double w0 = shl(S.d2[M][1]), e1 = shr(S.d2[M][0]);
    S.d4[N][0] = (S.d2[O][0] + w0 + S.d2[M][1] + S.d2[N][0] - 4.0 * S.d2[M][0]) * c.dxm2;
    S.d4[N][1] = (S.d2[O][1] + S.d2[M][0] + e1 + S.d2[N][1] - 4.0 * S.d2[M][1]) * c.dxm2;
  }
  const double d4w0 = shl(S.d4[M][1]), d4e1 = shr(S.d4[M][0]);
  if (!OUT) return;
  // row r-3: del6, Jacobian, forcing, leapfrog
  const double2 qold = slot[96];
  double qn[2];
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const double d4w = cc == 0 ? d4w0 : S.d4[M][0], d4e = cc == 0 ? S.d4[M][1] : d4e1;
    const double d6p = c.dxm2 * (S.d4[O][cc] + d4w + d4e + S.d4[N][cc] - 4.0 * S.d4[M][cc]);
    const double jac = cc == 0 ? qg_jac(S.p[O][0], S.pw[O], S.p[O][1], S.pw[M], S.p[M][1], S.p[N][0], S.pw[N], S.p[N][1],
                                        S.q[O][0], S.qw[O], S.q[O][1], S.qw[M], S.q[M][1], S.q[N][0], S.qw[N], S.q[N][1])
                               : qg_jac(S.p[O][1], S.p[O][0], S.pe[O], S.p[M][0], S.pe[M], S.p[N][1], S.p[N][0], S.pe[N],
                                        S.q[O][1], S.q[O][0], S.qe[O], S.q[M][0], S.qe[M], S.q[N][1], S.q[N][0], S.qe[N]);
    double dqdt;
    if (c.atmos) {
      dqdt = c.adf * jac - c.ah4f * d6p;
    } else {
      const double diffus = c.ah2f * S.d4[M][cc] - c.ah4f * d6p;
      dqdt = c.adf * jac + diffus;
    }
    double qdot = dqdt;
    if (c.kind == 0) {
      const double2 wk = slot[128], en = slot[160];
      const double wkc = cc == 0 ? wk.x : wk.y, enc = cc == 0 ? en.x : en.y;
      qdot = c.atmos ? dqdt + c.foh * (enc - wkc) : dqdt + c.foh * (wkc - enc);
    } else if (c.kind == 1) {
      const double2 en = slot[160];
      const double enc = cc == 0 ? en.x : en.y;
      qdot = c.atmos ? dqdt - c.foh * enc : dqdt + c.foh * enc;
    } else if (c.kind == 3) {
      qdot = qdot - c.bdr * S.d2[O][cc];      // del2p(i, jo)
    }
    qn[cc] = (cc == 0 ? qold.x : qold.y) + c.tdt * qdot;
  }
  if (outl) *reinterpret_cast<double2 *>(qm_row) = make_double2(qn[0], qn[1]);
}

// rows [ja, jb) of march `by`
__device__ __forceinline__ void qg_march_rows(const QgArgs &a, int by, int nby, int nyp, int &ja, int &jb) {
  if (a.erows > 0) {
    if (by == 0) { ja = 0; jb = a.erows; }
    else if (by == nby - 1) { ja = nyp - a.erows; jb = nyp; }
    else { ja = a.erows + (by - 1) * a.mrows; jb = min(nyp - a.erows, ja + a.mrows); }
  } else {
    ja = by * a.mrows; jb = min(nyp, ja + a.mrows);
  }
}

// PART 0: the warps that touch a wall column or a boundary row (general loop); PART 1: the interior
// warps (fast path).  Two kernels so that each gets its own register allocation; every warp belongs to
// exactly one of them and each element of q is read and written by exactly one lane.
template <int PART, int BLK>
__global__ void __launch_bounds__(128, BLK) k_qgstep2(QgArgs a) {
  extern __shared__ double2 ring2_all[];
  const Grid &g = a.g;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int k = blockIdx.z;
  const int nxp = g.nxp, nyp = g.nyp, ld = g.ld, cyc = g.cyclic, per = nxp - 1;
  int wx = blockIdx.x * 4 + wib, ja, jb;
  qg_march_rows(a, blockIdx.y, gridDim.y, nyp, ja, jb);
  if (PART == 0 && a.pmode == 1) {        // one warp per block: the warps next to the walls, every row
    wx = blockIdx.x == 0 ? 0 : a.w_hi + (int)blockIdx.x - 1;
    ja = blockIdx.y * a.mrows_edge; jb = min(nyp, ja + a.mrows_edge);
  } else if (PART == 0 && a.pmode == 2) { // the two boundary marches
    ja = blockIdx.y ? nyp - a.erows : 0; jb = ja + a.erows;
  }
  if (wx * W2OUT >= nxp) return;   // whole warp exits together
  bool interior = false;
  {
    // interior warp: columns g0 .. g0+63 all strictly inside the walls (box) or inside one period
    // without wrap (channel), and the rows the outputs depend on (ja-3 .. jb+2) inside the domain;
    // a two-layer model's second layer is also the bottom layer: forcing and drag together are
    // left to the general loop
    const int gw = wx * W2OUT - 4;
    const bool cols_in = cyc ? (gw >= 0 && gw + 64 <= per) : (gw >= 1 && gw + 64 <= nxp - 1);
    const bool rows_in = ja >= 3 && jb <= nyp - 3;
    const bool simple_kind = !(k < 2 && !a.atmos && k == g.nl - 1);
    interior = cols_in && rows_in && simple_kind;
    if (PART == 0 && a.pmode == 0 && interior) return;      // PART 2: every warp takes the general loop; PART 3: each warp its own
    if (PART == 0 && a.pmode == 2 && !cols_in) return;      // the wall warps did these rows themselves (pmode 1)
    if (PART == 1 && !interior) return;
  }
  double2 *ring = ring2_all + (size_t)wib * (Q2_D * QG_NF * 32) + lane;
  const int g0 = wx * W2OUT - 4 + 2 * lane;      // first column of the pair (even)
  int c0 = g0;                                    // canonical column for loads
  if (cyc) { if (c0 < 0) c0 += per; if (c0 >= per) c0 -= per; }
  const bool ld0 = c0 >= 0 && c0 < nxp;           // the pair is loaded when its first column exists
  const bool v0 = ld0, v1 = ld0 && (cyc || c0 + 1 < nxp);   // column validity (second one may be row padding)
  const bool wallW0 = !cyc && g0 == 0, wallE0 = !cyc && g0 == nxp - 1;
  const bool wallE1 = !cyc && g0 + 1 == nxp - 1;  // the odd column is never the western wall
  const bool outl = lane >= 2 && lane < 30;
  const bool out0 = outl && g0 < nxp, out1 = outl && g0 + 1 < nxp;
  const size_t lo = (size_t)k * g.lsz;
  const int cc = ld0 ? c0 : 0;
  const double *__restrict__ pm = a.pm + lo + cc;
  const double *__restrict__ p = a.p + lo + cc;
  const double *__restrict__ q = a.q + lo + cc;
  double *__restrict__ qm = a.qm + lo + (g0 >= 0 && g0 < nxp ? g0 : 0);
  const double *__restrict__ wek = a.wek + cc;
  const double *__restrict__ ent = a.ent + cc;
  const double dxm2 = g.dxm2, bcf = a.bcfac;
  const double ah2f = a.ah2fac[k], ah4f = a.ah4fac[k], adf = a.adfac, tdt = g.tdt;
  const int nl = g.nl;
  const bool forced = k < 2;
  const bool ldq = g0 >= 0 && g0 < nxp;

#pragma unroll
  for (int s = 0; s < Q2_D * QG_NF; ++s) ring[s * 32] = make_double2(0.0, 0.0);
  __syncwarp();
  // stage r carries pom(r), p(r-2), q(r-2), qm(r-3), wek(r-3), ent(r-3); rows outside the
  // domain are clamped (their values only reach results that are never used)
  auto issue = [&](int r) {
    double2 *slot = ring + (size_t)((r + 8 * Q2_D) % Q2_D) * (QG_NF * 32);
    const int r0c = min(max(r, 0), nyp - 1), r2c = min(max(r - 2, 0), nyp - 1), r3c = min(max(r - 3, 0), nyp - 1);
    if (ld0) {
      cp_async16(slot, pm + (size_t)r0c * ld);
      cp_async16(slot + 32, p + (size_t)r2c * ld);
      cp_async16(slot + 64, q + (size_t)r2c * ld);
      if (forced) {
        if (k == 0) cp_async16(slot + 128, wek + (size_t)r3c * ld);      // only the top layer feels the Ekman pumping
        cp_async16(slot + 160, ent + (size_t)r3c * ld);
      }
    }
    if (ldq) cp_async16(slot + 96, qm + (size_t)r3c * ld);
    cp_async_commit();
  };

  // [0] / [1]: the lane's two columns; w0 = west neighbour of column 0 (lane-1's second column),
  // e1 = east neighbour of column 1 (lane+1's first column)
  double pm0[2] = {0, 0}, pm1[2] = {0, 0}, pm2[2] = {0, 0}, d2a[2] = {0, 0}, d2b[2] = {0, 0}, d2c[2] = {0, 0};
  double d4a[2] = {0, 0}, d4b[2] = {0, 0}, d4c[2] = {0, 0};
  double pA[2] = {0, 0}, pB[2] = {0, 0}, pC[2] = {0, 0}, qA[2] = {0, 0}, qB[2] = {0, 0}, qC[2] = {0, 0};
  double pAw0 = 0, pAe1 = 0, pBw0 = 0, pBe1 = 0, pCw0 = 0, pCe1 = 0, qAw0 = 0, qAe1 = 0, qBw0 = 0, qBe1 = 0, qCw0 = 0, qCe1 = 0;
  const int r0 = ja - 3;
#pragma unroll
  for (int s = 0; s < Q2_D - 1; ++s) issue(r0 + s);
  {
    if (PART == 1 || (PART == 3 && interior)) {
      const QgRowK ck = {a, k, k == 0 ? 0 : k == 1 ? 1 : (!a.atmos && k == nl - 1) ? 3 : 2};
      QgWin S;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) S.pm[i][cc] = S.d2[i][cc] = S.d4[i][cc] = S.p[i][cc] = S.q[i][cc] = 0.0;
        S.pw[i] = S.pe[i] = S.qw[i] = S.qe[i] = 0.0;
      }
      int r = r0;
      auto slot_of = [&](int rr) { return ring + (size_t)((rr + 8 * Q2_D) % Q2_D) * (QG_NF * 32); };
      // pipeline fill: six rows without output (phases 0,1,2,0,1,2)
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        issue(r + Q2_D - 1); cp_async_wait<Q2_D - 1>(); qg_row_interior<0, false>(S, slot_of(r), ck, nullptr, false); ++r;
        issue(r + Q2_D - 1); cp_async_wait<Q2_D - 1>(); qg_row_interior<1, false>(S, slot_of(r), ck, nullptr, false); ++r;
        issue(r + Q2_D - 1); cp_async_wait<Q2_D - 1>(); qg_row_interior<2, false>(S, slot_of(r), ck, nullptr, false); ++r;
      }
      // steady state: row r produces output row r-3
      double *qrow = qm + (size_t)(r - 3) * ld;
      const int rend = jb + 3;
      for (; r + 2 < rend; r += 3, qrow += 3 * (size_t)ld) {
        issue(r + Q2_D - 1); cp_async_wait<Q2_D - 1>(); qg_row_interior<0, true>(S, slot_of(r), ck, qrow, outl);
        issue(r + Q2_D); cp_async_wait<Q2_D - 1>(); qg_row_interior<1, true>(S, slot_of(r + 1), ck, qrow + ld, outl);
        issue(r + Q2_D + 1); cp_async_wait<Q2_D - 1>(); qg_row_interior<2, true>(S, slot_of(r + 2), ck, qrow + 2 * (size_t)ld, outl);
      }
      if (r < rend) {
        issue(r + Q2_D - 1); cp_async_wait<Q2_D - 1>(); qg_row_interior<0, true>(S, slot_of(r), ck, qrow, outl); ++r;
        if (r < rend) {
          issue(r + Q2_D - 1); cp_async_wait<Q2_D - 1>(); qg_row_interior<1, true>(S, slot_of(r), ck, qrow + ld, outl);
        }
      }
      cp_async_wait<0>();
      return;
    }
  }
  if (PART == 1) return;
  for (int r = r0; r < jb + 3; ++r) {
    issue(r + Q2_D - 1);
    cp_async_wait<Q2_D - 1>();
    const double2 *slot = ring + (size_t)((r + 8 * Q2_D) % Q2_D) * (QG_NF * 32);
    {
      const double2 vpm = slot[0], vp = slot[32], vq = slot[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        pm0[c] = pm1[c]; pm1[c] = pm2[c];
        pA[c] = pB[c]; pB[c] = pC[c];
        qA[c] = qB[c]; qB[c] = qC[c];
      }
      pm2[0] = vpm.x; pm2[1] = v1 ? vpm.y : 0.0;
      pC[0] = vp.x; pC[1] = v1 ? vp.y : 0.0;
      qC[0] = vq.x; qC[1] = v1 ? vq.y : 0.0;
      pAw0 = pBw0; pAe1 = pBe1; pBw0 = pCw0; pBe1 = pCe1; pCw0 = shl(pC[1]); pCe1 = shr(pC[0]);
      qAw0 = qBw0; qAe1 = qBe1; qBw0 = qCw0; qBe1 = qCe1; qCw0 = shl(qC[1]); qCe1 = shr(qC[0]);
    }
    // ---- del2 at row r-1
    {
      const int j2 = r - 1;
      const double w0 = shl(pm1[1]), e1 = shr(pm1[0]);
      const double a0 = qg_lap(pm1[0], w0, pm1[1], pm0[0], pm2[0], j2, nyp, wallW0, wallE0, bcf, dxm2);
      const double a1 = qg_lap(pm1[1], pm1[0], e1, pm0[1], pm2[1], j2, nyp, false, wallE1, bcf, dxm2);
      d2a[0] = d2b[0]; d2b[0] = d2c[0]; d2c[0] = a0;
      d2a[1] = d2b[1]; d2b[1] = d2c[1]; d2c[1] = a1;
    }
    // ---- del4 at row r-2
    {
      const int j4 = r - 2;
      const double w0 = shl(d2b[1]), e1 = shr(d2b[0]);
      const double a0 = qg_lap(d2b[0], w0, d2b[1], d2a[0], d2c[0], j4, nyp, wallW0, wallE0, bcf, dxm2);
      const double a1 = qg_lap(d2b[1], d2b[0], e1, d2a[1], d2c[1], j4, nyp, false, wallE1, bcf, dxm2);
      d4a[0] = d4b[0]; d4b[0] = d4c[0]; d4c[0] = a0;
      d4a[1] = d4b[1]; d4b[1] = d4c[1]; d4c[1] = a1;
    }
    // ---- row r-3: del6, Jacobian, forcing, leapfrog
    const int jo = r - 3;
    const double d4w0 = shl(d4b[1]), d4e1 = shr(d4b[0]);
    if (jo < ja || jo >= jb || !out0) continue;
    const size_t ro = (size_t)jo * ld;
    const double2 qold = slot[96];
    double qn[2];
    if (jo == 0 || jo == nyp - 1) {
      // zonal boundary rows are not stepped: after the pointer rotation both time
      // levels hold the current boundary value (qgosubs.F:214-219)
      qn[0] = qB[0]; qn[1] = qB[1];
    } else {
      double2 wk = make_double2(0.0, 0.0), en = make_double2(0.0, 0.0);
      if (forced) { if (k == 0) wk = slot[128]; en = slot[160]; }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const bool wall = c == 0 ? (wallW0 || wallE0) : wallE1;
        double dqdt;
        if (wall) {
          dqdt = 0.0;   // qgosubs.F:371, :397
        } else {
          const double d4w = c == 0 ? d4w0 : d4b[0], d4e = c == 0 ? d4b[1] : d4e1;
          const double d6p = dxm2 * (d4a[c] + d4w + d4e + d4c[c] - 4.0 * d4b[c]);
          const double jac = c == 0 ? qg_jac(pA[0], pAw0, pA[1], pBw0, pB[1], pC[0], pCw0, pC[1], qA[0], qAw0, qA[1], qBw0, qB[1],
                                             qC[0], qCw0, qC[1])
                                    : qg_jac(pA[1], pA[0], pAe1, pB[0], pBe1, pC[1], pC[0], pCe1, qA[1], qA[0], qAe1, qB[0], qBe1,
                                             qC[1], qC[0], qCe1);
          if (a.atmos) {
            dqdt = adf * jac - ah4f * d6p;
          } else {
            const double diffus = ah2f * d4b[c] - ah4f * d6p;
            dqdt = adf * jac + diffus;
          }
        }
        double qdot = dqdt;
        if (forced) {
          const double wkc = c == 0 ? wk.x : wk.y, enc = c == 0 ? en.x : en.y;
          if (a.atmos) {
            if (k == 0) qdot = dqdt + a.fohfac[0] * (enc - wkc);
            if (k == 1) qdot = dqdt - a.fohfac[1] * enc;
          } else {
            if (k == 0) qdot = dqdt + a.fohfac[0] * (wkc - enc);
            if (k == 1) qdot = dqdt + a.fohfac[1] * enc;
          }
        }
        if (!a.atmos && k == nl - 1) qdot = qdot - a.bdrfac * d2a[c];   // d2a = del2p(i, jo) after the shifts above
        qn[c] = (c == 0 ? qold.x : qold.y) + tdt * qdot;
      }
    }
    // qm is updated in place; each element is read (prefetched) and written by exactly one lane
    if (out1) *reinterpret_cast<double2 *>(qm + ro) = make_double2(qn[0], qn[1]);
    else qm[ro] = qn[0];
  }
  cp_async_wait<0>();
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
  a.mrows = RCH;
  a.erows = 0;
  a.pmode = 0; a.w_hi = 0; a.mrows_edge = RCH;
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
  {
    // marches sized to whole waves of resident blocks, at least 24 rows each (6 fill rows per march)
    const int nwx = (g.nxp + W2OUT - 1) / W2OUT;
    const size_t smem = 4 * Q2_D * QG_NF * 32 * sizeof(double2);
    int &resident = m->qg_resident;
    if (!resident) {
      QG_CUDA(cudaFuncSetAttribute((k_qgstep2<0, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QG_CUDA(cudaFuncSetAttribute((k_qgstep2<2, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QG_CUDA(cudaFuncSetAttribute((k_qgstep2<3, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QG_CUDA(cudaFuncSetAttribute((k_qgstep2<1, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      QG_CUDA(cudaFuncSetAttribute((k_qgstep2<1, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0, sms = 0;
      if (env_int("QGCM_QG_BLK", 2) == 1)
        QG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (k_qgstep2<3, 3>), 128, smem));
      else if (env_int("QGCM_QG_BLK", 2) == 3)
        QG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (k_qgstep2<1, 3>), 128, smem));
      else
        QG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (k_qgstep2<1, 4>), 128, smem));
      QG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->cfg.device));
      resident = std::max(1, per_sm * sms);
    }
    const int bx = ((nwx + 3) / 4) * g.nl;
    // the three rows next to a zonal boundary need the boundary formulas: they get two short marches
    // of their own, every march between them is interior in y
    a.erows = (g.nyp >= 64 && env_int("QGCM_QG_EROWS", 1)) ? 4 : 0;
    const int inner = g.nyp - 2 * a.erows;
    a.mrows = pick_march_rows(inner, bx, resident, 24, RCH, a.erows ? 2 : 0);
    a.mrows = std::max(8, std::min(1024, env_int("QGCM_QG_MROWS", a.mrows)));
    dim3 grid((nwx + 3) / 4, (inner + a.mrows - 1) / a.mrows + (a.erows ? 2 : 0), g.nl);
    // 2 (default): interior warps in one launch at four blocks per SM; the general loop in two compact
    //   launches -- the warps next to the walls, one warp per block in short marches (their rows are a
    //   serial chain: 128-row marches would make this launch as long as a whole wave of the main one),
    //   and the two boundary marches of everyone else;
    // 1: one launch, every warp takes the path that fits it (three blocks per SM);
    // 0: the general loop for every warp (round-1 kernel); 3 / 4: A/B experiments
    // (a two-layer ocean, whose second layer is forced and dragged at once, stays on the general loop)
    const int mode = (a.erows == 0 || (!atmos && g.nl == 2)) ? 0 : env_int("QGCM_QG_BLK", 2);
    if (mode == 2) {
      // the wall-warp launches run beside the interior launch on a second stream (they write
      // disjoint elements); while profiling they stay on the main stream so that they are timed
      // (ocean only: the atmosphere's launches are a few blocks each, and inside a coupled cycle its steps
      // run beside the ocean step, which owns the side stream)
      const bool side = !m->prof && !atmos && env_int("QGCM_QG_SIDE", 1);
      if (side) {
        if (!m->side_stream) {
          QG_CUDA(cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking));
          QG_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
          QG_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
        }
        QG_CUDA(cudaEventRecord(m->ev_fork, m->stream));
        QG_CUDA(cudaStreamWaitEvent(m->side_stream, m->ev_fork, 0));
      }
      QG_LAUNCH(m, "k_qgstep", grid, 128, smem, (k_qgstep2<1, 4>), a);
      QgArgs e = a;
      // warps that are not interior in x: wx = 0 and wx >= w_hi
      int w_hi = nwx;
      while (w_hi > 1) {
        const int gw = (w_hi - 1) * W2OUT - 4;
        const bool in = g.cyclic ? (gw >= 0 && gw + 64 <= g.nxp - 1) : (gw >= 1 && gw + 64 <= g.nxp - 1);
        if (in) break;
        --w_hi;
      }
      e.pmode = 1; e.w_hi = w_hi; e.mrows_edge = std::max(8, env_int("QGCM_QG_EDGE_ROWS", 16));
      const dim3 g1(1 + (nwx - w_hi), (g.nyp + e.mrows_edge - 1) / e.mrows_edge, g.nl), g2((nwx + 3) / 4, 2, g.nl);
      if (side) {
        k_qgstep2<0, 4><<<g1, 32, smem / 4, m->side_stream>>>(e);
        launch_check(m, "k_qgstep_edge");
        e.pmode = 2;
        k_qgstep2<0, 4><<<g2, 128, smem, m->side_stream>>>(e);
        launch_check(m, "k_qgstep_edge");
        m->launches += 2;
        QG_CUDA(cudaEventRecord(m->ev_join, m->side_stream));
        QG_CUDA(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
      } else {
        QG_LAUNCH(m, "k_qgstep_edge", g1, 32, smem / 4, (k_qgstep2<0, 4>), e);
        e.pmode = 2;
        QG_LAUNCH(m, "k_qgstep_edge", g2, 128, smem, (k_qgstep2<0, 4>), e);
      }
    } else if (mode == 1) {
      QG_LAUNCH(m, "k_qgstep", grid, 128, smem, (k_qgstep2<3, 3>), a);
    } else if (mode == 0) {
      QG_LAUNCH(m, "k_qgstep", grid, 128, smem, (k_qgstep2<2, 4>), a);
    } else {
      if (mode == 3)
        QG_LAUNCH(m, "k_qgstep", grid, 128, smem, (k_qgstep2<1, 3>), a);        // interior warps
      else
        QG_LAUNCH(m, "k_qgstep", grid, 128, smem, (k_qgstep2<1, 4>), a);
      QG_LAUNCH(m, "k_qgstep_edge", grid, 128, smem, (k_qgstep2<0, 4>), a);     // warps on walls and boundary rows
    }
  }
  QG_CUDA(cudaGetLastError());
  m->swapf(nq, nqm);   // new q lives in the old qom buffer; old q becomes qom
}

void launch_qgostep(qgcm_model *m) { launch(m, false); }
void launch_qgastep(qgcm_model *m) { launch(m, true); }

}  // namespace qg

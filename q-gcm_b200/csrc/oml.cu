// Ocean mixed layer: oml + omladf (src/omlsubs.F:47-763).
//
// Pass 1 (k_oml_step): per T-cell tile, stage sstm (halo 2), sst (halo 1) and the p-grid
//   corner values of po(:,:,1), tauxo, tauyo in shared memory; compute del2(sstm) with
//   every boundary-condition variant, the C-grid flux-form advection, del2+del4
//   diffusion, the sst prediction, entrainment and convective adjustment; write the new
//   sst to a third buffer (sstm is read with a stencil, so it cannot be updated in
//   place) and xfo, and emit per-block partial sums for xfosum / cfrasm / centsm.
// Pass 2 (k_oml_entoc): subtract the global mean from xfo and average the four
//   surrounding T cells onto p points (entoc), with xintp row sums for xon(1).
// The reference needs ~24 field passes for this; here it is 11.
#include <algorithm>
#include <cstdlib>

#include "qgcm_internal.h"

namespace qg {


struct OmlArgs {
  Grid g;
  int sb, nb;
  double tsbdy, tnbdy;
  double uvgfac, rhf0hm, d2tfac, d4tfac, hdxm1;
  double hmoinv, dtoinv, entfac, rrcpoc, toc1, tdt;
  const double *po1, *taux, *tauy, *sst, *sstm, *wekt, *fnet;
  double *sstnew, *xfo;
  double *part;      // [3][nblocks] partial sums
  int nblocks;
  double *entoc;
  double *rowsum;    // [nyp] xintp row sums of entoc
  qgcm_scalars *sc;
  // y-slabs: T rows [t0, t1) and p rows [p0, p1) are owned by this rank; the three sums and
  // the entoc integral travel through the reduction vector cv
  int t0, t1, p0, p1, multi;
  double *cv;
  double *part2;          // [3][ORB] slice sums of the block partials
  unsigned int *ticket;   // last-block-done counter of k_oml_reduce
  int mrows;              // rows marched by one warp of k_oml_march
  PeerCtx peer;           // y-slabs over peer memory: k_oml_reduce all-reduces its three sums itself
  int *peer_err;
};

__device__ __forceinline__ int wrapt(int i, int nxt, int cyc) {
  if (!cyc) return i;
  if (i < 0) i += nxt;
  if (i >= nxt) i -= nxt;
  return i;
}

// ------------------------------------------------------------------------------------------
// The step as a marching pipeline: a warp owns 64 T columns (each lane the even/odd pair, 60
// outputs + halo 2 for del4 of sstm) and marches north; every lane keeps three rows of sstm,
// del2t and sst and two rows of the p-grid fields of its columns in registers, E/W neighbours
// come from warp shuffles and the rows ahead are prefetched with 16-byte cp.async into a
// per-warp ring.  (A 64x8 shared-memory tile version ran at 51 % of the HBM peak, this one at
// 72 %.)
// ------------------------------------------------------------------------------------------
constexpr int MW = 60;      // output columns per warp: 32 lanes x 2 columns minus a halo of 2 columns on either side
constexpr int MR = 256;     // most rows marched by one warp (fewer on small grids / slabs, to fill the GPU)
#ifndef QGCM_OML_MD
#define QGCM_OML_MD 4
#endif
constexpr int MD = QGCM_OML_MD;       // prefetch depth (row stages in flight; 14 KB of shared memory per warp)
constexpr int MNF = 7;      // fields per stage: sstm, sst, p, taux, tauy, wekt, fnet

__device__ __forceinline__ void oml_cp16(double2 *smem_dst, const double *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ double shw(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }     // value of lane-1 (west)
__device__ __forceinline__ double she(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }   // value of lane+1 (east)

// del2t of the lagged temperature at one T cell with every boundary-condition variant
// (omlsubs.F:291-682); neighbour order as in the reference so the sums round identically
__device__ __forceinline__ double oml_del2(const OmlArgs &a, int gi, int gj, int nxt, int nyt, int cyc, double c, double w,
                                           double e, double s, double n) {
  const bool hasW = cyc || gi > 0, hasE = cyc || gi < nxt - 1;
  double sum, cnt;
  if (gj == 0) {
    sum = 0.0; cnt = 0.0;
    if (hasW) { sum = w; cnt += 1.0; }
    if (hasE) { sum = (cnt > 0.0) ? sum + e : e; cnt += 1.0; }
    sum = sum + n; cnt += 1.0;
    if (a.sb) { sum = sum + a.tsbdy; cnt += 1.0; }
    return sum - cnt * c;
  }
  if (gj == nyt - 1) {
    const bool ne_cyc = cyc && gi == nxt - 1;
    sum = s; cnt = 1.0;
    if (hasW) { sum = sum + w; cnt += 1.0; }
    if (a.nb && !ne_cyc) { sum = sum + a.tnbdy; cnt += 1.0; }
    if (hasE) { sum = sum + e; cnt += 1.0; }
    if (a.nb && ne_cyc) return sum - 4.0 * c + a.tnbdy;
    return sum - cnt * c;
  }
  sum = s; cnt = 1.0;
  if (hasW) { sum = sum + w; cnt += 1.0; }
  if (hasE) { sum = sum + e; cnt += 1.0; }
  sum = sum + n; cnt += 1.0;
  return sum - cnt * c;
}

// advection + diffusion + prediction + entrainment + convection at one T cell
// (omlsubs.F:297-384, :728-758, :94-127); p/x/y are po(1), tauxo, tauyo at the cell's corners
// (S/N row, plain = west corner, e = east corner)
struct OmlOut { double sst, xf, cfr, cen; };
__device__ __forceinline__ OmlOut oml_cell(const OmlArgs &a, int gi, int gj, int nxt, int nyt, int cyc, double tc, double tW,
                                           double tE, double tS, double tN, double dcen, double dW, double dE, double dS, double dN,
                                           double tmc, double pS, double pN, double pSe, double pNe, double xS, double xN, double xSe,
                                           double xNe, double yS, double yN, double ySe, double yNe, double wek, double fnet) {
  const bool wallW = !cyc && gi == 0, wallE = !cyc && gi == nxt - 1;
  double um = -a.uvgfac * (pN - pS) + a.rhf0hm * (yN + yS);
  double up = -a.uvgfac * (pNe - pSe) + a.rhf0hm * (yNe + ySe);
  double tm = tc + tW, tp = tc + tE;
  if (wallW) { um = 0.0; tm = 0.0; }
  if (wallE) { up = 0.0; tp = 0.0; }
  const double hxadv = a.hdxm1 * (up * tp - um * tm);
  const double vs = a.uvgfac * (pSe - pS) - a.rhf0hm * (xSe + xS);
  const double vn = a.uvgfac * (pNe - pN) - a.rhf0hm * (xNe + xN);
  double hyadv;
  if (gj == 0) {
    const double tpn = tc + tN;
    if (a.sb) {
      const double vm = -a.rhf0hm * (xSe + xS);
      const double tms = tc + a.tsbdy;
      hyadv = a.hdxm1 * (vn * tpn - vm * tms);
    } else {
      hyadv = a.hdxm1 * (vn * tpn);
    }
  } else if (gj == nyt - 1) {
    const double tms = tS + tc;
    if (a.nb) {
      const double vp = -a.rhf0hm * (xNe + xN);
      const double tpn = tc + a.tnbdy;
      hyadv = a.hdxm1 * (vp * tpn - vs * tms);
    } else {
      hyadv = a.hdxm1 * (-vs * tms);
    }
  } else {
    hyadv = a.hdxm1 * (vn * (tN + tc) - vs * (tc + tS));
  }
  double rhs = -(hxadv + hyadv);
  // dummy columns of del2t: no diffusive flux through solid W/E walls
  const double dw = wallW ? dcen : dW;
  const double de = wallE ? dcen : dE;
  double d4;
  if (gj == 0) d4 = dw + de + dN - 3.0 * dcen;
  else if (gj == nyt - 1) d4 = dS + dw + de - 3.0 * dcen;
  else d4 = dS + dw + de + dN - 4.0 * dcen;
  rhs = rhs + a.d2tfac * dcen - a.d4tfac * d4;
  const double diabat = 0.5 * wek * (tmc + a.toc1);
  double sstnew = tmc + a.tdt * (rhs + a.hmoinv * (a.rrcpoc * fnet + diabat));
  const double xfoent = -(0.5 * a.dtoinv) * wek * (tmc - a.toc1);
  const double dtonew = a.toc1 - sstnew;
  const double coneno = a.entfac * fmax(0.0, dtonew);
  OmlOut o;
  o.xf = xfoent - coneno;
  o.sst = sstnew + fmax(0.0, dtonew);
  o.cfr = (-dtonew >= 0.0) ? 0.0 : 1.0;   // 0.5 - sign(0.5, -dtonew)
  o.cen = -coneno;
  return o;
}

__device__ void oml_monitors(const OmlArgs &a, double (*red)[8]);

// grid (ceil(warps/4), ceil(nyt/MR)); each lane owns the even/odd column pair (g0, g0+1):
// 16-byte cp.async and 16-byte stores, one shuffle pair per field row for two cells
__global__ void __launch_bounds__(128) k_oml_march(OmlArgs a) {
  extern __shared__ double2 oml_ring[];
  __shared__ double red[3][4];
  const Grid &g = a.g;
  const int nxt = g.nxt, nyt = g.nyt, ld = g.ld, cyc = g.cyclic;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wx = blockIdx.x * 4 + wib;
  const bool active = wx * MW < nxt;                 // whole warp
  double2 *ring = oml_ring + (size_t)wib * (MD * MNF * 32) + lane;
  const int g0 = wx * MW - 2 + 2 * lane;             // first T column of this lane's pair (even)
  int c0 = g0;                                       // canonical column of the pair
  if (cyc) { if (c0 < 0) c0 += nxt; if (c0 >= nxt) c0 -= nxt; }
  // the pair is loaded when its first column exists; a second column beyond the row's end is padding
  const bool pairT = c0 >= 0 && c0 < nxt, pairP = cyc ? true : (c0 >= 0 && c0 <= nxt);
  const bool t0ok = pairT, t1ok = pairT && (cyc || g0 + 1 < nxt);
  const bool p1ok = cyc ? true : (pairP && g0 + 1 <= nxt);
  const bool out0 = lane >= 1 && lane < 31 && g0 < nxt, out1 = lane >= 1 && lane < 31 && g0 + 1 < nxt;
  const int ja = blockIdx.y * a.mrows, jb = min(nyt, ja + a.mrows);
  const int cc = (pairT || pairP) ? c0 : 0;
  double pxfo = 0.0, pcfr = 0.0, pcen = 0.0;
  if (active) {
#pragma unroll
    for (int s = 0; s < MD * MNF; ++s) ring[s * 32] = make_double2(0.0, 0.0);   // rows/columns outside the domain read as zero
    __syncwarp();
    // stage r carries sstm(r), sst(r-1), p/taux/tauy at p row r-1, wekt/fnet(r-2)
    auto issue = [&](int r) {
      double2 *slot = ring + (size_t)((r + 8 * MD) % MD) * (MNF * 32);
      if (r <= jb + 1) {     // not past the last row this march needs
        if (pairT && r >= 0 && r < nyt) oml_cp16(slot, a.sstm + (size_t)r * ld + cc);
        if (pairT && r - 1 >= 0 && r - 1 < nyt) oml_cp16(slot + 32, a.sst + (size_t)(r - 1) * ld + cc);
        if (pairP && r - 1 >= 0 && r - 1 <= nyt) {
          const size_t o = (size_t)(r - 1) * ld + cc;
          oml_cp16(slot + 64, a.po1 + o);
          oml_cp16(slot + 96, a.taux + o);
          oml_cp16(slot + 128, a.tauy + o);
        }
        if (pairT && r - 2 >= 0 && r - 2 < nyt) {
          const size_t o = (size_t)(r - 2) * ld + cc;
          oml_cp16(slot + 160, a.wekt + o);
          oml_cp16(slot + 192, a.fnet + o);
        }
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    // [.][0] / [.][1]: the lane's two columns
    double tm0[2] = {0, 0}, tm1[2] = {0, 0}, tm2[2] = {0, 0}, da[2] = {0, 0}, db[2] = {0, 0}, dc[2] = {0, 0};
    double tA[2] = {0, 0}, tB[2] = {0, 0}, tC[2] = {0, 0};
    double pS[2] = {0, 0}, pN[2] = {0, 0}, xS[2] = {0, 0}, xN[2] = {0, 0}, yS[2] = {0, 0}, yN[2] = {0, 0};
    double pSx = 0, pNx = 0, xSx = 0, xNx = 0, ySx = 0, yNx = 0;   // p-grid values at the column east of the pair (lane+1's first)
    const int r0 = ja - 2;
#pragma unroll
    for (int s = 0; s < MD - 1; ++s) issue(r0 + s);
    for (int r = r0; r <= jb + 1; ++r) {
      issue(r + MD - 1);
      asm volatile("cp.async.wait_group %0;\n" ::"n"(MD - 1) : "memory");
      const double2 *slot = ring + (size_t)((r + 8 * MD) % MD) * (MNF * 32);
      const bool in_tm = r >= 0 && r < nyt, in_t = r - 1 >= 0 && r - 1 < nyt, in_p = r - 1 >= 0 && r - 1 <= nyt;
      {
        const double2 vtm = slot[0], vt = slot[32], vp = slot[64], vx = slot[96], vy = slot[128];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tm0[c] = tm1[c]; tm1[c] = tm2[c];
          tA[c] = tB[c]; tB[c] = tC[c];
          pS[c] = pN[c]; xS[c] = xN[c]; yS[c] = yN[c];
        }
        pSx = pNx; xSx = xNx; ySx = yNx;
        tm2[0] = (in_tm && t0ok) ? vtm.x : 0.0; tm2[1] = (in_tm && t1ok) ? vtm.y : 0.0;
        tC[0] = (in_t && t0ok) ? vt.x : 0.0;    tC[1] = (in_t && t1ok) ? vt.y : 0.0;
        pN[0] = (in_p && pairP) ? vp.x : 0.0;   pN[1] = (in_p && p1ok) ? vp.y : 0.0;
        xN[0] = (in_p && pairP) ? vx.x : 0.0;   xN[1] = (in_p && p1ok) ? vx.y : 0.0;
        yN[0] = (in_p && pairP) ? vy.x : 0.0;   yN[1] = (in_p && p1ok) ? vy.y : 0.0;
        pNx = she(pN[0]); xNx = she(xN[0]); yNx = she(yN[0]);
      }
      // ---- del2t at T row r-1
      {
        const int gj = r - 1;
        const double w0 = shw(tm1[1]), e1 = she(tm1[0]);
        double v0 = 0.0, v1 = 0.0;
        if (gj >= 0 && gj < nyt) {
          if (t0ok) v0 = oml_del2(a, g0, gj, nxt, nyt, cyc, tm1[0], w0, tm1[1], tm0[0], tm2[0]);
          if (t1ok) v1 = oml_del2(a, g0 + 1, gj, nxt, nyt, cyc, tm1[1], tm1[0], e1, tm0[1], tm2[1]);
        }
        da[0] = db[0]; db[0] = dc[0]; dc[0] = v0;
        da[1] = db[1]; db[1] = dc[1]; dc[1] = v1;
      }
      // ---- T row r-2: the two cells of this lane
      const int gj = r - 2;
      const double tW0 = shw(tB[1]), tE1 = she(tB[0]), dW0 = shw(db[1]), dE1 = she(db[0]);
      if (gj < ja || gj >= jb || !out0) continue;
      const double2 wk = slot[160], fn = slot[192];
      const OmlOut o0 = oml_cell(a, g0, gj, nxt, nyt, cyc, tB[0], tW0, tB[1], tA[0], tC[0], db[0], dW0, db[1], da[0], dc[0], tm0[0],
                                 pS[0], pN[0], pS[1], pN[1], xS[0], xN[0], xS[1], xN[1], yS[0], yN[0], yS[1], yN[1], wk.x, fn.x);
      OmlOut o1 = o0;
      if (out1)
        o1 = oml_cell(a, g0 + 1, gj, nxt, nyt, cyc, tB[1], tB[0], tE1, tA[1], tC[1], db[1], db[0], dE1, da[1], dc[1], tm0[1],
                      pS[1], pN[1], pSx, pNx, xS[1], xN[1], xSx, xNx, yS[1], yN[1], ySx, yNx, wk.y, fn.y);
      const size_t idx = (size_t)gj * ld + g0;
      if (out1) {
        *reinterpret_cast<double2 *>(a.xfo + idx) = make_double2(o0.xf, o1.xf);
        *reinterpret_cast<double2 *>(a.sstnew + idx) = make_double2(o0.sst, o1.sst);
      } else {
        a.xfo[idx] = o0.xf;
        a.sstnew[idx] = o0.sst;
      }
      if (gj >= a.t0 && gj < a.t1) {            // halo rows of a slab belong to the neighbour's sums
        pxfo += o0.xf; pcfr += o0.cfr; pcen += o0.cen;
        if (out1) { pxfo += o1.xf; pcfr += o1.cfr; pcen += o1.cen; }
      }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  }
  // block partial sums (fixed order)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pxfo += __shfl_down_sync(0xffffffffu, pxfo, o);
    pcfr += __shfl_down_sync(0xffffffffu, pcfr, o);
    pcen += __shfl_down_sync(0xffffffffu, pcen, o);
  }
  if (lane == 0) { red[0][wib] = pxfo; red[1][wib] = pcfr; red[2][wib] = pcen; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int i = 0; i < 4; ++i) t += red[threadIdx.x][i];
    a.part[(size_t)threadIdx.x * a.nblocks + blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
  // The block that finishes last adds the block partials in a fixed order (thread t takes partials
  // t, t+128, ...; shuffle tree; the four warp sums in index order), evaluates the boundary
  // monitors and, on y-slabs over peer memory, all-reduces the three sums with the other ranks:
  // what used to be a second, latency-bound launch (k_oml_reduce).
  __shared__ bool last;
  __shared__ double mred[6][8];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(a.ticket + 2, 1u) == (unsigned int)a.nblocks - 1u);
  __syncthreads();
  if (!last) return;
  __threadfence();
  double s3[3] = {0.0, 0.0, 0.0};
  for (int q = 0; q < 3; ++q)
    for (int i = threadIdx.x; i < a.nblocks; i += 128) s3[q] += __ldcg(a.part + (size_t)q * a.nblocks + i);
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s3[q] += __shfl_down_sync(0xffffffffu, s3[q], o);
    if (lane == 0) mred[q][wib] = s3[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.ticket[2] = 0u;
    double t[3];
    for (int q = 0; q < 3; ++q) t[q] = (mred[q][0] + mred[q][1]) + (mred[q][2] + mred[q][3]);
    a.cv[0] = t[0];                         // xfosum, read by k_oml_entoc (after the all-reduce on slabs)
    a.cv[1] = t[1];
    a.cv[2] = t[2];
    if (!a.multi) {
      a.sc->cfraoc = t[1] * a.g.norm;       // omlsubs.F:211
      a.sc->centoc = t[2] * (a.g.dx * a.g.dx);   // omlsubs.F:212
    }
  }
  if (a.sb || a.nb) {
    __syncthreads();
    oml_monitors(a, mred);                  // they only read the old fields
  }
  if (a.peer.n) {       // the slab ranks' sums meet here, over NVLink peer memory
    __syncthreads();
    peer_allreduce_block(a.peer, a.cv, 3, a.cv, a.peer_err);
  }
}

// boundary-flux monitors of the sb_hflux / nb_hflux options (omlsubs.F:684-726); called by
// one whole block of 256 threads.  They read the old sst/sstm, which the step leaves intact.
__device__ void oml_monitors(const OmlArgs &a, double (*red)[8]) {
  const Grid &g = a.g;
  const int nxt = g.nxt, nyt = g.nyt, ld = g.ld;
  double s[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 4
  for (int i = threadIdx.x; i < nxt; i += blockDim.x) {
    if (a.sb && g.wall_s()) {   // y-slabs: the rank that holds the wall owns these monitors
      const double vm = -a.rhf0hm * (a.taux[i + 1] + a.taux[i]);
      const double tm = a.sst[i] + a.tsbdy;
      s[0] += vm; s[1] += vm * tm; s[2] -= (a.sstm[i] - a.tsbdy);
    }
    if (a.nb && g.wall_n()) {
      const double vp = -a.rhf0hm * (a.taux[(size_t)nyt * ld + i + 1] + a.taux[(size_t)nyt * ld + i]);
      const double tp = a.sst[(size_t)(nyt - 1) * ld + i] + a.tnbdy;
      s[3] -= vp; s[4] -= vp * tp; s[5] += (a.tnbdy - a.sstm[(size_t)(nyt - 1) * ld + i]);
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_down_sync(0xffffffffu, s[q], o);
    if (lane == 0) red[q][w] = s[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[6];
    const int nw = blockDim.x >> 5;
    for (int q = 0; q < 6; ++q) { t[q] = 0.0; for (int i = 0; i < nw; ++i) t[q] += red[q][i]; }
    const double n = (double)nxt;
    a.sc->vfmads = t[0] / n; a.sc->ttmads = a.hdxm1 * t[1] / n; a.sc->ttmdfs = a.d2tfac * t[2] / n;
    a.sc->vfmadn = t[3] / n; a.sc->ttmadn = a.hdxm1 * t[4] / n; a.sc->ttmdfn = a.d2tfac * t[5] / n;
  }
}

__device__ void oml_finish(const OmlArgs &a, double dx, double *red);

// entoc = 4-point average of (xfo - mean) with the edge/corner rules (omlsubs.F:151-205);
// one block per p row; also the xintp row sum of that row (intsubs.f:103-131)
__global__ void __launch_bounds__(256) k_oml_entoc(OmlArgs a) {
  __shared__ double red[8];
  const Grid &g = a.g;
  const int j = blockIdx.x;   // 0-based p row
  const int nxp = g.nxp, nyp = g.nyp, nxt = g.nxt, nyt = g.nyt, ld = g.ld, cyc = g.cyclic;
  const double mean = a.cv[0] * g.norm;   // xfosum*ocnorm
  const bool rowS = (j == 0), rowN = (j == nyp - 1);
  const int jm = j - 1, jc = j;
#define X(ii, jj) (a.xfo[(size_t)(jj) * ld + (ii)] - mean)
  // one p point with every edge/corner rule
  auto point = [&](int i) {
    // T cells around p point (i,j): (i-1,j-1), (i,j-1), (i-1,j), (i,j) in 0-based T indices
    int im = i - 1, ic = i;
    const bool colW = (i == 0), colE = (i == nxp - 1);
    if (cyc && (colW || colE)) { im = nxt - 1; ic = 0; }
    if (!rowS && !rowN) {
      if (!cyc && colW) return 0.5 * (X(0, jm) + X(0, jc));
      if (!cyc && colE) return 0.5 * (X(nxt - 1, jm) + X(nxt - 1, jc));
      return 0.25 * (X(im, jm) + X(ic, jm) + X(im, jc) + X(ic, jc));
    }
    const int jt = rowS ? 0 : nyt - 1;
    if (!cyc && colW) return X(0, jt);
    if (!cyc && colE) return X(nxt - 1, jt);
    return 0.5 * (X(im, jt) + X(ic, jt));
  };
  double part = 0.0;
  // two p columns per thread: 16-byte loads of the T rows and a 16-byte store
#pragma unroll 2
  for (int i0 = 2 * threadIdx.x; i0 < nxp; i0 += 512) {
    const int i1 = i0 + 1;
    double v0, v1 = 0.0;
    if (!rowS && !rowN && i0 >= 2 && i1 <= nxp - 2) {
      const double2 am = *reinterpret_cast<const double2 *>(a.xfo + (size_t)jm * ld + i0);
      const double2 ac = *reinterpret_cast<const double2 *>(a.xfo + (size_t)jc * ld + i0);
      const double wm = X(i0 - 1, jm), wc = X(i0 - 1, jc);
      const double m0 = am.x - mean, m1 = am.y - mean, c0 = ac.x - mean, c1 = ac.y - mean;
      v0 = 0.25 * (wm + m0 + wc + c0);
      v1 = 0.25 * (m0 + m1 + c0 + c1);
      *reinterpret_cast<double2 *>(a.entoc + (size_t)j * ld + i0) = make_double2(v0, v1);
      part += v0 + v1;
    } else {
      v0 = point(i0);
      a.entoc[(size_t)j * ld + i0] = v0;
      part += (i0 == 0 || i0 == nxp - 1) ? 0.5 * v0 : v0;
      if (i1 < nxp) {
        v1 = point(i1);
        a.entoc[(size_t)j * ld + i1] = v1;
        part += (i1 == nxp - 1) ? 0.5 * v1 : v1;
      }
    }
  }
#undef X
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  if (lane == 0) red[w] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    a.rowsum[j] = t;
  }
  // the block that finishes last forms xon(1) from the row sums (what used to be k_oml_finish)
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(a.ticket + 1, 1u) == gridDim.x - 1u);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x == 0) a.ticket[1] = 0u;
  oml_finish(a, a.g.dx, red);
}

// xon(1) = dx*dy*xintp(entoc); channel: enisoc(1), eninoc(1) (omlsubs.F:214-233); one block of 256 threads
__device__ void oml_finish(const OmlArgs &a, double dx, double *red) {
  const int nyp = a.g.nyp;
  if (a.multi) {
    // owned rows only; the rows on the real walls carry half weight (intsubs.f:120-131)
    const int lo = a.p0 + (a.g.wall_s() ? 1 : 0), hi = a.p1 - (a.g.wall_n() ? 1 : 0);
    const double sump = block256_range_sum(a.rowsum, lo, hi, red);
    if (threadIdx.x != 0) return;
    double x = sump;
    if (a.g.wall_s()) x += 0.5 * a.rowsum[0];
    if (a.g.wall_n()) x += 0.5 * a.rowsum[nyp - 1];
    a.cv[3] = x * dx * dx;
    a.sc->cfraoc = a.cv[1] * a.g.norm;
    a.sc->centoc = a.cv[2] * dx * dx;
    return;
  }
  const double sump = block256_range_sum(a.rowsum, 1, nyp - 1, red);
  if (threadIdx.x != 0) return;
  const double x = sump + 0.5 * (a.rowsum[0] + a.rowsum[nyp - 1]);
  a.sc->xon[0] = x * dx * dx;
  if (a.g.cyclic) {
    a.sc->enisoc[0] = dx * a.rowsum[0];
    a.sc->eninoc[0] = dx * a.rowsum[nyp - 1];
  }
}

static OmlArgs oml_args(qgcm_model *m, dim3 &grid) {
  OmlArgs a;
  const Grid &g = m->go;
  const qgcm_config &c = m->cfg;
  a.g = g;
  a.sb = m->sb_hflux; a.nb = m->nb_hflux;
  a.tsbdy = c.tsbdy; a.tnbdy = c.tnbdy;
  a.uvgfac = c.ycexp * g.rdxf0;
  a.rhf0hm = 0.5 / (m->fnot * c.hmoc);
  a.d2tfac = c.st2d * g.dxm2;
  a.d4tfac = c.st4d * g.dxm2 * g.dxm2;
  a.hdxm1 = g.hdxm1;
  a.hmoinv = 1.0 / c.hmoc;
  a.dtoinv = 1.0 / (c.toc[0] - c.toc[1]);
  a.entfac = c.hmoc * a.dtoinv / g.tdt;
  a.rrcpoc = m->rrcpoc;
  a.toc1 = c.toc[0];
  a.tdt = g.tdt;
  a.po1 = m->F("po"); a.taux = m->F("tauxo"); a.tauy = m->F("tauyo");
  a.sst = m->F("sst"); a.sstm = m->F("sstm"); a.wekt = m->F("wekto"); a.fnet = m->F("fnetoc");
  a.sstnew = m->sstnew; a.xfo = m->xfo;
  a.mrows = MR;
  a.peer = PeerCtx{};
  a.peer_err = m->d_peer_err;
  {
    // marches sized to whole waves of resident blocks, at least 16 rows each (4 fill rows per march)
    const int xw = (g.nxt + MW - 1) / MW;
    int &resident = m->oml_resident;
    if (!resident) {
      const size_t smem = 4 * MD * MNF * 32 * sizeof(double2);
      QG_CUDA(cudaFuncSetAttribute(k_oml_march, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0, sms = 0;
      QG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_oml_march, 128, smem));
      QG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->cfg.device));
      resident = std::max(1, per_sm * sms);
    }
    a.mrows = pick_march_rows(g.nyt, (xw + 3) / 4, resident, 16, MR);
    a.mrows = std::max(8, std::min(1024, env_int("QGCM_OML_MROWS", a.mrows)));
    grid = dim3((xw + 3) / 4, (g.nyt + a.mrows - 1) / a.mrows);
  }
  a.nblocks = grid.x * grid.y;
  a.part = m->d_red;
  a.rowsum = m->d_red + 3 * (size_t)a.nblocks;
  a.part2 = a.rowsum + g.nyp;
  a.ticket = m->d_ticket;
  a.entoc = m->F("entoc");
  a.sc = m->d_scal;
  a.multi = m->nranks > 1;
  a.p0 = g.own0; a.p1 = g.own1;
  a.t0 = g.own0; a.t1 = std::min(g.own1, g.nyt);
  a.cv = m->d_cv;
  if (m->red_elems < 3 * (size_t)a.nblocks + g.nyp + 3 * 64) throw std::runtime_error("oml: reduction scratch too small");
  return a;
}

// monitors, the fused advection/diffusion/entrainment step, and its three sums (local to the
// rank on y-slabs: d_cv[0..2] are all-reduced before phase b)
void oml_phase_a(qgcm_model *m) {
  dim3 grid;
  OmlArgs a = oml_args(m, grid);
  const Grid &g = m->go;
  a.peer = peer_next_vec(m);    // y-slabs over peer memory: the last block of the march all-reduces the sums itself
  const size_t smem = 4 * MD * MNF * 32 * sizeof(double2);   // attribute set per model in oml_args
  QG_LAUNCH(m, "k_oml_step", grid, 128, smem, k_oml_march, a);
  (void)g;
}

// entoc from xfo minus the global mean, its integral (y-slabs: the rank's share in d_cv[3])
void oml_phase_b(qgcm_model *m) {
  dim3 grid;
  OmlArgs a = oml_args(m, grid);
  const Grid &g = m->go;
  QG_LAUNCH(m, "k_oml_entoc", g.nyp, 256, 0, k_oml_entoc, a);      // its last block also forms xon(1)
  QG_CUDA(cudaGetLastError());
  // sstm <- sst, sst <- new: three-buffer rotation (omlsubs.F:124-125)
  double *old_m = m->fields.at("sstm").d;
  m->fields.at("sstm").d = m->fields.at("sst").d;
  m->fields.at("sst").d = m->sstnew;
  m->sstnew = old_m;
}

void launch_oml(qgcm_model *m) {
  if (m->nranks > 1) throw std::runtime_error("qgcm_oml: a y-slab model is stepped with qgcm_ocean_step");
  oml_phase_a(m);
  oml_phase_b(m);
}

}  // namespace qg

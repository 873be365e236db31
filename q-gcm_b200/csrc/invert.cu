// PV inversion around the Helmholtz solver: ocinvq (src/ocisubs.F:64-407) and atinvq
// (src/atisubs.F:60-293): layer->mode RHS, the constraint algebra on the device-resident
// scalars, homogeneous-solution add and mode->layer projection; plus the init-time
// producers that use the same solver: homsol and constr (src/conhoms.F:44-818).
#include "qgcm_internal.h"

namespace qg {

struct InvArgs {
  Grid g;
  int atmos, nl;
  double f0, beta;
  double ctl2m[NLMAX * NLMAX], ctm2l[NLMAX * NLMAX];
  const double *q, *ddyn, *yrel;
  double *wrk;
  const double *hom;        // box: ochom [nl-1][nyp][ld]
  const double *pch1, *pch2, *pbh;   // channel: [nl-1][nyp], [nyp]
  double *pnew;             // written over the lagged-p buffer
  const double *coef;       // device: box hclco[nl-1]; channel c3, c1[nl-1], c2[nl-1]
};

// wrk_m = f0 * sum_k ctl2m(k,m) (q_k - beta*y - [k==kbot] ddyn), rows 2..nyp-1, all i
// (src/ocisubs.F:117-139, src/atisubs.F:106-126)
__global__ void __launch_bounds__(256) k_l2m(InvArgs a) {
  const Grid &g = a.g;
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);   // two columns per thread: 16-byte accesses
  const int j = blockIdx.y + 1;   // 0-based interior row
  if (i >= g.nxp) return;
  const double betay = a.beta * a.yrel[j];
  const size_t idx = (size_t)j * g.ld + i;     // rows are 128-byte aligned and ld >= nxp + 1 when nxp is odd
  const int nl = a.nl, kbot = a.atmos ? 0 : nl - 1;
  double2 ql[NLMAX];
  for (int k = 0; k < nl; ++k) {
    const double2 q = *reinterpret_cast<const double2 *>(a.q + k * g.lsz + idx);
    ql[k] = make_double2(q.x - betay, q.y - betay);
  }
  if (a.ddyn) {      // null over a flat bottom: x - 0 is x, so the field need not be read
    const double2 dd = *reinterpret_cast<const double2 *>(a.ddyn + idx);
    ql[kbot].x = ql[kbot].x - dd.x;
    ql[kbot].y = ql[kbot].y - dd.y;
  }
  for (int m = 0; m < nl; ++m) {
    double qx = 0.0, qy = 0.0;
    for (int k = 0; k < nl; ++k) {
      qx = qx + a.ctl2m[k + nl * m] * ql[k].x;
      qy = qy + a.ctl2m[k + nl * m] * ql[k].y;
    }
    // the pad column beyond nxp (odd nxp) receives a value nobody reads
    *reinterpret_cast<double2 *>(a.wrk + m * g.lsz + idx) = make_double2(a.f0 * qx, a.f0 * qy);
  }
}

// p_k = sum_m ctm2l(m,k) (wrk_m + homogeneous_m) at every point
// (src/ocisubs.F:300-327 channel, :377-401 box; src/atisubs.F:264-291)
__global__ void __launch_bounds__(256) k_m2l(InvArgs a) {
  const Grid &g = a.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= g.nxp) return;
  const size_t idx = (size_t)j * g.ld + i;
  const int nl = a.nl;
  double pm[NLMAX];
  if (g.cyclic) {
    pm[0] = a.wrk[idx] + a.coef[0] * a.pbh[j];
    for (int m = 1; m < nl; ++m) {
      const double homcor = a.coef[m] * a.pch1[(m - 1) * g.nyp + j] + a.coef[nl - 1 + m] * a.pch2[(m - 1) * g.nyp + j];
      pm[m] = a.wrk[m * g.lsz + idx] + homcor;
    }
  } else {
    pm[0] = a.wrk[idx];
    for (int m = 1; m < nl; ++m) pm[m] = a.wrk[m * g.lsz + idx] + a.coef[m - 1] * a.hom[(m - 1) * g.lsz + idx];
  }
  for (int k = 0; k < nl; ++k) {
    double pl = 0.0;
    for (int m = 0; m < nl; ++m) pl = pl + a.ctm2l[m + nl * k] * pm[m];
    a.pnew[k * g.lsz + idx] = pl;
  }
}

// dense solve with partial pivoting plus one refinement sweep: DGETRS + DGERFS at
// src/ocisubs.F:359-370 (LAPACK is not vendored by the reference; published algorithm)
__device__ void lu_solve_refine(const double *a, int n, const double *rhs, double *x) {
  double lu[NLMAX * NLMAX];
  int piv[NLMAX];
  for (int i = 0; i < n * n; ++i) lu[i] = a[i];
  for (int kk = 0; kk < n; ++kk) {
    int p = kk;
    for (int i = kk + 1; i < n; ++i)
      if (fabs(lu[i + n * kk]) > fabs(lu[p + n * kk])) p = i;
    piv[kk] = p;
    if (p != kk)
      for (int j = 0; j < n; ++j) { double t = lu[kk + n * j]; lu[kk + n * j] = lu[p + n * j]; lu[p + n * j] = t; }
    for (int i = kk + 1; i < n; ++i) {
      lu[i + n * kk] /= lu[kk + n * kk];
      for (int j = kk + 1; j < n; ++j) lu[i + n * j] -= lu[i + n * kk] * lu[kk + n * j];
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    double b[NLMAX];
    if (pass == 0) {
      for (int i = 0; i < n; ++i) b[i] = rhs[i];
    } else {
      for (int i = 0; i < n; ++i) {
        double acc = rhs[i];
        for (int j = 0; j < n; ++j) acc -= a[i + n * j] * x[j];
        b[i] = acc;
      }
    }
    for (int kk = 0; kk < n; ++kk) { double t = b[kk]; b[kk] = b[piv[kk]]; b[piv[kk]] = t; }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) b[i] -= lu[i + n * j] * b[j];
    for (int i = n - 1; i >= 0; --i) {
      for (int j = i + 1; j < n; ++j) b[i] -= lu[i + n * j] * b[j];
      b[i] /= lu[i + n * i];
    }
    for (int i = 0; i < n; ++i) x[i] = pass == 0 ? b[i] : x[i] + b[i];
  }
}

// the same with a compile-time order: every index is a constant after unrolling (row exchanges become
// conditional swaps), so the factors live in registers instead of local memory; same operations, same order
template <int N>
__device__ __forceinline__ void lu_solve_refine_t(const double *a, const double *rhs, double *x) {
  double lu[N * N];
  int piv[N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) lu[i] = a[i];
#pragma unroll
  for (int kk = 0; kk < N; ++kk) {
    int p = kk;
    double best = fabs(lu[kk + N * kk]);
#pragma unroll
    for (int i = kk + 1; i < N; ++i)
      if (fabs(lu[i + N * kk]) > best) { p = i; best = fabs(lu[i + N * kk]); }
    piv[kk] = p;
#pragma unroll
    for (int i = kk + 1; i < N; ++i)
      if (p == i) {
#pragma unroll
        for (int j = 0; j < N; ++j) { const double t = lu[kk + N * j]; lu[kk + N * j] = lu[i + N * j]; lu[i + N * j] = t; }
      }
#pragma unroll
    for (int i = kk + 1; i < N; ++i) {
      lu[i + N * kk] /= lu[kk + N * kk];
#pragma unroll
      for (int j = kk + 1; j < N; ++j) lu[i + N * j] -= lu[i + N * kk] * lu[kk + N * j];
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    double b[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = rhs[i];
      if (pass == 1) {
#pragma unroll
        for (int j = 0; j < N; ++j) acc -= a[i + N * j] * x[j];
      }
      b[i] = acc;
    }
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
#pragma unroll
      for (int i = kk + 1; i < N; ++i)
        if (piv[kk] == i) { const double t = b[kk]; b[kk] = b[i]; b[i] = t; }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int j = 0; j < i; ++j) b[i] -= lu[i + N * j] * b[j];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
#pragma unroll
      for (int j = i + 1; j < N; ++j) b[i] -= lu[i + N * j] * b[j];
      b[i] /= lu[i + N * i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = pass == 0 ? b[i] : x[i] + b[i];
  }
}

struct ScalArgs {
  int atmos, cyclic, nl, nyp;
  double dx, f0, tdt, xl, yl;
  double h[NLMAX], gp[NLMAX], ctl2m[NLMAX * NLMAX], ctm2l[NLMAX * NLMAX];
  const double *rowsum;   // [nl][nyp]
  // where the modal area integrals come from: xintp row sums of the inverse transform
  // (sumsrc = rowsum, stride nyp, rows [sumlo, sumhi)) or, on the fused box path, the per-block
  // spectral shares left by k_tri3 (sumsrc = spec, stride nspec, [0, nspec))
  const double *sumsrc;
  int sumstride, sumlo, sumhi;
  qgcm_scalars *sc;
  double *coef;
  // y-slabs (box ocean): cv[3] = xon(1), cv[4+m] = xinhom(m), already summed over the ranks
  const double *cv;
  // y-slabs over peer memory: this kernel also forms the rank's share of xinhom(m) from the
  // solved rows [lo, hi) and sums cv[3 .. 4+nl) over the ranks itself (k_inv_partials and the
  // all-reduce that would follow it are folded in)
  PeerCtx peer;
  int *peer_err;
  double *cvw;
  int lo, hi;
};

// y-slabs: this rank's share of the xintp integrals of the modal solutions -> cv[4+m]
// (with the peer-memory transport it also sums cv[3 .. 4+nl) over the ranks)
__global__ void __launch_bounds__(256) k_inv_partials(const double *sumsrc, int nl, int stride, int lo, int hi, double dx, double *cv,
                                                      PeerCtx peer, int *peer_err) {
  __shared__ double red[8];
  for (int m = 0; m < nl; ++m) {
    const double s = block256_range_sum(sumsrc + (size_t)m * stride, lo, hi, red);   // wall rows are exactly zero
    if (threadIdx.x == 0) cv[4 + m] = s * dx * dx;
    __syncthreads();
  }
  if (peer.n) peer_allreduce_block(peer, cv + 3, 1 + nl, cv + 3, peer_err);
}

// Single-thread constraint algebra on device-resident scalars, so the step never
// synchronises with the host (src/ocisubs.F:146-162, :174-294, :333-370;
// src/atisubs.F:137-258).  xinhom(m) = dx*dy * sum of the xintp row sums.
// (One block of INV_NT threads.  The kernel is a chain of dependent round trips to L2 -- the scalar state, the
// partial sums, the algebra's operands -- so everything it reads is fetched by all threads at once: the
// scalar block into shared memory, the partial sums eight strides at a time.)
constexpr int INV_NT = 512;

// tot[m] = sum of src[m*stride + lo .. hi): fixed strided partials (thread t takes lo+t, lo+t+INV_NT, ... in
// increasing order), fixed-order tree; the loads of eight strides are in flight together
__device__ __forceinline__ void inv_mode_sums(const double *src, int stride, int lo, int hi, int nl, double (*redm)[INV_NT / 32], double *tot) {
  constexpr int U = 8;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int m = 0; m < nl; ++m) {
    const double *v = src + (size_t)m * stride;
    double acc = 0.0;
    for (int i0 = lo + (int)threadIdx.x; i0 < hi; i0 += U * INV_NT) {
      double x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = (i0 + u * INV_NT < hi) ? v[i0 + u * INV_NT] : 0.0;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u * INV_NT < hi) acc += x[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) redm[m][w] = acc;
  }
  __syncthreads();
  if ((int)threadIdx.x < nl) {
    double t = 0.0;
    for (int i = 0; i < INV_NT / 32; ++i) t += redm[threadIdx.x][i];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
}

// The constraint algebra itself, one thread (src/ocisubs.F:146-162, :174-294, :333-370; src/atisubs.F:137-258).
// NLC = the layer count as a compile-time constant (2, 3, 4: every loop unrolls and every small array lives in
// registers) or 0 = read it from the arguments (any count up to NLMAX; the arrays then live in local memory,
// and the thread spends most of its 15 microseconds waiting on them).  Same statements either way.
template <int NLC>
__device__ __forceinline__ void inv_algebra(const ScalArgs &a, qgcm_scalars *s, const qgcm_scalars *ls, const double *tot) {
  constexpr int NA = NLC ? NLC : NLMAX;
  const int nl = NLC ? NLC : a.nl, nyp = a.nyp;
  const double ecrit = 1.0e-13;
  double xinhom[NA], ayis[NA], ayin[NA];
  double sums[NA];
  for (int m = 0; m < nl; ++m) sums[m] = a.cv ? 0.0 : tot[m];
  // box: the scalar state the constraint algebra needs
  double dpi0[NA], dpip0[NA], cdf[NA * NA], cdh[NA * NA], xon_own = 0.0;
  if (!a.cyclic) {
    xon_own = ls->xon[0];
    for (int k = 0; k < nl - 1; ++k) { dpi0[k] = ls->dpioc[k]; dpip0[k] = ls->dpiocp[k]; }
    for (int i = 0; i < nl * (nl - 1); ++i) cdf[i] = ls->cdiffo[i];
    for (int i = 0; i < (nl - 1) * (nl - 1); ++i) cdh[i] = ls->cdhoc[i];
  }
  if (a.cv) s->xon[0] = a.cv[3];
  for (int m = 0; m < nl; ++m) {
    const double sump = sums[m];
    xinhom[m] = a.cv ? a.cv[4 + m] : sump * a.dx * a.dx;   // boundary rows are exactly zero
    if (a.cyclic) {
      ayis[m] = a.rowsum[m * nyp + 1];              // dx/dy = 1
      ayin[m] = -a.rowsum[m * nyp + nyp - 2];
    }
    if (a.atmos) s->xinhom_at[m] = xinhom[m]; else s->xinhom_oc[m] = xinhom[m];
  }
  if (!a.cyclic) {
    // finite box: mass constraints (src/ocisubs.F:333-370)
    double aient[NA], rhs[NA], hclco[NA];
    aient[0] = a.cv ? a.cv[3] : xon_own;
    for (int k = 1; k < nl - 1; ++k) aient[k] = 0.0;
    for (int k = 0; k < nl - 1; ++k) {
      const double aitmp = dpi0[k];
      const double dnew = dpip0[k] - a.tdt * a.gp[k] * aient[k];
      s->dpioc[k] = dnew;
      s->dpiocp[k] = aitmp;
      double rhsum = 0.0;
      for (int m = 0; m < nl; ++m) rhsum = rhsum + cdf[m + nl * k] * xinhom[m];
      rhs[k] = dnew - rhsum;
    }
    if (NLC > 1) lu_solve_refine_t<(NLC > 1 ? NLC - 1 : 1)>(cdh, rhs, hclco);
    else lu_solve_refine(cdh, nl - 1, rhs, hclco);
    for (int k = 0; k < nl - 1; ++k) a.coef[k] = hclco[k];
    return;
  }
  // periodic channel: momentum constraints
  double rhss[NA], rhsn[NA], snew[NA], nnew[NA], clhss[NA], clhsn[NA];
  double c1[NA], c2[NA], c3, aipmod[NA], aiplay[NA];
  const double entfac = 0.5 * a.dx * a.f0 * a.f0;
  const double *h = a.h;
  if (!a.atmos) {
    rhss[0] = (entfac / h[0]) * ls->enisoc[0] + (a.f0 / h[0]) * ls->txisoc + ls->ajisoc[0] - ls->ap3soc[0] + ls->ap5soc[0];
    rhsn[0] = (entfac / h[0]) * ls->eninoc[0] - (a.f0 / h[0]) * ls->txinoc + ls->ajinoc[0] + ls->ap3noc[0] - ls->ap5noc[0];
    for (int k = 1; k < nl - 1; ++k) {
      rhss[k] = (entfac / h[k]) * (ls->enisoc[k] - ls->enisoc[k - 1]) + ls->ajisoc[k] - ls->ap3soc[k] + ls->ap5soc[k];
      rhsn[k] = (entfac / h[k]) * (ls->eninoc[k] - ls->eninoc[k - 1]) + ls->ajinoc[k] + ls->ap3noc[k] - ls->ap5noc[k];
    }
    rhss[nl - 1] = -(entfac / h[nl - 1]) * ls->enisoc[nl - 2] + ls->ajisoc[nl - 1] - ls->ap3soc[nl - 1] + ls->ap5soc[nl - 1] +
                   (a.f0 / h[nl - 1]) * ls->bdrins;
    rhsn[nl - 1] = -(entfac / h[nl - 1]) * ls->eninoc[nl - 2] + ls->ajinoc[nl - 1] + ls->ap3noc[nl - 1] - ls->ap5noc[nl - 1] -
                   (a.f0 / h[nl - 1]) * ls->bdrinn;
  } else {
    rhss[0] = -(entfac / h[0]) * ls->enisat[0] - (a.f0 / h[0]) * ls->txisat + ls->ajisat[0] + ls->ap5sat[0];
    rhsn[0] = -(entfac / h[0]) * ls->eninat[0] + (a.f0 / h[0]) * ls->txinat + ls->ajinat[0] - ls->ap5nat[0];
    for (int k = 1; k < nl - 1; ++k) {
      rhss[k] = -(entfac / h[k]) * (ls->enisat[k] - ls->enisat[k - 1]) + ls->ajisat[k] + ls->ap5sat[k];
      rhsn[k] = -(entfac / h[k]) * (ls->eninat[k] - ls->eninat[k - 1]) + ls->ajinat[k] - ls->ap5nat[k];
    }
    rhss[nl - 1] = (entfac / h[nl - 1]) * ls->enisat[nl - 2] + ls->ajisat[nl - 1] + ls->ap5sat[nl - 1];
    rhsn[nl - 1] = (entfac / h[nl - 1]) * ls->eninat[nl - 2] + ls->ajinat[nl - 1] - ls->ap5nat[nl - 1];
  }
  double *cs = a.atmos ? s->atmcs : s->ocncs, *cn = a.atmos ? s->atmcn : s->ocncn;
  double *csp = a.atmos ? s->atmcsp : s->ocncsp, *cnp = a.atmos ? s->atmcnp : s->ocncnp;
  const double *cs0 = a.atmos ? ls->atmcs : ls->ocncs, *cn0 = a.atmos ? ls->atmcn : ls->ocncn;
  const double *csp0 = a.atmos ? ls->atmcsp : ls->ocncsp, *cnp0 = a.atmos ? ls->atmcnp : ls->ocncnp;
  for (int k = 0; k < nl; ++k) {
    snew[k] = csp0[k] + a.tdt * rhss[k];
    nnew[k] = cnp0[k] + a.tdt * rhsn[k];
    csp[k] = cs0[k];
    cnp[k] = cn0[k];
    cs[k] = snew[k];
    cn[k] = nnew[k];
  }
  for (int m = 0; m < nl; ++m) {
    clhss[m] = 0.0;
    clhsn[m] = 0.0;
    for (int k = 0; k < nl; ++k) {
      clhss[m] = clhss[m] + a.ctl2m[k + nl * m] * snew[k];
      clhsn[m] = clhsn[m] + a.ctl2m[k + nl * m] * nnew[k];
    }
    clhss[m] = clhss[m] + ayis[m];
    clhsn[m] = clhsn[m] - ayin[m];
  }
  const double *hc1s = a.atmos ? ls->hc1sat : ls->hc1soc, *hc2s = a.atmos ? ls->hc2sat : ls->hc2soc;
  const double *hc1n = a.atmos ? ls->hc1nat : ls->hc1noc, *hc2n = a.atmos ? ls->hc2nat : ls->hc2noc;
  const double *aipch = a.atmos ? ls->aipcha : ls->aipcho;
  const double hbsi = a.atmos ? ls->hbsiat : ls->hbsioc, aipbh = a.atmos ? ls->aipbha : ls->aipbho;
  c3 = clhss[0] * hbsi;
  for (int m = 0; m < nl - 1; ++m) {
    c1[m] = hc2n[m] * clhss[m + 1] - hc2s[m] * clhsn[m + 1];
    c2[m] = hc1s[m] * clhsn[m + 1] - hc1n[m] * clhss[m + 1];
  }
  aipmod[0] = xinhom[0] + c3 * aipbh;
  for (int m = 1; m < nl; ++m) aipmod[m] = xinhom[m] + (c1[m - 1] + c2[m - 1]) * aipch[m - 1];
  for (int k = 0; k < nl; ++k) {
    double pl = 0.0;
    for (int m = 0; m < nl; ++m) pl = pl + a.ctm2l[m + nl * k] * aipmod[m];
    aiplay[k] = pl;
  }
  double *dpi = a.atmos ? s->dpiat : s->dpioc, *dpip = a.atmos ? s->dpiatp : s->dpiocp;
  const double *dpi0c = a.atmos ? ls->dpiat : ls->dpioc, *dpip0c = a.atmos ? ls->dpiatp : ls->dpiocp;
  const double *xn = a.atmos ? ls->xan : ls->xon;
  double *erma = a.atmos ? s->ermasa : s->ermaso, *emfr = a.atmos ? s->emfrat : s->emfroc;
  for (int k = 0; k < nl - 1; ++k) {
    // sign conventions: ocean dpioc = p(k+1)-p(k) (ocisubs.F:272), atmosphere p(k)-p(k+1) (atisubs.F:237)
    const double est1 = a.atmos ? aiplay[k] - aiplay[k + 1] : aiplay[k + 1] - aiplay[k];
    const double est2 = dpip0c[k] - a.tdt * a.gp[k] * xn[k];
    const double edif = est1 - est2;
    const double esum = fabs(est1) + fabs(est2);
    erma[k] = edif;
    emfr[k] = (esum > (ecrit * a.xl * a.yl * a.tdt * a.gp[k])) ? 2.0 * edif / esum : 0.0;
    dpip[k] = dpi0c[k];
    dpi[k] = est1;
  }
  a.coef[0] = c3;
  for (int m = 1; m < nl; ++m) {
    a.coef[m] = c1[m - 1];
    a.coef[nl - 1 + m] = c2[m - 1];
  }
}

__global__ void __launch_bounds__(INV_NT) k_inv_scalars(ScalArgs a) {
  // the fused inverse transform that follows needs this kernel's result only in its epilogues: let it start
  asm volatile("griddepcontrol.launch_dependents;");
  __shared__ qgcm_scalars sh;      // snapshot of the scalar block: every read below comes from it, every write goes to *s
  __shared__ double redm[NLMAX][INV_NT / 32];
  __shared__ double tot[NLMAX];
  qgcm_scalars *s = a.sc;
  const qgcm_scalars *ls = &sh;
  {
    static_assert(sizeof(qgcm_scalars) % 8 == 0, "qgcm_scalars is copied in 8-byte words");
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(a.sc);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(&sh);
    for (int i = threadIdx.x; i < (int)(sizeof(qgcm_scalars) / 8); i += INV_NT) dst[i] = src[i];
  }
  const int nl = a.nl;
  if (a.peer.n) {
    // y-slabs over peer memory: this rank's share of the modal integrals, then their sum over the ranks
    inv_mode_sums(a.sumsrc, a.sumstride, a.sumlo, a.sumhi, nl, redm, tot);
    if ((int)threadIdx.x < nl) a.cvw[4 + threadIdx.x] = tot[threadIdx.x] * a.dx * a.dx;
    __syncthreads();
    peer_allreduce_block(a.peer, a.cvw + 3, 1 + nl, a.cvw + 3, a.peer_err);
  }
  if (!a.cv) inv_mode_sums(a.sumsrc, a.sumstride, a.sumlo, a.sumhi, nl, redm, tot);
  __syncthreads();      // the snapshot is complete
  if (threadIdx.x != 0) return;
  switch (nl) {
    case 2: inv_algebra<2>(a, s, ls, tot); break;
    case 3: inv_algebra<3>(a, s, ls, tot); break;
    case 4: inv_algebra<4>(a, s, ls, tot); break;
    default: inv_algebra<0>(a, s, ls, tot); break;
  }
}

static void fill_inv(qgcm_model *m, bool atmos, InvArgs &a) {
  const Grid &g = atmos ? m->ga : m->go;
  const LayerConsts &lc = atmos ? m->la : m->lo;
  a.g = g;
  a.atmos = atmos;
  a.nl = g.nl;
  a.f0 = m->fnot;
  a.beta = m->beta;
  for (int i = 0; i < NLMAX * NLMAX; ++i) { a.ctl2m[i] = lc.ctl2m[i]; a.ctm2l[i] = lc.ctm2l[i]; }
  a.q = m->F(atmos ? "qa" : "qo");
  // topset 'flat' (src/topsubs.F:99-107, :248-257) leaves ddyn identically zero: qgcm_set_field
  // notes that and the right-hand side kernel then skips one of its seven field passes
  a.ddyn = (atmos ? m->ddynat_flat : m->ddynoc_flat) ? nullptr : m->F(atmos ? "ddynat" : "ddynoc");
  a.yrel = atmos ? m->yparel : m->yporel;
  a.wrk = atmos ? m->wrk_a : m->wrk_o;
  a.hom = (!atmos && !g.cyclic) ? m->F("ochom") : nullptr;
  a.pch1 = g.cyclic ? m->F(atmos ? "pch1at" : "pch1oc") : nullptr;
  a.pch2 = g.cyclic ? m->F(atmos ? "pch2at" : "pch2oc") : nullptr;
  a.pbh = g.cyclic ? m->F(atmos ? "pbhat" : "pbhoc") : nullptr;
  a.pnew = m->F(atmos ? "pam" : "pom");
  a.coef = m->d_coef + (atmos ? 64 : 0);
}

static FusedInv fused_args(qgcm_model *m, const InvArgs &a) {
  FusedInv f;
  f.q = a.q; f.pnew = a.pnew; f.yrel = a.yrel; f.beta = a.beta; f.f0 = a.f0; f.ddyn = a.ddyn; f.hom = a.hom; f.coef = a.coef;
  for (int i = 0; i < NLMAX * NLMAX; ++i) { f.ctl2m[i] = a.ctl2m[i]; f.ctm2l[i] = a.ctm2l[i]; }
  return f;
}

static void inv_scalars_m2l(qgcm_model *m, bool atmos, const InvArgs &a, bool fused = false) {
  const Grid &g = a.g;
  HelmPlan &hp = atmos ? m->hpa : m->hpo;
  ScalArgs s;
  const LayerConsts &lc = atmos ? m->la : m->lo;
  s.atmos = atmos; s.cyclic = g.cyclic; s.nl = g.nl; s.nyp = g.nyp;
  s.dx = g.dx; s.f0 = m->fnot; s.tdt = g.tdt; s.xl = g.xl; s.yl = g.yl;
  for (int k = 0; k < NLMAX; ++k) { s.h[k] = lc.h[k]; s.gp[k] = lc.gp[k]; }
  for (int i = 0; i < NLMAX * NLMAX; ++i) { s.ctl2m[i] = lc.ctl2m[i]; s.ctm2l[i] = lc.ctm2l[i]; }
  s.rowsum = hp.rowsum;
  s.sumsrc = fused ? hp.spec : hp.rowsum;
  s.sumstride = fused ? hp.nspec : g.nyp;
  // one GPU: interior rows 1 .. nyp-2 (the wall rows are exactly zero); y-slabs: the solved rows
  s.sumlo = fused ? 0 : (m->nranks > 1 ? hp.row0 : 1);
  s.sumhi = fused ? hp.nspec : (m->nranks > 1 ? hp.row0 + hp.nrows : g.nyp - 1);
  s.sc = m->d_scal;
  s.coef = (double *)a.coef;
  s.cv = (!atmos && m->nranks > 1) ? m->d_cv : nullptr;
  s.peer.n = 0; s.peer_err = m->d_peer_err; s.cvw = m->d_cv; s.lo = hp.row0; s.hi = hp.row0 + hp.nrows;
  if (!atmos && m->nranks > 1) s.peer = peer_next_vec(m);
  QG_LAUNCH(m, "k_inv_scalars", 1, INV_NT, 0, k_inv_scalars, s);
  if (fused) {
    helm_fused_inverse(m, hp, a.wrk, g.nl, fused_args(m, a));
  } else {
    dim3 gm((g.nxp + 255) / 256, g.nyp);
    QG_LAUNCH(m, "k_m2l", gm, 256, 0, k_m2l, a);
  }
  QG_CUDA(cudaGetLastError());
  // pom <- po, po <- new: pointer rotation (src/ocisubs.F:392, src/atisubs.F:282)
  m->swapf(atmos ? "pa" : "po", atmos ? "pam" : "pom");
}

static void invert(qgcm_model *m, bool atmos) {
  InvArgs a;
  fill_inv(m, atmos, a);
  const Grid &g = a.g;
  HelmPlan &hp = atmos ? m->hpa : m->hpo;
  if (!atmos && helm_can_fuse(m, hp, g.nl)) {
    // box ocean, fast DST plan: no k_l2m / k_m2l, the projections ride on the solver's kernels
    const FusedInv fz = fused_args(m, a);
    helm_solve_a(m, hp, a.wrk, g.nl, &fz);
    helm_solve_b(m, hp, a.wrk, g.nl, &fz);
    inv_scalars_m2l(m, atmos, a, true);
    return;
  }
  dim3 gi((g.nxp + 511) / 512, g.nyp - 2);
  QG_LAUNCH(m, "k_l2m", gi, 256, 0, k_l2m, a);
  helm_solve(m, hp, a.wrk, g.nl);
  inv_scalars_m2l(m, atmos, a);
}

// y-slab pieces of ocinvq (the drivers in slab.cu put an all-gather after a and an all-reduce
// after b)
void ocinvq_phase_a(qgcm_model *m) {
  InvArgs a;
  fill_inv(m, false, a);
  const Grid &g = a.g;
  if (helm_can_fuse(m, m->hpo, g.nl)) {
    const FusedInv fz = fused_args(m, a);
    helm_solve_a(m, m->hpo, a.wrk, g.nl, &fz);
    return;
  }
  dim3 gi((g.nxp + 511) / 512, g.nyp - 2);
  QG_LAUNCH(m, "k_l2m", gi, 256, 0, k_l2m, a);
  helm_solve_a(m, m->hpo, a.wrk, g.nl);
}
void ocinvq_phase_b(qgcm_model *m) {
  const Grid &g = m->go;
  HelmPlan &hp = m->hpo;
  const bool fused = helm_can_fuse(m, hp, g.nl);
  if (fused) {
    InvArgs a;
    fill_inv(m, false, a);
    const FusedInv fz = fused_args(m, a);
    helm_solve_b(m, hp, m->wrk_o, g.nl, &fz);
  } else {
    helm_solve_b(m, hp, m->wrk_o, g.nl);
  }
  if (peer_active(m)) return;      // k_inv_scalars forms and all-reduces the integrals itself
  PeerCtx none = {};
  if (fused)
    QG_LAUNCH(m, "k_inv_partials", 1, 256, 0, k_inv_partials, hp.spec, g.nl, hp.nspec, 0, hp.nspec, g.dx, m->d_cv, none, m->d_peer_err);
  else
    QG_LAUNCH(m, "k_inv_partials", 1, 256, 0, k_inv_partials, hp.rowsum, g.nl, g.nyp, hp.row0, hp.row0 + hp.nrows, g.dx, m->d_cv,
              none, m->d_peer_err);
}
void ocinvq_phase_c(qgcm_model *m) {
  InvArgs a;
  fill_inv(m, false, a);
  inv_scalars_m2l(m, false, a, helm_can_fuse(m, m->hpo, m->go.nl));
}

void launch_ocinvq(qgcm_model *m) { invert(m, false); }
void launch_atinvq(qgcm_model *m) { invert(m, true); }

// ------------------------------------------------------------------------------------
// homsol (src/conhoms.F:318-818) through the device solver
// ------------------------------------------------------------------------------------
__global__ void k_fill_rows(double *w, int ld, int nyp, int nxp, const double *rowval, double cst, int use_row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= nxp) return;
  w[(size_t)j * ld + i] = use_row ? rowval[j] : cst;
}
// out = base + rdm2 * sol, and record xintp row sums of `out`
__global__ void __launch_bounds__(256) k_hom_finish(double *out, const double *sol, int ld, int nyp, int nxp,
                                                    const double *rowval, double cst, int use_row, double rdm2,
                                                    double *rowsum) {
  __shared__ double red[32];
  const int j = blockIdx.x;
  double part = 0.0;
  for (int i = threadIdx.x; i < nxp; i += blockDim.x) {
    const double v = (use_row ? rowval[j] : cst) + rdm2 * sol[(size_t)j * ld + i];
    out[(size_t)j * ld + i] = v;
    part += (i == 0 || i == nxp - 1) ? 0.5 * v : v;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  if (lane == 0) red[w] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    rowsum[j] = t;
  }
}

static double xintp_from_rowsums(const std::vector<double> &rs) {
  const int nyp = (int)rs.size();
  double sump = 0.0;
  for (int j = 1; j < nyp - 1; ++j) sump += rs[j];
  return sump + 0.5 * (rs[0] + rs[nyp - 1]);
}
// the same integral restricted to the rows a y-slab owns (its share of the global sum)
static double xintp_share(const Grid &g, const std::vector<double> &rs) {
  const int lo = g.own0 + (g.wall_s() ? 1 : 0), hi = g.own1 - (g.wall_n() ? 1 : 0);
  double sump = 0.0;
  for (int j = lo; j < hi; ++j) sump += rs[j];
  if (g.wall_s()) sump += 0.5 * rs[0];
  if (g.wall_n()) sump += 0.5 * rs[g.nyp - 1];
  return sump;
}

static void homsol_channel(qgcm_model *m, bool atmos) {
  const Grid &g = atmos ? m->ga : m->go;
  const LayerConsts &lc = atmos ? m->la : m->lo;
  HelmPlan &hp = atmos ? m->hpa : m->hpo;
  const std::vector<double> &yp = atmos ? m->h_ypa : m->h_ypo;
  const int nyp = g.nyp, nl = g.nl;
  qgcm_scalars s;
  QG_CUDA(cudaMemcpyAsync(&s, m->d_scal, sizeof(s), cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
  std::vector<double> pbh(nyp), l1(nyp), l2(nyp), rs(nyp), col(nyp);
  for (int j = 1; j <= nyp; ++j) pbh[j - 1] = (double)(nyp - j) / (double)(nyp - 1);
  double *d_pbh = m->F(atmos ? "pbhat" : "pbhoc");
  QG_CUDA(cudaMemcpy(d_pbh, pbh.data(), sizeof(double) * nyp, cudaMemcpyHostToDevice));
  (atmos ? s.hbsiat : s.hbsioc) = g.yl / g.xl;
  (atmos ? s.aipbha : s.aipbho) = 0.5 * g.xl * g.yl;
  double *wrk = atmos ? m->wrk_a : m->wrk_o;   // modes 0,1 of the work array as the two RHS
  double *d_row = m->d_red;                     // 2*nyp + nyp doubles of scratch
  double *d_p1 = m->F(atmos ? "pch1at" : "pch1oc"), *d_p2 = m->F(atmos ? "pch2at" : "pch2oc");
  std::vector<double> p1(nyp), p2(nyp);
  for (int mo = 1; mo <= nl - 1; ++mo) {
    const double rdm2 = lc.rdm2[mo];
    for (int j = 0; j < nyp; ++j) {
      l1[j] = (yp[nyp - 1] - yp[j]) / g.yl;
      l2[j] = (yp[j] - yp[0]) / g.yl;
    }
    // both right-hand sides solved as a batch of two "modes" with mode mo's operator:
    // build a temporary diagonal table with the same coefficients in slots 0 and 1
    std::vector<double> b((size_t)nl * hp.n);
    {
      const double PI2 = 6.28318530717958648;
      std::vector<double> bd2(hp.n);
      const double a = hp.a;
      for (int i = 2; i <= hp.n / 2; ++i) {
        int i1 = 2 * i - 1;
        bd2[i1 - 2] = -2.0 * a + 2.0 * g.dxm2 * (cos((i - 1) * PI2 / hp.n) - 1.0);
        bd2[i1 - 1] = bd2[i1 - 2];
      }
      bd2[0] = -2.0 * a;
      bd2[hp.n - 1] = -2.0 * a - 4.0 * g.dxm2;
      for (int q = 0; q < nl; ++q)
        for (int i = 0; i < hp.n; ++i) b[(size_t)q * hp.n + i] = bd2[i] - rdm2;
    }
    helm_set_diag(m, hp, b.data());
    QG_CUDA(cudaMemcpy(d_row, l1.data(), sizeof(double) * nyp, cudaMemcpyHostToDevice));
    QG_CUDA(cudaMemcpy(d_row + nyp, l2.data(), sizeof(double) * nyp, cudaMemcpyHostToDevice));
    dim3 gf((g.nxp + 255) / 256, nyp);
    QG_LAUNCH(m, "k_fill_rows", gf, 256, 0, k_fill_rows, wrk, g.ld, nyp, g.nxp, d_row, 0.0, 1);
    QG_LAUNCH(m, "k_fill_rows", gf, 256, 0, k_fill_rows, wrk + g.lsz, g.ld, nyp, g.nxp, d_row + nyp, 0.0, 1);
    hp.walls_dirty = true;
    helm_solve(m, hp, wrk, 2);
    double aip[2];
    for (int q = 0; q < 2; ++q) {
      QG_LAUNCH(m, "k_hom_finish", nyp, 256, 0, k_hom_finish, wrk + q * g.lsz, wrk + q * g.lsz, g.ld, nyp, g.nxp, d_row + q * nyp, 0.0, 1,
                                               rdm2, d_row + 2 * nyp);
      QG_CUDA(cudaMemcpyAsync(rs.data(), d_row + 2 * nyp, sizeof(double) * nyp, cudaMemcpyDeviceToHost, m->stream));
      // column 1 of the solution is the 1-D profile (src/conhoms.F:478-479)
      QG_CUDA(cudaMemcpy2DAsync(col.data(), sizeof(double), wrk + q * g.lsz, sizeof(double) * g.ld, sizeof(double), nyp,
                                cudaMemcpyDeviceToHost, m->stream));
      QG_CUDA(cudaStreamSynchronize(m->stream));
      aip[q] = xintp_from_rowsums(rs);
      (q == 0 ? p1 : p2) = col;
    }
    QG_CUDA(cudaMemcpy(d_p1 + (size_t)(mo - 1) * nyp, p1.data(), sizeof(double) * nyp, cudaMemcpyHostToDevice));
    QG_CUDA(cudaMemcpy(d_p2 + (size_t)(mo - 1) * nyp, p2.data(), sizeof(double) * nyp, cudaMemcpyHostToDevice));
    const double dx = g.dx, dy = g.dx, xl = g.xl;
    (atmos ? s.aipcha : s.aipcho)[mo - 1] = 0.5 * (aip[0] + aip[1]) * dx * dy;
    double pch1ys = (p1[1] - p1[0]) / dy, pch2ys = (p2[1] - p2[0]) / dy;
    double pch1yn = (p1[nyp - 1] - p1[nyp - 2]) / dy, pch2yn = (p2[nyp - 1] - p2[nyp - 2]) / dy;
    pch1ys = -pch1ys + 0.5 * dy * rdm2 * p1[0];
    pch2ys = -pch2ys + 0.5 * dy * rdm2 * p2[0];
    pch1yn = pch1yn + 0.5 * dy * rdm2 * p1[nyp - 1];
    pch2yn = pch2yn + 0.5 * dy * rdm2 * p2[nyp - 1];
    pch1ys *= xl; pch2ys *= xl; pch1yn *= xl; pch2yn *= xl;
    const double det = pch1ys * pch2yn - pch2ys * pch1yn;
    (atmos ? s.hc1sat : s.hc1soc)[mo - 1] = pch1ys / det;
    (atmos ? s.hc2sat : s.hc2soc)[mo - 1] = pch2ys / det;
    (atmos ? s.hc1nat : s.hc1noc)[mo - 1] = pch1yn / det;
    (atmos ? s.hc2nat : s.hc2noc)[mo - 1] = pch2yn / det;
  }
  // restore the per-mode operators for the time loop
  {
    std::vector<double> b((size_t)nl * hp.n), bd2(hp.n);
    const double PI2 = 6.28318530717958648, a = hp.a;
    for (int i = 2; i <= hp.n / 2; ++i) {
      int i1 = 2 * i - 1;
      bd2[i1 - 2] = -2.0 * a + 2.0 * g.dxm2 * (cos((i - 1) * PI2 / hp.n) - 1.0);
      bd2[i1 - 1] = bd2[i1 - 2];
    }
    bd2[0] = -2.0 * a;
    bd2[hp.n - 1] = -2.0 * a - 4.0 * g.dxm2;
    for (int q = 0; q < nl; ++q)
      for (int i = 0; i < hp.n; ++i) b[(size_t)q * hp.n + i] = bd2[i] - lc.rdm2[q];
    helm_set_diag(m, hp, b.data());
  }
  helm_clean_walls(m, hp, wrk, nl);     // k_hom_finish wrote the channel profiles over the wall rows
  QG_CUDA(cudaMemcpy(m->d_scal, &s, sizeof(s), cudaMemcpyHostToDevice));
}

// homsol for the box ocean (src/conhoms.F:549-640) in three pieces so that the y-slab driver
// can put the solver's all-gather after a and the all-reduce of the integrals after b.
// ochom(:,:,m) = 1 + rdm2(m+1) * sol0, sol0 solving (del2 - rdm2(m+1)) sol0 = 1.  The per-mode
// operator table already holds mode m+1 in slot m, so all nl slots are solved with rhs = 1
// and slots 1..nl-1 are kept.
void homsol_box_a(qgcm_model *m) {
  const Grid &g = m->go;
  dim3 gf((g.nxp + 255) / 256, g.nyp);
  for (int q = 0; q < g.nl; ++q)
    QG_LAUNCH(m, "k_fill_rows", gf, 256, 0, k_fill_rows, m->wrk_o + q * g.lsz, g.ld, g.nyp, g.nxp, nullptr, 1.0, 0);
  m->hpo.walls_dirty = true;
  helm_solve_a(m, m->hpo, m->wrk_o, g.nl);
}
void homsol_box_b(qgcm_model *m, std::vector<double> &share) {
  const Grid &g = m->go;
  const LayerConsts &lc = m->lo;
  const int nl = g.nl, nyp = g.nyp;
  helm_solve_b(m, m->hpo, m->wrk_o, nl);
  double *ochom = m->F("ochom");
  std::vector<double> rs(nyp);
  share.assign(nl - 1, 0.0);
  for (int mo = 1; mo <= nl - 1; ++mo) {
    QG_LAUNCH(m, "k_hom_finish", nyp, 256, 0, k_hom_finish, ochom + (size_t)(mo - 1) * g.lsz, m->wrk_o + (size_t)mo * g.lsz, g.ld, nyp,
                                             g.nxp, nullptr, 1.0, 0, lc.rdm2[mo], m->d_red);
    QG_CUDA(cudaMemcpyAsync(rs.data(), m->d_red, sizeof(double) * nyp, cudaMemcpyDeviceToHost, m->stream));
    QG_CUDA(cudaStreamSynchronize(m->stream));
    share[mo - 1] = xintp_share(g, rs) * g.dx * g.dx;
  }
}
void homsol_box_c(qgcm_model *m, const std::vector<double> &aipohs) {
  const Grid &g = m->go;
  const LayerConsts &lc = m->lo;
  const int nl = g.nl;
  qgcm_scalars s;
  QG_CUDA(cudaMemcpyAsync(&s, m->d_scal, sizeof(s), cudaMemcpyDeviceToHost, m->stream));
  QG_CUDA(cudaStreamSynchronize(m->stream));
  for (int mo = 1; mo <= nl - 1; ++mo) s.aipohs[mo - 1] = aipohs[mo - 1];
  for (int k = 1; k <= nl - 1; ++k) {
    for (int mo = 1; mo <= nl; ++mo)
      s.cdiffo[(mo - 1) + nl * (k - 1)] = lc.ctm2l[(mo - 1) + nl * k] - lc.ctm2l[(mo - 1) + nl * (k - 1)];
    for (int mo = 1; mo <= nl - 1; ++mo)
      s.cdhoc[(k - 1) + (nl - 1) * (mo - 1)] = (lc.ctm2l[mo + nl * k] - lc.ctm2l[mo + nl * (k - 1)]) * s.aipohs[mo - 1];
  }
  helm_clean_walls(m, m->hpo, m->wrk_o, nl);
  QG_CUDA(cudaMemcpy(m->d_scal, &s, sizeof(s), cudaMemcpyHostToDevice));
}

static void homsol_box(qgcm_model *m) {
  std::vector<double> aip;
  homsol_box_a(m);
  homsol_box_b(m, aip);
  homsol_box_c(m, aip);
}

void launch_homsol(qgcm_model *m) {
  if (m->nranks > 1) { slab_homsol(ranks_of(m)); return; }
  if (m->has_ocean) {
    if (m->cyclic) homsol_channel(m, false); else homsol_box(m);
  }
  if (m->has_atmos) homsol_channel(m, true);
}

}  // namespace qg

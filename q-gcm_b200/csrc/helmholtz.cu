// Batched modified-Helmholtz solver: replaces hsbxoc / hscyoc / hscyat
// (src/ocisubs.F:415-618, src/atisubs.F:301-395) and the FFTPACK routines they call
// (src/fftpack/newbihar/dsint.f, drfftf.f, drfftb.f).
//
//   x-direction : one thread block per (mode, row); the row lives in shared memory, a
//                 Stockham mixed-radix complex FFT of length n/2 plus the real / sine
//                 pre- and post-processing (published FFTPACK algorithm, dsint.f:17-43).
//   y-direction : the constant-coefficient tridiagonal systems (one per wavenumber)
//                 are partitioned into chunks of TRI_L rows.  Each chunk is solved
//                 locally in registers (Thomas), the chunk interfaces are coupled by a
//                 small block-tridiagonal system per wavenumber, and the correction
//                 (two precomputed spike vectors, Toeplitz => identical for every
//                 chunk) is applied on the fly when the inverse transform loads a row.
//                 All global accesses are full-row coalesced; no transpose.
#include <algorithm>
#include <cmath>

#include "qgcm_internal.h"

namespace qg {

// --------------------------------------------------------------------------------------
// complex helpers
// --------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmulmi(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// Internal twiddles W_R^{n2*k1} of the two-level (Cooley-Tukey) in-register butterflies,
// filled by helm_plan_create.  Uniform index across a warp => constant-cache broadcast.
__constant__ double2 c_ctw[160];
__host__ __device__ constexpr int ctw_off(int R) {
  return R == 6 ? 0 : R == 9 ? 6 : R == 10 ? 15 : R == 12 ? 25 : R == 15 ? 37 : R == 16 ? 52 : R == 20 ? 68 : R == 25 ? 88 : 113;
}

template <int R>
__device__ __forceinline__ void dft(double2 *v);

template <>
__device__ __forceinline__ void dft<2>(double2 *v) {
  double2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft<4>(double2 *v) {
  double2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = cmulmi(csub(v[1], v[3]));
  v[0] = cadd(a, c);
  v[1] = cadd(b, d);
  v[2] = csub(a, c);
  v[3] = csub(b, d);
}
template <>
__device__ __forceinline__ void dft<3>(double2 *v) {
  const double s = 0.86602540378443864676;
  double2 t1 = cadd(v[1], v[2]);
  double2 t2 = make_double2(v[0].x - 0.5 * t1.x, v[0].y - 0.5 * t1.y);
  double2 t3 = cscale(cmulmi(csub(v[1], v[2])), s);
  v[0] = cadd(v[0], t1);
  v[1] = cadd(t2, t3);
  v[2] = csub(t2, t3);
}
template <>
__device__ __forceinline__ void dft<5>(double2 *v) {
  const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
  const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
  double2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
  double2 b1 = cmulmi(csub(v[1], v[4])), b2 = cmulmi(csub(v[2], v[3]));
  double2 r1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
  double2 r2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
  double2 i1 = make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
  double2 i2 = make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
  v[0] = cadd(v[0], cadd(a1, a2));
  v[1] = cadd(r1, i1);
  v[4] = csub(r1, i1);
  v[2] = cadd(r2, i2);
  v[3] = csub(r2, i2);
}
template <>
__device__ __forceinline__ void dft<8>(double2 *v) {
  const double h = 0.70710678118654752440;
  double2 e[4] = {v[0], v[2], v[4], v[6]};
  double2 o[4] = {v[1], v[3], v[5], v[7]};
  dft<4>(e);
  dft<4>(o);
  double2 w1 = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));     // o1 * e^{-i pi/4}
  double2 w2 = cmulmi(o[2]);                                                   // o2 * (-i)
  double2 w3 = make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));    // o3 * e^{-3i pi/4}
  v[0] = cadd(e[0], o[0]);
  v[4] = csub(e[0], o[0]);
  v[1] = cadd(e[1], w1);
  v[5] = csub(e[1], w1);
  v[2] = cadd(e[2], w2);
  v[6] = csub(e[2], w2);
  v[3] = cadd(e[3], w3);
  v[7] = csub(e[3], w3);
}

// R = RA*RB point DFT in registers: RB sub-DFTs of length RA over n1 (n = n1*RB + n2),
// constant twiddles W_R^{n2 k1}, then RA sub-DFTs of length RB; output k = k1 + RA*k2.
template <int RA, int RB>
__device__ __forceinline__ void dft_ct(double2 *v) {
  constexpr int R = RA * RB;
  double2 y[R];
#pragma unroll
  for (int n2 = 0; n2 < RB; ++n2) {
    double2 t[RA];
#pragma unroll
    for (int n1 = 0; n1 < RA; ++n1) t[n1] = v[n1 * RB + n2];
    dft<RA>(t);
#pragma unroll
    for (int k1 = 0; k1 < RA; ++k1)
      y[k1 * RB + n2] = (n2 == 0 || k1 == 0) ? t[k1] : cmul(t[k1], c_ctw[ctw_off(R) + k1 * RB + n2]);
  }
#pragma unroll
  for (int k1 = 0; k1 < RA; ++k1) {
    double2 t[RB];
#pragma unroll
    for (int n2 = 0; n2 < RB; ++n2) t[n2] = y[k1 * RB + n2];
    dft<RB>(t);
#pragma unroll
    for (int k2 = 0; k2 < RB; ++k2) v[k1 + RA * k2] = t[k2];
  }
}
template <> __device__ __forceinline__ void dft<6>(double2 *v) { dft_ct<2, 3>(v); }
template <> __device__ __forceinline__ void dft<9>(double2 *v) { dft_ct<3, 3>(v); }
template <> __device__ __forceinline__ void dft<10>(double2 *v) { dft_ct<2, 5>(v); }
template <> __device__ __forceinline__ void dft<12>(double2 *v) { dft_ct<4, 3>(v); }
template <> __device__ __forceinline__ void dft<15>(double2 *v) { dft_ct<3, 5>(v); }
template <> __device__ __forceinline__ void dft<16>(double2 *v) { dft_ct<4, 4>(v); }

// One Stockham pass of radix R over the complex array `in` (length M) into `out`.
// Ns = product of the radices of the previous passes; tw = this pass's twiddle table,
// laid out [r-1][k] so that consecutive lanes read consecutive entries.
template <int R>
__device__ __forceinline__ void fft_pass(const double2 *__restrict__ in, double2 *__restrict__ out, int M, int Ns,
                                         const double2 *__restrict__ tw) {
  const int L = M / R;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const int k = (Ns == 1) ? 0 : j % Ns;
    const int j0 = (j - k) * R + k;
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[j + r * L];
    if (Ns > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], __ldg(&tw[(r - 1) * Ns + k]));
    }
    dft<R>(v);
#pragma unroll
    for (int r = 0; r < R; ++r) out[j0 + r * Ns] = v[r];
  }
}

struct FftDev {
  int n, m, nrad;
  int radix[8];
  int twoff[8];           // offset of each pass's twiddle table in tw
  const double2 *tw, *wn;
  const double *sintw;
};

// forward complex FFT of length p.m; data in a, scratch b; returns pointer holding the result
__device__ __forceinline__ double2 *cfft_smem(const FftDev &p, double2 *a, double2 *b) {
  int Ns = 1;
  for (int s = 0; s < p.nrad; ++s) {
    const int R = p.radix[s];
    const double2 *tw = p.tw + p.twoff[s];
    switch (R) {
      case 2: fft_pass<2>(a, b, p.m, Ns, tw); break;
      case 3: fft_pass<3>(a, b, p.m, Ns, tw); break;
      case 4: fft_pass<4>(a, b, p.m, Ns, tw); break;
      case 5: fft_pass<5>(a, b, p.m, Ns, tw); break;
      case 6: fft_pass<6>(a, b, p.m, Ns, tw); break;
      case 8: fft_pass<8>(a, b, p.m, Ns, tw); break;
      case 9: fft_pass<9>(a, b, p.m, Ns, tw); break;
      case 10: fft_pass<10>(a, b, p.m, Ns, tw); break;
      case 12: fft_pass<12>(a, b, p.m, Ns, tw); break;
      case 15: fft_pass<15>(a, b, p.m, Ns, tw); break;
      default: fft_pass<16>(a, b, p.m, Ns, tw); break;
    }
    __syncthreads();
    double2 *t = a;
    a = b;
    b = t;
    Ns *= R;
  }
  return a;
}

// block-wide sum, result valid in every thread; red must hold 32 doubles
__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];   // fixed order: deterministic
  return t;
}

// Parameters of one x-transform launch
struct XfArgs {
  FftDev f;
  int kind;        // 0 DST-I, 1 real FFT
  int inverse;     // 0 forward (rhs -> spectrum), 1 inverse (tridiagonal solution -> field)
  int ld, nyp, nxp, row0;
  size_t lsz;      // mode stride
  double *wrk;
  double *rowsum;                // [nmodes][nyp] (inverse only)
};

// F_k of the length-n real transform from the half-length complex transform Z (n = 2m):
// F_k = (Z_k + conj Z_{m-k})/2 - (i/2) e^{-2 pi i k/n} (Z_k - conj Z_{m-k})
__device__ __forceinline__ double2 real_post(double2 za, double2 zmk, double2 w) {
  const double2 zb = cconj(zmk);
  const double2 e = cscale(cadd(za, zb), 0.5);
  const double2 o = cmul(cscale(cmulmi(csub(za, zb)), 0.5), w);
  return cadd(e, o);
}

// grid (nrows, nmodes); dynamic smem: two complex buffers of length m, then 40 doubles
__global__ void __launch_bounds__(256, 2) k_xform(XfArgs a) {
  extern __shared__ double2 smem2[];
  const int N = a.f.n, M = a.f.m;
  double2 *A = smem2, *B = smem2 + M;
  double *red = reinterpret_cast<double *>(smem2 + 2 * M);
  const int r = blockIdx.x;            // solved row index, local row r + row0
  const int mode = blockIdx.y;
  double *__restrict__ row = a.wrk + (size_t)mode * a.lsz + (size_t)(r + a.row0) * a.ld;
  double *Ar = reinterpret_cast<double *>(A);
  const int T = blockDim.x, t = threadIdx.x;
  const int lane = t & 31, w = t >> 5, nw = T >> 5;

  if (a.kind == 0) {
    // ---------------- DST-I of row(2:nxto), dsint.f:17-43 ----------------
    // load + pre-processing fused: t_k = (x_k - x_{N-k}) + 2 sin(k pi/N)(x_k + x_{N-k})
    for (int k = t; k <= M; k += T) {
      if (k == 0) {
        Ar[0] = 0.0;
      } else if (k == M) {
        Ar[M] = 4.0 * row[M];
      } else {
        const double xa = row[k], xb = row[N - k];
        const double t1 = xa - xb, t2 = __ldg(&a.f.sintw[k]) * (xa + xb);
        Ar[k] = t1 + t2;
        Ar[N - k] = t2 - t1;
      }
    }
    __syncthreads();
    double2 *Z = cfft_smem(a.f, A, B);
    double2 *O = (Z == A) ? B : A;
    double *Or = reinterpret_cast<double *>(O);
    // real post-processing for the pair (k, M-k); even outputs out_{2k} = -Im F_k go to
    // Or[2k], Re F_k (the summand of the odd outputs) is parked in Or[2k+1]
    for (int k = t; k <= M / 2; k += T) {
      if (k == 0) {
        Or[1] = 0.5 * (Z[0].x + Z[0].y);          // out_1 = F_0 / 2
      } else {
        const double2 za = Z[k], zb = Z[M - k];
        const double2 Fa = real_post(za, zb, __ldg(&a.f.wn[k]));
        Or[2 * k] = -Fa.y;
        Or[2 * k + 1] = Fa.x;
        if (k != M - k) {
          const double2 Fb = real_post(zb, za, __ldg(&a.f.wn[M - k]));
          Or[2 * (M - k)] = -Fb.y;
          Or[2 * (M - k) + 1] = Fb.x;
        }
      }
    }
    __syncthreads();
    // odd outputs are a running sum (dsint.f:33-37): out_{2k+1} = sum_{l<=k} Or[2l+1].
    // Each warp scans a contiguous segment 32 elements at a time, then adds the offset
    // of the preceding segments.
    const int seg = (M + nw - 1) / nw;
    const int s0 = w * seg, s1 = min(M, s0 + seg);
    double carry = 0.0;
    for (int base = s0; base < s1; base += 32) {
      const int k = base + lane;
      double v = (k < s1) ? Or[2 * k + 1] : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double nb = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += nb;
      }
      v += carry;
      if (k < s1) Or[2 * k + 1] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) red[w] = carry;
    __syncthreads();
    double off = 0.0;
    for (int i = 0; i < w; ++i) off += red[i];
    __syncthreads();
    for (int base = s0; base < s1; base += 32) {
      const int k = base + lane;
      if (k < s1) Or[2 * k + 1] += off;
    }
    __syncthreads();
    double part = 0.0;
    for (int k = t + 1; k < N; k += T) {
      const double v = Or[k];
      row[k] = v;
      part += v;
    }
    if (a.inverse) {
      if (t == 0) {
        row[0] = 0.0;
        row[a.nxp - 1] = 0.0;
      }
      const double s = block_sum(part, red);
      if (t == 0) a.rowsum[(size_t)mode * a.nyp + (r + a.row0)] = s;
    }
  } else if (!a.inverse) {
    // ---------------- forward real FFT, packed order (fft.doc:96-114) ----------------
    for (int k = t; k < N; k += T) Ar[k] = row[k];
    __syncthreads();
    double2 *Z = cfft_smem(a.f, A, B);
    double2 *O = (Z == A) ? B : A;
    double *Or = reinterpret_cast<double *>(O);
    for (int k = t; k <= M / 2; k += T) {
      if (k == 0) {
        Or[0] = Z[0].x + Z[0].y;
        Or[N - 1] = Z[0].x - Z[0].y;
      } else {
        const double2 za = Z[k], zb = Z[M - k];
        const double2 Fa = real_post(za, zb, __ldg(&a.f.wn[k]));
        Or[2 * k - 1] = Fa.x;
        Or[2 * k] = Fa.y;
        if (k != M - k) {
          const double2 Fb = real_post(zb, za, __ldg(&a.f.wn[M - k]));
          Or[2 * (M - k) - 1] = Fb.x;
          Or[2 * (M - k)] = Fb.y;
        }
      }
    }
    __syncthreads();
    for (int k = t; k < N; k += T) row[k] = Or[k];
  } else {
    // ---------------- inverse real FFT (drfftb), then periodic wrap ----------------
    double *Br = reinterpret_cast<double *>(B);
    for (int k = t; k < N; k += T) Br[k] = row[k];
    __syncthreads();
    for (int k = t; k < M; k += T) {
      double2 xa, xb;   // X_k, conj X_{M-k}
      if (k == 0) {
        xa = make_double2(Br[0], 0.0);
        xb = make_double2(Br[N - 1], 0.0);
      } else {
        xa = make_double2(Br[2 * k - 1], Br[2 * k]);
        xb = make_double2(Br[2 * (M - k) - 1], -Br[2 * (M - k)]);
      }
      const double2 e = cadd(xa, xb);
      const double2 d = csub(xa, xb);
      const double2 dw = cmul(d, cconj(__ldg(&a.f.wn[k])));
      const double2 o = make_double2(-dw.y, dw.x);   // i * dw
      A[k] = cconj(cadd(e, o));
    }
    __syncthreads();
    double2 *Z = cfft_smem(a.f, A, B);
    const double *Zr = reinterpret_cast<const double *>(Z);
    // x_{2n} = Re Z_n, x_{2n+1} = -Im Z_n: the buffer is the output up to the sign of odd entries
    double part = 0.0;
    for (int k = t; k < N; k += T) {
      const double v = (k & 1) ? -Zr[k] : Zr[k];
      row[k] = v;
      if (k > 0) part += v;
    }
    if (t == 0) row[a.nxp - 1] = Zr[0];
    const double s = block_sum(part, red);
    // xintp row sum: 0.5*v(1) + sum_{2}^{nxp-1} + 0.5*v(nxp), v(nxp) = v(1)
    if (t == 0) a.rowsum[(size_t)mode * a.nyp + (r + a.row0)] = 0.5 * Zr[0] + s + 0.5 * Zr[0];
  }
}

// --------------------------------------------------------------------------------------
// Fast path for the box solver: DST-I rows whose half length is M = 16*15*R3 = 240*R3
// (all box benchmark decks: 2400 = 240*10, 1200 = 240*5, 480 = 240*2).
//
//   * persistent blocks (two of 256 threads per SM; four of 128 threads for M <= 1200, see dst3_threads),
//     each walks rows blockIdx.x, +gridDim.x, ...
//   * the raw row is fetched by the TMA engine (cp.async.bulk global->shared, mbarrier
//     completion) while the previous row is post-processed and stored, so no warp ever
//     waits on HBM;
//   * the FFTPACK pre-processing (dsint.f:17-30) is fused into the register load of the
//     first butterfly pass, its sine weights are rebuilt from a per-thread base angle and
//     R1 constants (no table traffic);
//   * three register butterfly passes (16, 15, R3: one butterfly per thread in every pass -- the last one
//     in two rounds on 128-thread blocks -- under 128 registers) with two shared-memory exchanges laid out
//     free of bank conflicts (the first one skewed by i >> 4); the exchanges ping-pong
//     between the exchange buffer and the raw row's buffer (free once pass 1 has read it), so
//     no pass needs a barrier between its loads and its stores: six block barriers per row;
//     the last pass builds its twiddles as powers of a per-thread base (two loads);
//   * real post-processing on the (k, M-k) pair, the running sum of dsint.f:33-37 as a
//     one-sweep block scan over contiguous segments, and the interleaved result goes
//     straight from registers to HBM with coalesced 16-byte stores.
// All strides are compile-time constants.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// read-only load the compiler may not hoist out of the row loop (keeps the per-thread bases
// out of the register budget of the butterflies)
__device__ __forceinline__ double2 ldg2_nohoist(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// tw[r] = w^r, r = 1..R-1, from w and w^2: tw[2q] = tw[q]^2, tw[2q+1] = tw[q] tw[q+1]
// (depth <= 3 products for R <= 16)
template <int R>
__device__ __forceinline__ void twiddle_apply(double2 *v, double2 w1, double2 w2) {
  double2 tw[R > 2 ? R : 3];
  tw[1] = w1;
  tw[2] = w2;
#pragma unroll
  for (int r = 3; r < R; ++r) tw[r] = (r & 1) ? cmul(tw[r / 2], tw[r / 2 + 1]) : cmul(tw[r / 2], tw[r / 2]);
#pragma unroll
  for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw[r]);
}

__host__ __device__ constexpr int dst3_seg(int M, int NT) {
  // values per thread in the running-sum sweep: even, >= M/NT, divides M, preferably an
  // odd number of 16-byte units (conflict-free 128-bit shared loads)
  int best = 0;
  for (int s = 2; s <= 64; s += 2)
    if (M % s == 0 && s * NT >= M) {
      if (!best) best = s;
      if ((s / 2) & 1) return s;
    }
  return best;
}

// k_dst3 variants.  The DST is linear, so the layer<->mode projections of ocinvq commute with
// it: the fused forward transform works on the vorticity LAYERS (the projection is applied by
// the tridiagonal kernels, which hold all modes of a wavenumber anyway), and the fused inverse
// transform produces pressure LAYERS.  That removes the two pointwise kernels (k_l2m, k_m2l)
// and the 12 field passes they existed for.
// DST_FUSED_FT: DST_FUSED_F over topography (its own instantiation: the flat-bottom kernel carries no predicated
// ddynoc loads, which were 7 % of its instructions)
enum { DST_PLAIN_F = 0, DST_PLAIN_I = 1, DST_FUSED_F = 2, DST_FUSED_I = 3, DST_FUSED_FT = 4 };
struct Dst3Args {
  int nitems, nrows;        // (mode,row) work items; solved rows per mode
  int ld, nyp, nxp, row0;
  size_t lsz;
  double *wrk;
  double *rowsum;
  // fused variants (items are row-major: the nl layers of a row are adjacent in time, so the
  // rows every layer re-reads -- ochom -- come from L2)
  int nl, kbot;
  const double *src;        // DST_FUSED_F: q [nl][nyp][ld] (src/ocisubs.F:117-139 without the projection)
  double *dst;              // DST_FUSED_I: new p [nl][nyp][ld] (src/ocisubs.F:377-401)
  const double *yrel;       // DST_FUSED_F: beta*y is subtracted from the row
  double beta;
  const double *ddyn;       // DST_FUSED_F: topography term of the bottom layer, null over a flat bottom
  const double *hom;        // DST_FUSED_I: ochom [nl-1][nyp][ld]
  const double *coef;       // DST_FUSED_I: device hclco[nl-1]
  int wall_s, wall_n;       // DST_FUSED_I: this grid holds the southern / northern wall row (written by block 0)
  int pdl;                  // DST_FUSED_I: launched as a programmatic dependent of k_inv_scalars
  double ctm2l[NLMAX * NLMAX];
  const double2 *s1base;    // [L1][2]  (2 sin, 2 cos) of pi*(2t)/N and pi*(2t+1)/N
  const double2 *tw2;       // [R2-1][R1] twiddles of pass 2
  const double2 *tw3base;   // [L3][2]  w, w^2 of pass 3
  const double2 *wnbase;    // [L3]     exp(-2 pi i t/N)
  double c1[16], s1[16];    // cos, sin of pi q/R1
  double2 wnr[16];          // exp(-2 pi i q L3/N)
};

// Threads per block: the first two passes have M/16 and M/15 butterflies, so rows of half length M <= 1200
// (NAtl 2 km and coarser) keep less than a third of a 256-thread block busy in them, and their per-row
// latencies (six barriers, the exchanges) weigh twice as much against half the arithmetic.  Those plans run
// 128-thread blocks, four per SM instead of two -- twice the rows in flight per SM from the same registers
// and less shared memory -- and take the 240 butterflies of the last pass in two rounds.
#ifndef DST3_SMALL_NT
#define DST3_SMALL_NT 128      // 256: every plan on 256-thread blocks (A/B builds)
#endif
__host__ __device__ constexpr int dst3_threads(int R3) { return R3 <= 5 ? DST3_SMALL_NT : 256; }
__host__ __device__ constexpr int dst3_blocks(int R3) { return 512 / dst3_threads(R3); }

template <int R3, int MODE>
__global__ void __launch_bounds__(dst3_threads(R3), dst3_blocks(R3)) k_dst3(const Dst3Args a) {
  constexpr int NT = dst3_threads(R3), NW = NT / 32;
  constexpr bool INV = (MODE == DST_PLAIN_I);       // xintp row sums of the result, wall column zeroed
  constexpr bool FWD_F = (MODE == DST_FUSED_F || MODE == DST_FUSED_FT), TOPO = (MODE == DST_FUSED_FT);
  constexpr bool FUSED = (FWD_F || MODE == DST_FUSED_I);
  constexpr int R1 = 16, R2 = 15;
  constexpr int M = R1 * R2 * R3, N = 2 * M;
  constexpr int L1 = M / R1, L2 = M / R2, L3 = M / R3;
  constexpr int NS2 = R1, NS3 = R1 * R2;
  constexpr int WSZ = M + M / 16;              // exchange buffer incl. the skew of pass 1's output
  constexpr int SEG = dst3_seg(M, NT), NSC = M / SEG;
  constexpr int RND = (L3 + NT - 1) / NT;      // rounds of the last pass (and of everything after it)
  static_assert(L1 <= NT && L2 <= NT && NSC <= NT && SEG > 0 && L2 % 16 == 0, "plan does not fit one block");
  static_assert(NS3 == L3, "last pass must be a single sweep");
  extern __shared__ __align__(128) unsigned char smraw[];
  double *IN = reinterpret_cast<double *>(smraw);                 // N doubles: raw row (TMA target)
  const double2 *IN2 = reinterpret_cast<const double2 *>(smraw);
  double2 *Z = reinterpret_cast<double2 *>(smraw);      // M complex values: the raw row's buffer, reused between passes 2 and 3
  double2 *W = reinterpret_cast<double2 *>(smraw + (size_t)N * 8);   // exchange buffer
  double *SC = reinterpret_cast<double *>(smraw + (size_t)N * 8 + (size_t)WSZ * 16);   // M doubles: summands / running sums of the odd outputs
  double *red = SC + M;                                                                // 64 doubles
  uint64_t *mbar = reinterpret_cast<uint64_t *>(red + 64);
  const int t0 = threadIdx.x;
  const uint32_t bar = smem_u32(mbar), in_s = smem_u32(IN);
  int prev_slot = -1;      // rowsum slot of the previous row, finished one barrier later (INV)

  int item = blockIdx.x;
  if (t0 == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (item < a.nitems) {
      const int mode = FUSED ? item % a.nl : item / a.nrows, r = FUSED ? item / a.nl : item - mode * a.nrows;
      const double *in = FWD_F ? a.src : a.wrk;
      mbar_expect_tx(bar, N * 8);
      bulk_g2s(in_s, in + (size_t)mode * a.lsz + (size_t)(r + a.row0) * a.ld, N * 8, bar);
    }
  }
  __syncthreads();
  uint32_t parity = 0;
  for (; item < a.nitems; item += gridDim.x, parity ^= 1) {
    // the thread index is made opaque once per row: nothing derived from it (addresses,
    // twiddle bases, predicates) is hoisted out of the row loop into long-lived registers
    int t = t0;
    asm volatile("" : "+r"(t));
    const int lane = t & 31, wp = t >> 5;
    const int mode = FUSED ? item % a.nl : item / a.nrows, r = FUSED ? item / a.nl : item - mode * a.nrows;
    double *__restrict__ row = ((MODE == DST_FUSED_I) ? a.dst : a.wrk) + (size_t)mode * a.lsz + (size_t)(r + a.row0) * a.ld;
    // ---- pass 1 (radix 16, no twiddles) fused with the DST pre-processing (dsint.f:17-30):
    //      t_e = (x_e - x_{N-e}) + 2 sin(e pi/N) (x_e + x_{N-e}),  z_n = t_{2n} + i t_{2n+1}.
    //      Output position i = 16 t + q is stored at i + (i >> 4) = 17 t + q: conflict-free
    //      stores here, and pass 2 reads t + 160 q at t + (t >> 4) + 170 q ----
    if (t < L1) {
      double2 v1[R1];
      mbar_wait(bar, parity);
      const double2 b0 = ldg2_nohoist(a.s1base + 2 * t), b1 = ldg2_nohoist(a.s1base + 2 * t + 1);
      // fused forward transform: the row is q_k(:,j); the right-hand side of the layer is
      // q_k - beta*y_j (- ddynoc in the bottom layer), src/ocisubs.F:121-138
      double by = 0.0;
      if (FWD_F) by = a.beta * a.yrel[r + a.row0];
      const double *dd = nullptr;
      if (TOPO && mode == a.kbot) dd = a.ddyn + (size_t)(r + a.row0) * a.ld;
#pragma unroll
      for (int q = 0; q < R1; ++q) {
        const int n = t + q * L1;
        double2 xo = IN2[n];
        const int ib0 = (q == 0) ? ((t == 0) ? 0 : N - 2 * t) : N - 2 * n;
        double xb0 = IN[ib0];
        double xb1 = IN[N - 2 * n - 1];
        if (FWD_F) {
          xo.x -= by; xo.y -= by; xb0 -= by; xb1 -= by;
          if (TOPO && dd) { xo.x -= dd[2 * n]; xo.y -= dd[2 * n + 1]; xb0 -= dd[ib0]; xb1 -= dd[N - 2 * n - 1]; }
        }
        const double s0 = (q == 0) ? b0.x : fma(b0.x, a.c1[q], b0.y * a.s1[q]);
        const double s1 = (q == 0) ? b1.x : fma(b1.x, a.c1[q], b1.y * a.s1[q]);
        v1[q].x = fma(s0, xo.x + xb0, xo.x - xb0);
        v1[q].y = fma(s1, xo.y + xb1, xo.y - xb1);
      }
      if (t == 0) v1[0].x = 0.0;
      dft<R1>(v1);
#pragma unroll
      for (int q = 0; q < R1; ++q) W[t * (R1 + 1) + q] = v1[q];
    }
    __syncthreads();   // raw row consumed, pass 1 complete
    if (INV && t == 0 && prev_slot >= 0) {
      // row sum of the previous row: the per-warp partials were parked before this barrier
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < NW; ++i) sum += red[16 + i];
      a.rowsum[prev_slot] = sum;
    }
    // ---- pass 2 (radix 15): exchange buffer -> raw-row buffer (autosort stores) ----
    {
      double2 v[R2];
#pragma unroll
      for (int q = 0; q < R2; ++q) v[q] = make_double2(0.0, 0.0);   // defined on every path: nothing is carried between rows
      const int k2 = t % NS2, j0 = (t - k2) * R2 + k2;
      if (t < L2) {
        const double2 *src = W + t + (t >> 4);
#pragma unroll
        for (int q = 0; q < R2; ++q) v[q] = src[q * (L2 + L2 / 16)];
#pragma unroll
        for (int q = 1; q < R2; ++q) v[q] = cmul(v[q], ldg2_nohoist(a.tw2 + (q - 1) * NS2 + k2));
        dft<R2>(v);
      }
      // the raw row was consumed before the last barrier: its buffer takes this pass's output,
      // so no barrier separates the loads from the stores
      if (t < L2) {
#pragma unroll
        for (int q = 0; q < R2; ++q) Z[j0 + q * NS2] = v[q];
      }
    }
    __syncthreads();
    // ---- pass 3 (radix R3), raw-row buffer -> exchange buffer: butterfly tt = t + rr*NT ends with Z_k,
    //      k = tt + q*L3 (RND rounds: one with 256 threads, two with 128) ----
    double2 v[RND][R3];
#pragma unroll
    for (int rr = 0; rr < RND; ++rr) {
      const int tt = t + rr * NT;
#pragma unroll
      for (int q = 0; q < R3; ++q) v[rr][q] = make_double2(0.0, 0.0);
      if (tt < L3) {
#pragma unroll
        for (int q = 0; q < R3; ++q) v[rr][q] = Z[tt + q * L3];
        const double2 w1 = ldg2_nohoist(a.tw3base + 2 * tt), w2 = ldg2_nohoist(a.tw3base + 2 * tt + 1);
        twiddle_apply<R3>(v[rr], w1, w2);
        dft<R3>(v[rr]);
      }
      if (tt < L3) {
#pragma unroll
        for (int q = 0; q < R3; ++q) W[tt + q * L3] = v[rr][q];
      }
    }
    // generic-proxy stores of pass 2 into the buffer the bulk copy is about to overwrite
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (t == 0) {      // pass 3 has read the ping-pong buffer: the next raw row may land in it
      const int nxt = item + gridDim.x;
      if (nxt < a.nitems) {
        const int m2 = FUSED ? nxt % a.nl : nxt / a.nrows, r2 = FUSED ? nxt / a.nl : nxt - m2 * a.nrows;
        const double *in = FWD_F ? a.src : a.wrk;
        mbar_expect_tx(bar, N * 8);
        bulk_g2s(in_s, in + (size_t)m2 * a.lsz + (size_t)(r2 + a.row0) * a.ld, N * 8, bar);
      }
    }
    // ---- real post-processing: even outputs -Im F_k stay in registers, Re F_k is the
    //      summand of the odd outputs ----
    double ev[RND][R3];
#pragma unroll
    for (int rr = 0; rr < RND; ++rr) {
      const int tt = t + rr * NT;
      double cs[R3];
#pragma unroll
      for (int q = 0; q < R3; ++q) ev[rr][q] = cs[q] = 0.0;
      if (tt < L3) {
        const double2 wb = ldg2_nohoist(a.wnbase + tt);
#pragma unroll
        for (int q = 0; q < R3; ++q) {
          const int k = tt + q * L3;
          const double2 zb = (q == 0) ? ((tt == 0) ? v[rr][0] : W[M - tt]) : W[M - k];
          const double2 w = (q == 0) ? wb : cmul(wb, a.wnr[q]);
          const double2 F = real_post(v[rr][q], zb, w);
          ev[rr][q] = -F.y;
          cs[q] = F.x;
          asm volatile("" ::: "memory");   // keep the partner loads in step with their use (register budget)
        }
        if (tt == 0) {
          ev[rr][0] = 0.0;
          cs[0] = 0.5 * (v[rr][0].x + v[rr][0].y);   // out_1 = F_0 / 2
        }
      }
      if (tt < L3) {
#pragma unroll
        for (int q = 0; q < R3; ++q) SC[tt + q * L3] = cs[q];
      }
    }
    __syncthreads();
    // ---- running sum over k (dsint.f:33-37): contiguous segment per thread ----
    {
      double sg[SEG];
#pragma unroll
      for (int q = 0; q < SEG; ++q) sg[q] = 0.0;
      double tot = 0.0;
      if (t < NSC) {
        const double2 *src = reinterpret_cast<const double2 *>(SC + t * SEG);
#pragma unroll
        for (int q = 0; q < SEG / 2; ++q) {
          const double2 c2 = src[q];
          tot += c2.x;
          sg[2 * q] = tot;
          tot += c2.y;
          sg[2 * q + 1] = tot;
        }
      }
      double inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double nb = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += nb;
      }
      if (lane == 31) red[wp] = inc;
      __syncthreads();
      double off = inc - tot;
#pragma unroll
      for (int i = 0; i < NW; ++i)
        if (i < wp) off += red[i];
      if (t < NSC) {
        double2 *dst = reinterpret_cast<double2 *>(SC + t * SEG);
#pragma unroll
        for (int q = 0; q < SEG / 2; ++q) dst[q] = make_double2(sg[2 * q] + off, sg[2 * q + 1] + off);
      }
    }
    __syncthreads();
    // ---- interleave and store: row[2k] = even_k, row[2k+1] = odd_k ----
    double part = 0.0;
    if (MODE == DST_FUSED_I) {
      // coef comes from the kernel launched before this one, which may still be running
      if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
      // fused inverse transform: the row is sum_m ctm2l(m,k) wrk_m(:,j) already (the tridiagonal
      // kernel projected in spectral space); add the homogeneous solutions of the baroclinic
      // modes, sum_m ctm2l(m,k) hclco(m-1) ochom(:,j,m-1), and store p_k (src/ocisubs.F:377-401)
#pragma unroll
      for (int rr = 0; rr < RND; ++rr) {
        const int tt = t + rr * NT;
        if (tt < L3) {
          double2 *__restrict__ out = reinterpret_cast<double2 *>(row);
          const double2 *hrow = reinterpret_cast<const double2 *>(a.hom + (size_t)(r + a.row0) * a.ld);
          double2 acc[R3];
#pragma unroll
          for (int q = 0; q < R3; ++q) acc[q] = make_double2(ev[rr][q], SC[tt + q * L3]);
          for (int mm = 1; mm < a.nl; ++mm) {
            const double hc = a.coef[mm - 1], cm = a.ctm2l[mm + a.nl * mode];
            const double2 *h2 = hrow + (size_t)(mm - 1) * (a.lsz / 2);
#pragma unroll
            for (int q = 0; q < R3; ++q) {
              const double2 h = __ldg(h2 + tt + q * L3);
              acc[q].x = fma(cm, hc * h.x, acc[q].x);
              acc[q].y = fma(cm, hc * h.y, acc[q].y);
            }
          }
#pragma unroll
          for (int q = 0; q < R3; ++q) out[tt + q * L3] = acc[q];
        }
      }
      if (t == NT - 1) {      // the eastern wall column (the transform covers columns 0 .. nxp-2)
        double e = 0.0;
        for (int mm = 1; mm < a.nl; ++mm)
          e = fma(a.ctm2l[mm + a.nl * mode], a.coef[mm - 1] * a.hom[(size_t)(mm - 1) * a.lsz + (size_t)(r + a.row0) * a.ld + a.nxp - 1], e);
        row[a.nxp - 1] = e;
      }
    } else {
#pragma unroll
      for (int rr = 0; rr < RND; ++rr) {
        const int tt = t + rr * NT;
        if (tt < L3) {
          double2 *__restrict__ out = reinterpret_cast<double2 *>(row);
#pragma unroll
          for (int q = 0; q < R3; ++q) {
            const int k = tt + q * L3;
            const double od = SC[k];
            out[k] = make_double2(ev[rr][q], od);
            if (INV) part += ev[rr][q] + od;
          }
        }
      }
    }
    if (INV) {
      // xintp row sum (intsubs.f:105-112): warp partials now, the eight-term sum after the next
      // barrier (the next row's first one, or the one below the loop); W and SC are free for
      // the next row without a barrier here because its first pass only writes W
      if (t == 0) row[a.nxp - 1] = 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
      if (lane == 0) red[16 + wp] = part;
      prev_slot = mode * a.nyp + (r + a.row0);
    }
  }
  if (MODE == DST_FUSED_I) {
    if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    // wall rows of the new pressure: the inhomogeneous solution vanishes there, the homogeneous
    // solutions do not (src/ocisubs.F:377-401 on rows 1 and nypo); a slice of columns per block
    for (int side = 0; side < 2; ++side) {
      if (!(side == 0 ? a.wall_s : a.wall_n)) continue;
      const size_t ro = (size_t)(side == 0 ? 0 : a.nyp - 1) * a.ld;
      for (int i = blockIdx.x * NT + t0; i < a.nxp; i += gridDim.x * NT)
        for (int k = 0; k < a.nl; ++k) {
          double e = 0.0;
          for (int mm = 1; mm < a.nl; ++mm) e = fma(a.ctm2l[mm + a.nl * k], a.coef[mm - 1] * a.hom[(size_t)(mm - 1) * a.lsz + ro + i], e);
          a.dst[(size_t)k * a.lsz + ro + i] = e;
        }
    }
  }
  if (INV) {
    __syncthreads();
    if (t0 == 0 && prev_slot >= 0) {
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < NW; ++i) sum += red[16 + i];
      a.rowsum[prev_slot] = sum;
    }
  }
}

// --------------------------------------------------------------------------------------
// y-direction: partitioned tridiagonal solve
// --------------------------------------------------------------------------------------
struct TriArgs {
  int ld, nyp, nk, koff, nchunk, lastlen, nmodes, row0;
  size_t lsz;
  int use_yx, nranks;
  int slab_phase;              // k_tri_reduced on y-slabs: 1 = also emit the slab's first/last rows, 2 = outer neighbours are the adjacent slabs' rows
  double *slab_send;           // [nmodes][2][ld]
  const double *slab_outer;    // [nmodes][2][ld]
  PeerCtx peer;                // slab_phase 1 over peer memory: the rows go to the ranks' mailboxes instead of slab_send
  unsigned int *ticket;
  double a, ftnorm;
  double *wrk;
  const double *bcoef;
  double *binv, *vl, *vll, *pt, *fg, *yx;
};

// one thread per (column, mode): elimination reciprocals, spike vectors, interface pivots
__global__ void k_tri_tables(TriArgs t) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int mode = blockIdx.y;
  if (s >= t.nk) return;
  const int col = t.koff + s;
  const double b = t.bcoef[(size_t)mode * t.ld + col], a = t.a;
  double binv[TRI_L], w[TRI_L];
  binv[0] = 1.0 / b;
  for (int j = 1; j < TRI_L; ++j) binv[j] = 1.0 / (b - a * (a * binv[j - 1]));
  const size_t tb = ((size_t)mode * TRI_L) * t.ld + col;
  for (int j = 0; j < TRI_L; ++j) t.binv[tb + (size_t)j * t.ld] = binv[j];
  double alpha = 0.0, eps = 0.0, alphal = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    const int len = pass == 0 ? TRI_L : t.lastlen;
    w[0] = (-a) * binv[0];
    for (int j = 1; j < len; ++j) w[j] = (-a * w[j - 1]) * binv[j];
    for (int j = len - 2; j >= 0; --j) w[j] = w[j] - (a * binv[j]) * w[j + 1];
    double *dst = pass == 0 ? t.vl : t.vll;
    for (int j = 0; j < TRI_L; ++j) dst[tb + (size_t)j * t.ld] = j < len ? w[j] : 0.0;
    if (pass == 0) {
      alpha = w[0];
      eps = w[TRI_L - 1];
    } else {
      alphal = w[0];
    }
  }
  // interface c (between chunks c-1 and c), c = 1..C-1: D'_c = [[1, p_c],[q_c, 1]]
  const int C = t.nchunk;
  double p = -alpha;
  for (int c = 1; c < C; ++c) {
    const double q = (c == C - 1) ? -alphal : -alpha;
    const double dinv = 1.0 / (1.0 - p * q);
    t.pt[((size_t)mode * 2 * C + c) * t.ld + col] = p;
    t.pt[((size_t)mode * 2 * C + C + c) * t.ld + col] = dinv;
    p = -alpha + eps * eps * p * dinv;
  }
}

// Chunk-local Thomas solve in registers; the elimination reciprocals stream from an
// L2-resident table.  grid (ceil(nk/128), nchunk, nmodes).
//   FINAL = false : right-hand side with zero neighbours; only the first and last values of
//                   the local solution are stored (f_c, g_c: the interface system's rhs).
//   FINAL = true  : the neighbour values found by k_tri_reduced are moved to the right-hand
//                   side of the chunk's first and last rows, so the local solve returns the
//                   rows of the global solution; stored times ftnorm (src/ocisubs.F:484-487).
template <bool FINAL>
__global__ void __launch_bounds__(128, 3) k_tri_local(TriArgs t) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, mode = blockIdx.z;
  if (s >= t.nk) return;
  const int col = t.koff + s;
  const int len = (c == t.nchunk - 1) ? t.lastlen : TRI_L;
  double *__restrict__ base = t.wrk + (size_t)mode * t.lsz + (size_t)(t.row0 + c * TRI_L) * t.ld + col;
  const double *__restrict__ bi = t.binv + ((size_t)mode * TRI_L) * t.ld + col;
  const double a = t.a;
  const int ld = t.ld;
  const size_t fb = ((size_t)mode * 2 * t.nchunk) * ld + col;
  double u[TRI_L];
#pragma unroll
  for (int j = 0; j < TRI_L; ++j) u[j] = base[(size_t)min(j, len - 1) * ld];   // rows >= len: harmless duplicates
  if (FINAL && t.use_yx) {
    const double yp = t.yx[fb + (size_t)c * ld], xn = t.yx[fb + (size_t)(t.nchunk + c) * ld];
    u[0] -= a * yp;
#pragma unroll
    for (int j = 0; j < TRI_L; ++j) u[j] = fma(-a, (j == len - 1) ? xn : 0.0, u[j]);
  }
  u[0] = u[0] * __ldg(bi);
#pragma unroll
  for (int j = 1; j < TRI_L; ++j) u[j] = (u[j] - a * u[j - 1]) * __ldg(bi + (size_t)j * ld);
#pragma unroll
  for (int j = TRI_L - 2; j >= 0; --j) {
    const double v = u[j] - (a * __ldg(bi + (size_t)j * ld)) * u[j + 1];
    u[j] = (j < len - 1) ? v : u[j];
  }
  if (FINAL) {
    const double fn = t.ftnorm;
#pragma unroll
    for (int j = 0; j < TRI_L; ++j)
      if (j < len) base[(size_t)j * ld] = fn * u[j];
  } else if (t.nchunk > 1 || t.nranks > 1) {
    t.fg[fb + (size_t)c * ld] = u[0];                    // f_c : first row of the chunk
    // g_c : last row of the chunk (the ragged last chunk ends at row len-1; its g only matters
    // to the slab coupling of multi-GPU runs)
    double gl = 0.0;
#pragma unroll
    for (int j = 0; j < TRI_L; ++j) gl = fma((j == len - 1) ? 1.0 : 0.0, u[j], gl);   // arithmetic select keeps u in registers
    t.fg[fb + (size_t)(t.nchunk + c) * ld] = gl;
  }
}

// Fused variant of k_tri_local for the box ocean (NL layers = NL modes per wavenumber in one
// thread).  The work array holds spectral LAYER rows (the forward transform of q_k - beta*y): the
// thread projects them onto the modes, r_m = f0 sum_k ctl2m(k,m) s_k (src/ocisubs.F:117-139 in
// spectral space), solves the NL chunk systems in registers and
//   FINAL = false : keeps the first/last values of each mode (the interface system's rhs);
//   FINAL = true  : projects the solution back onto the layers, l_k = sum_m ctm2l(m,k) u_m
//                   (src/ocisubs.F:377-401 without the homogeneous part, which the inverse
//                   transform adds in physical space), stores it, and leaves the block's share of
//                   the area integrals xinhom(m) of the modal solutions: the x-sum of a sine
//                   series is sum_k 2 cot(k pi / 2N) X_k over odd k, so the integral the
//                   constraint algebra needs (src/ocisubs.F:146-162) is known before the
//                   inverse transform runs.
// grid (ceil(nk/128), nchunk).
struct Mix3 {
  double f0;
  double ctl2m[NLMAX * NLMAX], ctm2l[NLMAX * NLMAX];
  const double *wsum;     // [ld] 2 cot(k pi / 2N) at odd wavenumber columns, 0 elsewhere
  double *spec;           // [nl][nchunk * gridDim.x] partial integrals (FINAL)
};

// Work item = a tile of 2*TRI3_TP wavenumber columns x one chunk; a thread is part p (of NL) of one
// column PAIR (even column first: rows are 128-byte aligned, so every global and shared access
// is a 16-byte one -- the kernel is bound by the number of load/store instructions, not by
// bytes).  Persistent blocks (two per SM) take contiguous ranges of the chunk-fastest item list,
// so a block stays on one column tile for many chunks and its table reads hit L1.  Per item
//   1. part p loads rows j = p, p+NL, ... of ALL layers of its pair, projects them onto the
//      modes and parks them in shared memory ([mode][row][pair], conflict free);
//   2. part p owns mode p of its pair (two interleaved recurrences):
//      FINAL = false: the last value of the chunk-local solution is where the forward
//        elimination ends; the first value is where the same elimination ends when it runs over
//        the rows in reverse order (symmetric Toeplitz system: same reciprocals);
//      FINAL = true : Thomas solve in place in shared memory, reciprocals from the L2-resident
//        table a group of rows ahead of the recurrence;
//   3. FINAL: part p projects rows j = p, p+NL, ... of all modes back onto the layers and stores
//      them, and the block leaves its share of the modal area integrals.
// Columns outside the solved range (the wall columns 0 and nxp-1 of the box, the row padding)
// ride along with zero reciprocals and are never stored.
constexpr int TRI3_TP = 64;                                                    // column pairs per tile (32 with four blocks per SM measured slower)
constexpr int TRI3_BLK = 2;                                                    // blocks per SM (96 KB of shared memory each)
constexpr size_t tri3_smem = sizeof(double2) * 3 * TRI_L * TRI3_TP;            // 96 KB for NL = 3
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
template <int NL, bool FINAL>
__global__ void __launch_bounds__(TRI3_TP * NL, TRI3_BLK) k_tri3(TriArgs t, Mix3 mx, int ntx, int order) {
  constexpr int TP = TRI3_TP, G = 8;
  static_assert(TRI_L % G == 0, "row groups");
  extern __shared__ double2 us_raw2[];
  double2 (*us)[TRI_L][TP] = reinterpret_cast<double2 (*)[TRI_L][TP]>(us_raw2);      // [NL][TRI_L][TP]
  __shared__ double red[NL][TRI3_TP / 32];
  const int tp = threadIdx.x % TP, p = threadIdx.x / TP;      // warps are uniform in p
  const double a = t.a;
  const int ld = t.ld;
  const int nitems = ntx * t.nchunk;
  constexpr int NJ = (TRI_L + NL - 1) / NL;
  const int per = (nitems + (int)gridDim.x - 1) / (int)gridDim.x;
  const int first = order ? (int)blockIdx.x : blockIdx.x * per, last = order ? nitems : min(first + per, nitems);
  const int step = order ? (int)gridDim.x : 1;
  for (int item = first; item < last; item += step) {
    // order 0: chunk-fastest items in contiguous ranges (a block walks down one column tile);
    // order 1: tile-fastest items, strided (at any time the blocks work on neighbouring tiles of a
    // few chunk rows, i.e. on one narrow band of rows per layer)
    const int bx = order ? item % ntx : item / t.nchunk, c = order ? item / ntx : item - bx * t.nchunk;
    int c0 = 2 * (bx * TP + tp);                               // even column of the pair
    const bool inrow = c0 < ld;
    if (!inrow) c0 = 0;
    const bool live0 = inrow && c0 >= t.koff && c0 < t.koff + t.nk, live1 = inrow && c0 + 1 >= t.koff && c0 + 1 < t.koff + t.nk;
    const bool lastc = (c == t.nchunk - 1);
    const int len = lastc ? t.lastlen : TRI_L;
    double *__restrict__ base = t.wrk + (size_t)(t.row0 + c * TRI_L) * ld + c0;
    // ---- 1. layers -> modes for this part's rows
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = p + jj * NL;
      if (j < TRI_L) {
        double2 sk[NL];
        const size_t ro = (size_t)min(j, len - 1) * ld;      // rows >= len: harmless duplicates of the last row
#pragma unroll
        for (int k = 0; k < NL; ++k) sk[k] = *reinterpret_cast<const double2 *>(base + (size_t)k * t.lsz + ro);
#pragma unroll
        for (int m = 0; m < NL; ++m) {
          double ax = 0.0, ay = 0.0;
#pragma unroll
          for (int k = 0; k < NL; ++k) {
            ax = ax + mx.ctl2m[k + NL * m] * sk[k].x;
            ay = ay + mx.ctl2m[k + NL * m] * sk[k].y;
          }
          us[m][j][tp] = make_double2(live0 ? mx.f0 * ax : 0.0, live1 ? mx.f0 * ay : 0.0);
        }
      }
    }
    __syncthreads();
    // ---- 2. mode p of this pair
    const int m = p;
    const double *__restrict__ bi = t.binv + ((size_t)m * TRI_L) * ld + c0;
    const size_t fb = ((size_t)m * 2 * t.nchunk) * ld + c0;
    if (!FINAL) {
      double2 d = make_double2(0.0, 0.0), e = make_double2(0.0, 0.0);
      double2 bq[G];
#pragma unroll
      for (int u = 0; u < G; ++u) bq[u] = ldg2(bi + (size_t)u * ld);
#pragma unroll
      for (int g = 0; g < TRI_L / G; ++g) {
        double2 bn[G];
#pragma unroll
        for (int u = 0; u < G; ++u) bn[u] = ldg2(bi + (size_t)min((g + 1) * G + u, TRI_L - 1) * ld);
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const int j = g * G + u;
          const double2 r = us[m][j][tp], rr = us[m][max(len - 1 - j, 0)][tp];
          if (j < len) {
            d = make_double2((r.x - a * d.x) * bq[u].x, (r.y - a * d.y) * bq[u].y);
            e = make_double2((rr.x - a * e.x) * bq[u].x, (rr.y - a * e.y) * bq[u].y);
          }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) bq[u] = bn[u];
      }
      if (inrow && (t.nchunk > 1 || t.nranks > 1)) {
        *reinterpret_cast<double2 *>(t.fg + fb + (size_t)c * ld) = e;                      // f_c : first row of the chunk-local solution
        *reinterpret_cast<double2 *>(t.fg + fb + (size_t)(t.nchunk + c) * ld) = d;         // g_c : its last row
      }
      __syncthreads();      // the buffer is reused by the next item
      continue;
    }
    {
      if (t.use_yx) {
        const double2 yp = ldg2(t.yx + fb + (size_t)c * ld), xn = ldg2(t.yx + fb + (size_t)(t.nchunk + c) * ld);
        double2 v0 = us[m][0][tp];
        v0.x -= a * yp.x; v0.y -= a * yp.y;
        us[m][0][tp] = v0;
        double2 v1 = us[m][len - 1][tp];
        v1.x = fma(-a, xn.x, v1.x); v1.y = fma(-a, xn.y, v1.y);
        us[m][len - 1][tp] = v1;
      }
      double2 bq[G];
#pragma unroll
      for (int u = 0; u < G; ++u) bq[u] = ldg2(bi + (size_t)u * ld);
      double2 prev = make_double2(0.0, 0.0);
#pragma unroll
      for (int g = 0; g < TRI_L / G; ++g) {
        double2 bn[G];
#pragma unroll
        for (int u = 0; u < G; ++u) bn[u] = ldg2(bi + (size_t)min((g + 1) * G + u, TRI_L - 1) * ld);
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const int j = g * G + u;
          const double2 r = us[m][j][tp];
          prev = make_double2((r.x - a * prev.x) * bq[u].x, (r.y - a * prev.y) * bq[u].y);
          us[m][j][tp] = prev;      // rows >= len hold values nobody reads
        }
#pragma unroll
        for (int u = 0; u < G; ++u) bq[u] = bn[u];
      }
      // back substitution from row len-1 downwards, times ftnorm on the way out (src/ocisubs.F:484-487)
      const double fn = t.ftnorm;
      double2 nxt = us[m][len - 1][tp];
      double2 sm = make_double2(fn * nxt.x, fn * nxt.y);
      us[m][len - 1][tp] = sm;
#pragma unroll
      for (int u = 0; u < G; ++u) bq[u] = ldg2(bi + (size_t)(TRI_L - 1 - u) * ld);
#pragma unroll
      for (int g = 0; g < TRI_L / G; ++g) {
        double2 bn[G];
#pragma unroll
        for (int u = 0; u < G; ++u) bn[u] = ldg2(bi + (size_t)max(TRI_L - 1 - (g + 1) * G - u, 0) * ld);
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const int j = TRI_L - 1 - g * G - u;
          if (j < len - 1) {
            const double2 r = us[m][j][tp];
            nxt = make_double2(r.x - (a * bq[u].x) * nxt.x, r.y - (a * bq[u].y) * nxt.y);
            const double2 o = make_double2(fn * nxt.x, fn * nxt.y);
            us[m][j][tp] = o;
            sm.x += o.x; sm.y += o.y;
          }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) bq[u] = bn[u];
      }
      // the block's share of the area integral of mode p: fixed-order reduction over its columns
      const double2 w = ldg2(mx.wsum + c0);
      double v = inrow ? (w.x * sm.x + w.y * sm.y) : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if ((tp & 31) == 0) red[p][tp >> 5] = v;
    }
    __syncthreads();
    if (tp == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < TRI3_TP / 32; ++w) tot += red[p][w];
      mx.spec[(size_t)p * nitems + item] = tot;
    }
    // ---- 3. modes -> layers for this part's rows
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = p + jj * NL;
      if (j < len) {
        double2 um[NL];
#pragma unroll
        for (int mm = 0; mm < NL; ++mm) um[mm] = us[mm][j][tp];
#pragma unroll
        for (int k = 0; k < NL; ++k) {
          double ax = 0.0, ay = 0.0;
#pragma unroll
          for (int mm = 0; mm < NL; ++mm) {
            ax = ax + mx.ctm2l[mm + NL * k] * um[mm].x;
            ay = ay + mx.ctm2l[mm + NL * k] * um[mm].y;
          }
          double *dst = base + (size_t)k * t.lsz + (size_t)j * ld;
          if (live0 && live1) *reinterpret_cast<double2 *>(dst) = make_double2(ax, ay);
          else if (live0) dst[0] = ax;
          else if (live1) dst[1] = ay;
        }
      }
    }
    __syncthreads();      // the buffer is reused by the next item
  }
}

struct SlabArgs {
  int ld, nk, koff, nchunk, lastlen, nmodes, nranks, rank;
  int nrows_of[16];          // solved rows of every slab
  double a;
  const double *bcoef, *vl, *vll;
  double *fg, *yx;
  double *ae;                // [nmodes][nranks][2][ld]
  double *send;              // [nmodes][2][ld]
  const double *all;         // [nranks][nmodes][2][ld]
  double *outer;             // [nmodes][2][ld]
  PeerCtx peer;              // peer-memory transport: `all` is this rank's mailbox, filled by the peers' k_slab_push
  int *peer_err;
};

// y-slab coupling, per wavenumber column (the comment block "y-slab coupling" below explains
// the algebra).  The inter-slab system, the neighbour rows of this slab, and their effect on the
// first/last chunk's f, g.  Interface i (between slabs i-1 and i) couples Y_{i-1} and X_i:
//   Y_{i-1} - alpha_{i-1} X_i = G_{i-1} + eps_{i-1} Y_{i-2},   X_i - alpha_i Y_{i-1} = F_i + eps_i X_{i+1}
// which is an interleaved tridiagonal system: one forward sweep expressing
// Y_{i-1} = P_i + Q_i X_{i+1}, X_i = xc_i + xq_i X_{i+1}, one back substitution.
// peer-memory transport: the rows of every rank must have landed in the mailbox (whole block)
__device__ __forceinline__ void slab_wait(const SlabArgs &t) {
  if (t.peer.n) {
    if ((int)threadIdx.x < t.peer.n)
      peer_wait(reinterpret_cast<const volatile unsigned long long *>(t.peer.box[t.peer.rank] + peer_off_flagf(t.peer.n)) + threadIdx.x,
                t.peer.epoch, t.peer_err);
    __syncthreads();
    __threadfence_system();
  }
}
__device__ __forceinline__ void slab_solve_col(const SlabArgs &t, int s, int mode) {
  const int col = t.koff + s, N = t.nranks, ld = t.ld;
  double xc[8], xq[8], P[8], Q[8];
  double Pp = 0.0, Qp = 0.0;
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    if (i < N) {
      const double a_lo = t.ae[(((size_t)mode * N + i - 1) * 2 + 0) * ld + col], e_lo = t.ae[(((size_t)mode * N + i - 1) * 2 + 1) * ld + col];
      const double a_hi = t.ae[(((size_t)mode * N + i) * 2 + 0) * ld + col], e_hi = t.ae[(((size_t)mode * N + i) * 2 + 1) * ld + col];
      const double G = __ldcg(t.all + (((size_t)(i - 1) * t.nmodes + mode) * 2 + 1) * ld + col);
      const double F = __ldcg(t.all + (((size_t)i * t.nmodes + mode) * 2 + 0) * ld + col);
      const double gp = G + e_lo * Pp, A = a_lo + e_lo * Qp;
      const double den = 1.0 / (1.0 - a_hi * A);
      xc[i] = (F + a_hi * gp) * den;
      xq[i] = e_hi * den;
      P[i] = gp + A * xc[i];
      Q[i] = A * xq[i];
      Pp = P[i];
      Qp = Q[i];
    }
  }
  double xnext_run = 0.0, yprev = 0.0, xnext = 0.0;
#pragma unroll
  for (int i = 7; i >= 1; --i) {
    if (i < N) {
      const double X = xc[i] + xq[i] * xnext_run;      // X_i
      const double Y = P[i] + Q[i] * xnext_run;        // Y_{i-1}
      if (i == t.rank) yprev = Y;
      if (i == t.rank + 1) xnext = X;
      xnext_run = X;
    }
  }
  t.outer[((size_t)mode * 2 + 0) * ld + col] = yprev;
  t.outer[((size_t)mode * 2 + 1) * ld + col] = xnext;
  // neighbour rows act on the first chunk through its left spike and on the last chunk
  // through its right spike (mirror of its left spike)
  const int C = t.nchunk;
  if (C > 1) {
    const size_t fb = ((size_t)mode * 2 * C) * ld + col, tb = ((size_t)mode * TRI_L) * ld + col;
    const double alpha = t.vl[tb], eps = t.vl[tb + (size_t)(TRI_L - 1) * ld];
    const double alphal = t.vll[tb], epsl = t.vll[tb + (size_t)(t.lastlen - 1) * ld];
    t.fg[fb] += yprev * alpha;
    t.fg[fb + (size_t)C * ld] += yprev * eps;
    t.fg[fb + (size_t)(C - 1) * ld] += xnext * epsl;
    t.fg[fb + (size_t)(C + C - 1) * ld] += xnext * alphal;
  }
}

// interface system: one thread per (column, mode), block Thomas over the chunks.  All
// loads are independent of the recurrence (separate output array), so they pipeline.
__global__ void __launch_bounds__(128) k_tri_reduced(TriArgs t, SlabArgs sa) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int mode = blockIdx.y;
  if (t.slab_phase == 2) {
    // y-slabs, second pass: first the inter-slab system of this column (it fixes the outer
    // neighbours and adds their effect to the first/last chunk's f, g), then the chunk
    // interfaces again -- one launch for what used to be k_slab_solve + k_tri_reduced
    slab_wait(sa);
    if (s < t.nk) slab_solve_col(sa, s, mode);
  }
  if (s < t.nk) {
  const int col = t.koff + s;
  const int C = t.nchunk, ld = t.ld;
  const size_t tb = ((size_t)mode * TRI_L) * ld + col;
  const double alpha = t.vl[tb], eps = t.vl[tb + (size_t)(TRI_L - 1) * ld], alphal = t.vll[tb];
  const size_t fb = ((size_t)mode * 2 * C) * ld + col;
  // (not __restrict__: on y-slabs the inter-slab solve above has just updated four of these values)
  const double *f = t.fg + fb;
  const double *g = t.fg + fb + (size_t)C * ld;
  double *__restrict__ yp = t.yx + fb;
  double *__restrict__ xn = t.yx + fb + (size_t)C * ld;
  const double *__restrict__ pt = t.pt + ((size_t)mode * 2 * C) * ld + col;   // p_c
  const double *__restrict__ di = pt + (size_t)C * ld;                        // 1/(1 - p_c q_c)
  // Both sweeps are latency bound (one thread per wavenumber, ~C dependent steps), so the
  // loads of PF chunks are issued together ahead of their dependent chain: one exposed
  // memory latency per PF steps instead of one per step.
  constexpr int PF = 24;
  // forward elimination: h0_c is parked in yp[c] (overwritten by the back substitution)
  double h0 = g[0], h1 = f[ld], p = pt[ld], dinv = di[ld];
  yp[ld] = h0;
  for (int c0 = 2; c0 < C; c0 += PF) {
    double gc[PF], fc[PF], pc[PF], dc[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int c = min(c0 + u, C - 1);
      gc[u] = g[(size_t)(c - 1) * ld]; fc[u] = f[(size_t)c * ld];
      pc[u] = pt[(size_t)c * ld]; dc[u] = di[(size_t)c * ld];
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (c0 + u < C) {
        const double t0 = (h0 - p * h1) * dinv;
        h0 = gc[u] + eps * t0;
        yp[(size_t)(c0 + u) * ld] = h0;
        h1 = fc[u];
        p = pc[u];
        dinv = dc[u];
      }
    }
  }
  // back substitution
  double xnext = 0.0;
  xn[(size_t)(C - 1) * ld] = 0.0;
  for (int c0 = C - 1; c0 >= 1; c0 -= PF) {
    double pc[PF], dc[PF], hh[PF], fc[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int c = max(c0 - u, 1);
      pc[u] = pt[(size_t)c * ld]; dc[u] = di[(size_t)c * ld];
      hh[u] = yp[(size_t)c * ld]; fc[u] = f[(size_t)c * ld];
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int c = c0 - u;
      if (c >= 1) {
        const double qc = (c == C - 1) ? -alphal : -alpha;
        const double hh1 = fc[u] + eps * xnext;
        const double y = (hh[u] - pc[u] * hh1) * dc[u];    // y_{c-1}: last row of chunk c-1
        const double x = (hh1 - qc * hh[u]) * dc[u];       // x_c    : first row of chunk c
        yp[(size_t)c * ld] = y;
        xn[(size_t)(c - 1) * ld] = x;
        xnext = x;
      }
    }
  }
  yp[0] = 0.0;
  if (t.slab_phase == 1) {
    // first and last rows of the slab-local solution: f_0 + eps x_1, g_{C-1} + epsl y_{C-2}
    const double epsl = t.vll[tb + (size_t)(t.lastlen - 1) * ld];
    const double F = f[0] + eps * xnext, G = g[(size_t)(C - 1) * ld] + epsl * yp[(size_t)(C - 1) * ld];
    if (t.peer.n) {
      // peer-memory transport: the two rows go straight into every rank's mailbox
      const size_t o = peer_off_fg(t.peer.n, t.peer.fglen, (int)(t.peer.epoch & 1ull), t.peer.rank) + ((size_t)mode * 2) * ld + col;
      for (int r = 0; r < t.peer.n; ++r) {
        t.peer.box[r][o] = F;
        t.peer.box[r][o + ld] = G;
      }
    } else {
      t.slab_send[((size_t)mode * 2 + 0) * ld + col] = F;
      t.slab_send[((size_t)mode * 2 + 1) * ld + col] = G;
    }
  } else if (t.slab_phase == 2) {
    yp[0] = t.slab_outer[((size_t)mode * 2 + 0) * ld + col];
    xn[(size_t)(C - 1) * ld] = t.slab_outer[((size_t)mode * 2 + 1) * ld + col];
  }
  }
  if (t.slab_phase == 1 && t.peer.n) {
    // the block that finishes last publishes the epoch to every rank (k_slab_solve waits for it)
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(t.ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x == 0) *t.ticket = 0u;
    if ((int)threadIdx.x < t.peer.n)
      reinterpret_cast<volatile unsigned long long *>(t.peer.box[threadIdx.x] + peer_off_flagf(t.peer.n))[t.peer.rank] = t.peer.epoch;
  }
}

// --------------------------------------------------------------------------------------
// y-slab coupling (multi-GPU): a slab is one more level of the same partition.  Its local
// solution (zero neighbours) has first/last rows F_s, G_s; the true rows obey
//   X_s = F_s + alpha_s Y_{s-1} + eps_s X_{s+1},   Y_s = G_s + eps_s Y_{s-1} + alpha_s X_{s+1}
// with (alpha_s, eps_s) the first/last values of the slab's left spike (Toeplitz: the right
// spike is its mirror image).  F, G are all-gathered (2 rows per mode and rank), every rank
// solves the 2*nranks unknowns per wavenumber redundantly and keeps its neighbours' rows.
// --------------------------------------------------------------------------------------
// alpha_s = -a (T^-1)_{00}, eps_s = -a (T^-1)_{n-1,0}: one forward elimination per slab length
__global__ void k_slab_spikes(SlabArgs t) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x, mode = blockIdx.y;
  if (s >= t.nk) return;
  const int col = t.koff + s;
  const double b = t.bcoef[(size_t)mode * t.ld + col], a = t.a;
  for (int r = 0; r < t.nranks; ++r) {
    const int n = t.nrows_of[r];
    double denom = b, cp = a / b, dp = -a / b;
    for (int j = 1; j < n; ++j) {
      denom = b - a * cp;
      cp = a / denom;
      dp = (-a * dp) / denom;
    }
    t.ae[(((size_t)mode * t.nranks + r) * 2 + 0) * t.ld + col] = -a / denom;
    t.ae[(((size_t)mode * t.nranks + r) * 2 + 1) * t.ld + col] = dp;
  }
}

// first and last rows of the slab-local solution from the chunk interface solve
__global__ void k_slab_fg(SlabArgs t) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x, mode = blockIdx.y;
  if (s >= t.nk) return;
  const int col = t.koff + s, C = t.nchunk, ld = t.ld;
  const size_t fb = ((size_t)mode * 2 * C) * ld + col, tb = ((size_t)mode * TRI_L) * ld + col;
  double F = t.fg[fb], G = t.fg[fb + (size_t)(C + C - 1) * ld];
  if (C > 1) {
    const double eps = t.vl[tb + (size_t)(TRI_L - 1) * ld], epsl = t.vll[tb + (size_t)(t.lastlen - 1) * ld];
    F += eps * t.yx[fb + (size_t)C * ld];                 // xn[0]: first row of chunk 1
    G += epsl * t.yx[fb + (size_t)(C - 1) * ld];          // yp[C-1]: last row of chunk C-2
  }
  t.send[((size_t)mode * 2 + 0) * ld + col] = F;
  t.send[((size_t)mode * 2 + 1) * ld + col] = G;
}

// the inter-slab system, the neighbour rows of this slab, and their effect on the first/last
// chunk's f, g.  Interface i (between slabs i-1 and i) couples Y_{i-1} and X_i:
//   Y_{i-1} - alpha_{i-1} X_i = G_{i-1} + eps_{i-1} Y_{i-2},   X_i - alpha_i Y_{i-1} = F_i + eps_i X_{i+1}
// which is an interleaved tridiagonal system: one forward sweep expressing
// Y_{i-1} = P_i + Q_i X_{i+1}, X_i = xc_i + xq_i X_{i+1}, one back substitution.
__global__ void k_slab_solve(SlabArgs t) {
  slab_wait(t);
  const int s = blockIdx.x * blockDim.x + threadIdx.x, mode = blockIdx.y;
  if (s >= t.nk) return;
  slab_solve_col(t, s, mode);
}

// after the second interface solve: the outermost neighbours are the adjacent slabs' rows
__global__ void k_slab_outer(SlabArgs t) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x, mode = blockIdx.y;
  if (s >= t.nk) return;
  const int col = t.koff + s, C = t.nchunk, ld = t.ld;
  const size_t fb = ((size_t)mode * 2 * C) * ld + col;
  t.yx[fb] = t.outer[((size_t)mode * 2 + 0) * ld + col];
  t.yx[fb + (size_t)(C + C - 1) * ld] = t.outer[((size_t)mode * 2 + 1) * ld + col];
}

__global__ void k_zero_rows(double *wrk, size_t lsz, int ld, int nyp, int nxp, int nmodes, int south, int north) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nxp) return;
  for (int m = 0; m < nmodes; ++m) {
    if (south) wrk[(size_t)m * lsz + i] = 0.0;
    if (north) wrk[(size_t)m * lsz + (size_t)(nyp - 1) * ld + i] = 0.0;
  }
}

// --------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------
// ---- fast DST path: instantiated plans, tables, launch ----
template <int R3>
static void dst3_launch_t(qgcm_model *md, HelmPlan &hp, const Dst3Args &a, int mode) {
  constexpr int M = 16 * 15 * R3;
  const size_t smem = (size_t)M * 16 + (size_t)(M + M / 16) * 16 + (size_t)M * 8 + 64 * 8 + 16;
  auto kf = k_dst3<R3, DST_PLAIN_F>;
  auto ki = k_dst3<R3, DST_PLAIN_I>;
  auto kff = k_dst3<R3, DST_FUSED_F>;
  auto kfi = k_dst3<R3, DST_FUSED_I>;
  auto kfft = k_dst3<R3, DST_FUSED_FT>;
  if (!hp.fast_attr) {
    QG_CUDA(cudaFuncSetAttribute(kfft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QG_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QG_CUDA(cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QG_CUDA(cudaFuncSetAttribute(kff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QG_CUDA(cudaFuncSetAttribute(kfi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hp.fast_attr = 1;
  }
  constexpr int NT = dst3_threads(R3);
  const int grid = std::min(a.nitems, hp.fast_grid / 2 * dst3_blocks(R3));      // persistent blocks, two or four per SM
  switch (mode) {
    case DST_PLAIN_F: QG_LAUNCH(md, "k_xform", grid, NT, smem, kf, a); break;
    case DST_PLAIN_I: QG_LAUNCH(md, "k_xform_inv", grid, NT, smem, ki, a); break;
    case DST_FUSED_F:
      if (a.ddyn) QG_LAUNCH(md, "k_xform", grid, NT, smem, kfft, a);
      else QG_LAUNCH(md, "k_xform", grid, NT, smem, kff, a);
      break;
    default:
      // the fused inverse transform needs the constraint coefficients (k_inv_scalars, the launch before it)
      // only in its epilogues: it may start beside that kernel and waits for it there (griddepcontrol.wait)
      if (a.pdl) QG_LAUNCH_PDL(md, "k_xform_inv", grid, NT, smem, kfi, a);
      else QG_LAUNCH(md, "k_xform_inv", grid, NT, smem, kfi, a);
      break;
  }
}

// box decks whose half length is 240*R3 run the three-pass plan (16, 15, R3)
static int dst3_r3(const HelmPlan &hp) {
  if (hp.kind != 0 || hp.m % 240 != 0) return 0;
  const int r3 = hp.m / 240;
  return (r3 == 2 || r3 == 3 || r3 == 4 || r3 == 5 || r3 == 6 || r3 == 8 || r3 == 10) ? r3 : 0;
}

static void dst3_launch(qgcm_model *md, HelmPlan &hp, double *wrk, size_t lsz, int nmodes, int mode, const FusedInv *fz = nullptr) {
  Dst3Args a = {};
  a.nitems = nmodes * hp.nrows; a.nrows = hp.nrows; a.ld = hp.ld; a.nyp = hp.nyp; a.nxp = hp.nxp; a.row0 = hp.row0; a.lsz = lsz;
  a.wrk = wrk; a.rowsum = hp.rowsum;
  a.s1base = hp.s1base; a.tw2 = hp.tw2; a.tw3base = hp.tw3base; a.wnbase = hp.wnbase;
  for (int i = 0; i < 16; ++i) { a.c1[i] = hp.c1[i]; a.s1[i] = hp.s1c[i]; a.wnr[i] = hp.wnr[i]; }
  a.nl = nmodes; a.kbot = nmodes - 1;
  a.wall_s = hp.wall_s; a.wall_n = hp.wall_n;
  a.pdl = (mode == DST_FUSED_I && !md->prof && env_int("QGCM_PDL", 1)) ? 1 : 0;
  if (fz) {
    a.src = fz->q; a.dst = fz->pnew; a.yrel = fz->yrel; a.beta = fz->beta; a.ddyn = fz->ddyn; a.hom = fz->hom; a.coef = fz->coef;
    for (int i = 0; i < NLMAX * NLMAX; ++i) a.ctm2l[i] = fz->ctm2l[i];
  }
  switch (hp.fast) {
    case 2: dst3_launch_t<2>(md, hp, a, mode); break;
    case 3: dst3_launch_t<3>(md, hp, a, mode); break;
    case 4: dst3_launch_t<4>(md, hp, a, mode); break;
    case 5: dst3_launch_t<5>(md, hp, a, mode); break;
    case 6: dst3_launch_t<6>(md, hp, a, mode); break;
    case 8: dst3_launch_t<8>(md, hp, a, mode); break;
    case 10: dst3_launch_t<10>(md, hp, a, mode); break;
    default: throw std::runtime_error("helmholtz: no fast DST plan");
  }
}

static void dst3_plan(qgcm_model *md, HelmPlan &hp) {
  hp.fast = dst3_r3(hp);
  if (!hp.fast) return;
  const int R1 = 16, R2 = 15, R3 = hp.fast;
  const int M = hp.m, N = hp.n, L1 = M / R1, L3 = M / R3;
  const long double PI_L = 3.141592653589793238462643383279502884L;
  std::vector<double2> s1b(2 * L1), t2, t3(2 * L3), wb(L3);
  for (int j = 0; j < L1; ++j)
    for (int e = 0; e < 2; ++e) {
      const long double ang = PI_L * (2 * j + e) / N;
      s1b[2 * j + e] = make_double2((double)(2.0L * sinl(ang)), (double)(2.0L * cosl(ang)));
    }
  for (int q = 1; q < R2; ++q)
    for (int k = 0; k < R1; ++k) {
      const long double ang = -2.0L * PI_L * (long double)q * k / ((long double)R1 * R2);
      t2.push_back(make_double2((double)cosl(ang), (double)sinl(ang)));
    }
  for (int t = 0; t < L3; ++t) {
    for (int e = 1; e <= 2; ++e) {
      const long double ang = -2.0L * PI_L * e * t / M;
      t3[2 * t + e - 1] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    const long double ang = -2.0L * PI_L * t / N;
    wb[t] = make_double2((double)cosl(ang), (double)sinl(ang));
  }
  for (int q = 0; q < 16; ++q) {
    hp.c1[q] = (double)cosl(PI_L * q / R1);
    hp.s1c[q] = (double)sinl(PI_L * q / R1);
    const long double ang = -2.0L * PI_L * q * L3 / N;
    hp.wnr[q] = make_double2((double)cosl(ang), (double)sinl(ang));
  }
  auto up = [&](const std::vector<double2> &v) {
    double2 *d = (double2 *)dalloc(md, sizeof(double2) * v.size());
    QG_CUDA(cudaMemcpy(d, v.data(), sizeof(double2) * v.size(), cudaMemcpyHostToDevice));
    return d;
  };
  hp.s1base = up(s1b); hp.tw2 = up(t2); hp.tw3base = up(t3); hp.wnbase = up(wb);
  int dev = 0, sms = 0;
  QG_CUDA(cudaGetDevice(&dev));
  QG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  hp.fast_grid = 2 * sms;
  hp.fast_attr = 0;
}

// Radix plan: an odd radix first (its stride-R shared-memory scatter is conflict-free while
// Ns = 1), then the largest radices that divide what is left: three passes for every
// benchmark length (2400 = 15*16*10, 1200 = 15*16*5, 480 = 15*16*2, 2304 = 9*16*16).
static bool factorize(int m, int *radix, int &nrad) {
  nrad = 0;
  const int odd[4] = {15, 9, 5, 3};
  for (int i = 0; i < 4; ++i)
    if (m % odd[i] == 0) { radix[nrad++] = odd[i]; m /= odd[i]; break; }
  const int rest[11] = {16, 10, 12, 8, 6, 9, 5, 4, 3, 2, 0};
  while (m > 1 && nrad < 8) {
    int i = 0;
    while (rest[i] && m % rest[i] != 0) ++i;
    if (!rest[i]) return false;
    radix[nrad++] = rest[i];
    m /= rest[i];
  }
  return m == 1;
}

void helm_plan_create(qgcm_model *md, HelmPlan &hp, const Grid &g, int kind, const double *rdm2, int nmodes) {
  hp.kind = kind;
  hp.n = g.nxt;
  if (hp.n % 2 != 0) throw std::runtime_error("helmholtz: nxt must be even");
  hp.m = hp.n / 2;
  if (!factorize(hp.m, hp.radix, hp.nrad))
    throw std::runtime_error("helmholtz: nxt/2 must factor into 2,3,5 (src/parameters_data.F:63-67 recommends the same)");
  hp.nmodes = nmodes;
  hp.ld = g.ld;
  hp.nyp = g.nyp;
  hp.nxp = g.nxp;
  hp.nranks = md->nranks; hp.rank = md->rank;
  hp.row0 = g.own0 + (g.wall_s() ? 1 : 0);
  hp.nrows = (g.own1 - (g.wall_n() ? 1 : 0)) - hp.row0;
  hp.wall_s = g.wall_s(); hp.wall_n = g.wall_n();
  hp.nchunk = (hp.nrows + TRI_L - 1) / TRI_L;
  hp.lastlen = hp.nrows - (hp.nchunk - 1) * TRI_L;
  hp.nk = kind == 0 ? hp.n - 1 : hp.n;
  hp.koff = kind == 0 ? 1 : 0;
  hp.a = g.dxm2;  // dy = dx (src/q-gcm.F:413, :930)
  hp.ftnorm = kind == 0 ? 0.5 / hp.n : 1.0 / hp.n;
  hp.smem_bytes = (size_t)2 * hp.m * sizeof(double2) + 64 * sizeof(double);
  // twiddles in extended precision, rounded once
  std::vector<double2> wm, wn(hp.m + 1);
  std::vector<double> sw(hp.m);
  const long double PI_L = 3.141592653589793238462643383279502884L;
  for (int k = 0; k < hp.m; ++k) sw[k] = (double)(2.0L * sinl(PI_L * k / hp.n));
  // per-pass twiddle tables exp(-2 pi i r k / (Ns R)), laid out [r-1][k]
  {
    int Ns = 1;
    for (int s = 0; s < hp.nrad; ++s) {
      const int R = hp.radix[s];
      hp.twoff[s] = (int)wm.size();
      if (Ns > 1)
        for (int r = 1; r < R; ++r)
          for (int k = 0; k < Ns; ++k) {
            long double a = -2.0L * PI_L * (long double)r * k / ((long double)Ns * R);
            wm.push_back(make_double2((double)cosl(a), (double)sinl(a)));
          }
      Ns *= R;
    }
    if (wm.empty()) wm.push_back(make_double2(1.0, 0.0));
  }
  // constant internal twiddles of the compound butterflies
  {
    std::vector<double2> ct(160, make_double2(1.0, 0.0));
    const int comp[8][2] = {{2, 3}, {3, 3}, {2, 5}, {4, 3}, {3, 5}, {4, 4}, {4, 5}, {5, 5}};
    for (int c = 0; c < 8; ++c) {
      const int RA = comp[c][0], RB = comp[c][1], R = RA * RB;
      for (int k1 = 0; k1 < RA; ++k1)
        for (int n2 = 0; n2 < RB; ++n2) {
          long double a = -2.0L * PI_L * (long double)(n2 * k1) / R;
          ct[ctw_off(R) + k1 * RB + n2] = make_double2((double)cosl(a), (double)sinl(a));
        }
    }
    QG_CUDA(cudaMemcpyToSymbol(c_ctw, ct.data(), sizeof(double2) * 160));
  }
  for (int k = 0; k <= hp.m; ++k) {
    long double a = -2.0L * PI_L * k / hp.n;
    wn[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  hp.wm = (double2 *)dalloc(md, sizeof(double2) * wm.size());
  hp.wn = (double2 *)dalloc(md, sizeof(double2) * (hp.m + 1));
  hp.sintw = (double *)dalloc(md, sizeof(double) * hp.m);
  QG_CUDA(cudaMemcpy(hp.wm, wm.data(), sizeof(double2) * wm.size(), cudaMemcpyHostToDevice));
  QG_CUDA(cudaMemcpy(hp.wn, wn.data(), sizeof(double2) * (hp.m + 1), cudaMemcpyHostToDevice));
  QG_CUDA(cudaMemcpy(hp.sintw, sw.data(), sizeof(double) * hp.m, cudaMemcpyHostToDevice));
  const size_t row = (size_t)hp.ld;
  hp.bcoef = (double *)dalloc(md, sizeof(double) * nmodes * row);
  hp.binv = (double *)dalloc(md, sizeof(double) * nmodes * TRI_L * row);
  hp.vl = (double *)dalloc(md, sizeof(double) * nmodes * TRI_L * row);
  hp.vll = (double *)dalloc(md, sizeof(double) * nmodes * TRI_L * row);
  hp.pt = (double *)dalloc(md, sizeof(double) * nmodes * 2 * hp.nchunk * row);
  hp.fg = (double *)dalloc(md, sizeof(double) * nmodes * 2 * hp.nchunk * row);
  hp.yx = (double *)dalloc(md, sizeof(double) * nmodes * 2 * hp.nchunk * row);
  hp.rowsum = (double *)dalloc(md, sizeof(double) * nmodes * hp.nyp);
  if (kind == 0) {
    // x-sum of a sine series: sum_{i=1}^{n-1} sin(k i pi/n) = cot(k pi / 2n) for odd k, 0 for even
    // k; times 2 for dsint's definition (fft.doc: x(i) = sum_k 2 X(k) sin(k i pi/(n+1)), its n+1 = our n)
    std::vector<double> ws(row, 0.0);
    const long double PI_Q = 3.141592653589793238462643383279502884L;
    for (int k = 1; k < hp.n; k += 2) ws[k] = (double)(2.0L * cosl(PI_Q * k / (2.0L * hp.n)) / sinl(PI_Q * k / (2.0L * hp.n)));
    hp.wsum = (double *)dalloc(md, sizeof(double) * row);
    QG_CUDA(cudaMemcpy(hp.wsum, ws.data(), sizeof(double) * row, cudaMemcpyHostToDevice));
    hp.nspec = hp.nchunk * ((hp.koff + hp.nk + 2 * TRI3_TP - 1) / (2 * TRI3_TP));
    QG_CUDA(cudaFuncSetAttribute(k_tri3<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri3_smem));
    QG_CUDA(cudaFuncSetAttribute(k_tri3<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri3_smem));
    // (no carve-out preference: forcing the largest shared-memory carve-out shrinks L1, which these
    // kernels and k_dst3 use for their L2-resident tables -- measured 10-25 % slower)
    hp.spec = (double *)dalloc(md, sizeof(double) * nmodes * hp.nspec);
  }
  for (int r = 0; r < 16; ++r) hp.slab_rows[r] = 0;
  if (hp.nranks > 1) {
    if (hp.nranks > 8) throw std::runtime_error("helmholtz: at most 8 y-slabs");
    for (int r = 0; r < hp.nranks; ++r) {
      int p0, p1;
      slab_bounds(g.nyp_g, hp.nranks, r, &p0, &p1);
      hp.slab_rows[r] = (p1 - (p1 == g.nyp_g ? 1 : 0)) - (p0 + (p0 == 0 ? 1 : 0));
    }
    hp.slab_ae = (double *)dalloc(md, sizeof(double) * nmodes * hp.nranks * 2 * row);
    hp.slab_fg = (double *)dalloc(md, sizeof(double) * hp.nranks * nmodes * 2 * row);
    hp.slab_send = (double *)dalloc(md, sizeof(double) * nmodes * 2 * row);
    hp.slab_yx = (double *)dalloc(md, sizeof(double) * nmodes * 2 * row);
  }
  
  {
    // the attribute belongs to the kernel, not to the plan: only ever raise it (an ocean and
    // an atmosphere plan of different lengths coexist in coupled models)
    static size_t xform_smem_cap = 48 * 1024;
    if (hp.smem_bytes > xform_smem_cap) {
      QG_CUDA(cudaFuncSetAttribute(k_xform, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp.smem_bytes));
      xform_smem_cap = hp.smem_bytes;
    }
  }
  dst3_plan(md, hp);
  // diagonal b(i) = bd2(i) - rdm2(m), src/q-gcm.F:929-973 and src/ocisubs.F:148-150
  const double PI = 3.14159265358979324, TWOPI = 6.28318530717958648;
  std::vector<double> bd2(hp.n, 0.0), b((size_t)nmodes * hp.n);
  const double a = hp.a, dxm2 = g.dxm2;
  const int nxt = hp.n;
  if (kind == 1) {
    for (int i = 2; i <= nxt / 2; ++i) {
      int i1 = 2 * i - 1;
      bd2[i1 - 2] = -2.0 * a + 2.0 * dxm2 * (cos((i - 1) * TWOPI / nxt) - 1.0);
      bd2[i1 - 1] = bd2[i1 - 2];
    }
    bd2[0] = -2.0 * a;
    bd2[nxt - 1] = -2.0 * a - 4.0 * dxm2;
  } else {
    for (int i = 2; i <= nxt; ++i) bd2[i - 2] = -2.0 * a + 2.0 * dxm2 * (cos((i - 1) * PI / nxt) - 1.0);
    bd2[nxt - 1] = 0.0;
  }
  for (int mo = 0; mo < nmodes; ++mo)
    for (int i = 0; i < nxt; ++i) b[(size_t)mo * nxt + i] = bd2[i] - rdm2[mo];
  helm_set_diag(md, hp, b.data());
}

static TriArgs tri_args(HelmPlan &hp, double *wrk, size_t lsz, int nmodes) {
  TriArgs t;
  t.ld = hp.ld; t.nyp = hp.nyp; t.nk = hp.nk; t.koff = hp.koff; t.nchunk = hp.nchunk;
  t.lastlen = hp.lastlen; t.nmodes = nmodes; t.row0 = hp.row0;
  t.use_yx = (hp.nchunk > 1 || hp.nranks > 1) ? 1 : 0; t.nranks = hp.nranks;
  t.peer.n = 0; t.ticket = nullptr;
  t.slab_phase = 0; t.slab_send = hp.slab_send; t.slab_outer = hp.slab_yx; t.lsz = lsz; t.a = hp.a; t.ftnorm = hp.ftnorm; t.wrk = wrk;
  t.bcoef = hp.bcoef; t.binv = hp.binv; t.vl = hp.vl; t.vll = hp.vll; t.pt = hp.pt; t.fg = hp.fg; t.yx = hp.yx;
  return t;
}

static SlabArgs slab_args(HelmPlan &hp, int nmodes);

// b_host: [nmodes][n] in the reference's ordering: box b(i-1) multiplies wavenumber column
// i (src/ocisubs.F:472), periodic b(i) column i (src/ocisubs.F:577)
void helm_set_diag(qgcm_model *md, HelmPlan &hp, const double *b_host) {
  std::vector<double> tmp((size_t)hp.nmodes * hp.ld, 1.0);
  for (int mo = 0; mo < hp.nmodes; ++mo)
    for (int s = 0; s < hp.nk; ++s) tmp[(size_t)mo * hp.ld + hp.koff + s] = b_host[(size_t)mo * hp.n + s];
  QG_CUDA(cudaMemcpyAsync(hp.bcoef, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice, md->stream));
  QG_CUDA(cudaStreamSynchronize(md->stream));
  TriArgs t = tri_args(hp, nullptr, 0, hp.nmodes);
  dim3 grid((hp.nk + 127) / 128, hp.nmodes);
  QG_LAUNCH(md, "k_tri_tables", grid, 128, 0, k_tri_tables, t);
  if (hp.nranks > 1) {
    SlabArgs sa = slab_args(hp, hp.nmodes);
    QG_LAUNCH(md, "k_slab_spikes", grid, 128, 0, k_slab_spikes, sa);
  }
  QG_CUDA(cudaGetLastError());
}

static SlabArgs slab_args(HelmPlan &hp, int nmodes) {
  SlabArgs t;
  t.ld = hp.ld; t.nk = hp.nk; t.koff = hp.koff; t.nchunk = hp.nchunk; t.lastlen = hp.lastlen; t.nmodes = nmodes;
  t.nranks = hp.nranks; t.rank = hp.rank; t.a = hp.a;
  for (int r = 0; r < 16; ++r) t.nrows_of[r] = hp.slab_rows[r];
  t.bcoef = hp.bcoef; t.vl = hp.vl; t.vll = hp.vll; t.fg = hp.fg; t.yx = hp.yx;
  t.ae = hp.slab_ae; t.send = hp.slab_send; t.all = hp.slab_fg; t.outer = hp.slab_yx;
  t.peer = hp.slab_peer; t.peer_err = hp.slab_err;
  if (t.peer.n) t.all = t.peer.box[t.peer.rank] + peer_off_fg(t.peer.n, t.peer.fglen, (int)(t.peer.epoch & 1ull), 0);
  return t;
}

static XfArgs xf_args(HelmPlan &hp, double *wrk, size_t lsz) {
  XfArgs x;
  x.f.n = hp.n; x.f.m = hp.m; x.f.nrad = hp.nrad;
  for (int i = 0; i < 8; ++i) { x.f.radix[i] = hp.radix[i]; x.f.twoff[i] = hp.twoff[i]; }
  x.f.tw = hp.wm; x.f.wn = hp.wn; x.f.sintw = hp.sintw;
  x.kind = hp.kind; x.inverse = 0; x.ld = hp.ld; x.nyp = hp.nyp; x.nxp = hp.nxp; x.row0 = hp.row0; x.lsz = lsz;
  x.wrk = wrk; x.rowsum = hp.rowsum;
  return x;
}

// first half: forward transform, chunk-local solves, chunk interface system; with slabs also
// the first/last rows of the slab-local solution (hp.slab_send, to be all-gathered)
static Mix3 mix3_args(const HelmPlan &hp, const FusedInv &fz) {
  Mix3 mx;
  mx.f0 = fz.f0;
  for (int i = 0; i < NLMAX * NLMAX; ++i) { mx.ctl2m[i] = fz.ctl2m[i]; mx.ctm2l[i] = fz.ctm2l[i]; }
  mx.wsum = hp.wsum; mx.spec = hp.spec;
  return mx;
}

bool helm_can_fuse(const qgcm_model *m, const HelmPlan &hp, int nl) {
  static const bool off = env_int("QGCM_NOFUSE", 0) != 0;      // A/B switch (scripts/ab_env.sh)
  return !off && hp.kind == 0 && hp.fast && nl == 3 && hp.nmodes == 3 && hp.spec;
}

// after the constraint algebra: spectral layer rows -> pressure layers (+ homogeneous solutions)
void helm_fused_inverse(qgcm_model *md, HelmPlan &hp, double *wrk, int nl, const FusedInv &fz) {
  const size_t lsz = (size_t)hp.ld * hp.nyp;
  dst3_launch(md, hp, wrk, lsz, nl, DST_FUSED_I, &fz);      // the blocks also write the wall rows, a slice of columns each
  QG_CUDA(cudaGetLastError());
}

void helm_solve_a(qgcm_model *md, HelmPlan &hp, double *wrk, int nmodes, const FusedInv *fz) {
  const size_t lsz = (size_t)hp.ld * hp.nyp;
  XfArgs x = xf_args(hp, wrk, lsz);
  dim3 gx(hp.nrows, nmodes);
  if (fz)
    dst3_launch(md, hp, wrk, lsz, nmodes, DST_FUSED_F, fz);
  else if (hp.fast)
    dst3_launch(md, hp, wrk, lsz, nmodes, DST_PLAIN_F);
  else
    QG_LAUNCH(md, "k_xform", gx, 256, hp.smem_bytes, k_xform, x);
  TriArgs t = tri_args(hp, wrk, lsz, nmodes);
  dim3 gl((hp.nk + 127) / 128, hp.nchunk, nmodes), gr((hp.nk + 127) / 128, nmodes);
  if (hp.nchunk > 1 || hp.nranks > 1) {
    if (fz) {
      const int ntx = (hp.koff + hp.nk + 2 * TRI3_TP - 1) / (2 * TRI3_TP);      // tiles start at column 0
      QG_LAUNCH(md, "k_tri_fg", std::min(ntx * hp.nchunk, hp.fast_grid / 2 * TRI3_BLK), 3 * TRI3_TP, tri3_smem, (k_tri3<3, false>), t, mix3_args(hp, *fz), ntx, env_int("QGCM_TRI_ORDER", 1));
    } else {
      auto kfg = k_tri_local<false>;
      QG_LAUNCH(md, "k_tri_fg", gl, 128, 0, kfg, t);
    }
  }
  t.slab_phase = hp.nranks > 1 ? 1 : 0;
  hp.slab_pushed = false;
  if (hp.nranks > 1 && hp.nchunk > 1 && peer_active(md) && !env_int("QGCM_PEER_NOFUSE", 0)) {
    // peer-memory transport: the interface kernel itself delivers the slab's first/last rows
    t.peer = md->peer;
    t.peer.epoch = ++md->epoch_fg;
    t.ticket = md->d_ticket2;
    hp.slab_peer = t.peer;
    hp.slab_err = md->d_peer_err;
    hp.slab_pushed = true;
  }
  if (hp.nchunk > 1) QG_LAUNCH(md, "k_tri_reduced", gr, 128, 0, k_tri_reduced, t, slab_args(hp, nmodes));
  if (hp.nranks > 1 && hp.nchunk == 1) {   // a one-chunk slab has no interface system to piggyback on
    SlabArgs sa = slab_args(hp, nmodes);
    QG_LAUNCH(md, "k_slab_fg", gr, 128, 0, k_slab_fg, sa);
  }
}

// second half: (slabs: inter-slab system from the gathered rows, interface system again with
// the neighbour rows) final chunk solves, inverse transform, wall rows
void helm_solve_b(qgcm_model *md, HelmPlan &hp, double *wrk, int nmodes, const FusedInv *fz) {
  const size_t lsz = (size_t)hp.ld * hp.nyp;
  TriArgs t = tri_args(hp, wrk, lsz, nmodes);
  dim3 gl((hp.nk + 127) / 128, hp.nchunk, nmodes), gr((hp.nk + 127) / 128, nmodes);
  if (hp.nranks > 1) {
    SlabArgs sa = slab_args(hp, nmodes);
    t.slab_phase = 2;
    if (hp.nchunk > 1) {
      QG_LAUNCH(md, "k_tri_reduced", gr, 128, 0, k_tri_reduced, t, sa);      // includes the inter-slab solve
    } else {
      QG_LAUNCH(md, "k_slab_solve", gr, 128, 0, k_slab_solve, sa);
      QG_LAUNCH(md, "k_slab_outer", gr, 128, 0, k_slab_outer, sa);
    }
  }
  if (fz) {
    // fused: the final chunk solves project back onto the layers and leave the modal integrals;
    // the inverse transform follows the constraint algebra (helm_fused_inverse)
    const int ntx = (hp.koff + hp.nk + 2 * TRI3_TP - 1) / (2 * TRI3_TP);
    QG_LAUNCH(md, "k_tri_local", std::min(ntx * hp.nchunk, hp.fast_grid / 2 * TRI3_BLK), 3 * TRI3_TP, tri3_smem, (k_tri3<3, true>), t, mix3_args(hp, *fz), ntx, env_int("QGCM_TRI_ORDER", 1));
    QG_CUDA(cudaGetLastError());
    return;
  }
  auto kfin = k_tri_local<true>;
  QG_LAUNCH(md, "k_tri_local", gl, 128, 0, kfin, t);
  if (hp.fast) {
    dst3_launch(md, hp, wrk, lsz, nmodes, DST_PLAIN_I);
  } else {
    XfArgs x = xf_args(hp, wrk, lsz);
    dim3 gx(hp.nrows, nmodes);
    x.inverse = 1;
    QG_LAUNCH(md, "k_xform_inv", gx, 256, hp.smem_bytes, k_xform, x);
  }
  // the wall rows of the work array stay zero from one ocean/atmosphere step to the next (the
  // right-hand side kernel writes interior rows only); only homsol and qgcm_helmholtz fill them
  if (hp.walls_dirty) {
    QG_LAUNCH(md, "k_zero_rows", (hp.nxp + 255) / 256, 256, 0, k_zero_rows, wrk, lsz, hp.ld, hp.nyp, hp.nxp, nmodes, hp.wall_s,
              hp.wall_n);
    hp.walls_dirty = false;
  }
  QG_CUDA(cudaGetLastError());
}

// zero the wall rows of every mode of a work array whose rows a caller has filled (homsol):
// the time loop relies on them staying zero
void helm_clean_walls(qgcm_model *md, HelmPlan &hp, double *wrk, int nmodes) {
  const size_t lsz = (size_t)hp.ld * hp.nyp;
  QG_LAUNCH(md, "k_zero_rows", (hp.nxp + 255) / 256, 256, 0, k_zero_rows, wrk, lsz, hp.ld, hp.nyp, hp.nxp, nmodes, hp.wall_s, hp.wall_n);
  hp.walls_dirty = false;
}

// in place on wrk[nmodes][nyp][ld]: rhs -> solution with zero boundary values (one GPU; the
// slab drivers in slab.cu call the two halves around an all-gather)
void helm_solve(qgcm_model *md, HelmPlan &hp, double *wrk, int nmodes) {
  if (hp.nranks > 1) throw std::runtime_error("helm_solve: a y-slab model must be driven through the slab procedures");
  helm_solve_a(md, hp, wrk, nmodes);
  helm_solve_b(md, hp, wrk, nmodes);
}

}  // namespace qg

// TEST INFRASTRUCTURE ONLY -- see orc_fft.h.
#include "orc_fft.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace orc {

typedef std::complex<double> cd;

static const long double PI_L = 3.141592653589793238462643383279502884L;

void FftPlan::init(int n_real) {
  if (n_real < 2 || (n_real & 1)) throw std::runtime_error("orc::FftPlan: n must be even");
  n = n_real;
  m = n / 2;
  radices.clear();
  int rem = m;
  // factor preference 4,2,3,5 as drfti1.f:16, then any remaining primes
  const int pref[4] = {4, 2, 3, 5};
  for (int t = 0; t < 4; ++t)
    while (rem % pref[t] == 0) { radices.push_back(pref[t]); rem /= pref[t]; }
  for (int f = 7; rem > 1; f += 2)
    while (rem % f == 0) { radices.push_back(f); rem /= f; }
  wm.resize(m);
  for (int k = 0; k < m; ++k) {
    long double a = -2.0L * PI_L * (long double)k / (long double)m;
    wm[k] = cd((double)cosl(a), (double)sinl(a));
  }
  wn.resize(m + 1);
  for (int k = 0; k <= m; ++k) {
    long double a = -2.0L * PI_L * (long double)k / (long double)n;
    wn[k] = cd((double)cosl(a), (double)sinl(a));
  }
  sint_w.assign(m, 0.0);
  for (int k = 1; k < m; ++k)
    sint_w[k] = (double)(2.0L * sinl(PI_L * (long double)k / (long double)n));
}

static inline void dft2(cd *v) {
  cd a = v[0], b = v[1];
  v[0] = a + b; v[1] = a - b;
}
static inline cd mul_mi(cd a) { return cd(a.imag(), -a.real()); }  // a * (-i)
static inline void dft4(cd *v) {
  cd a = v[0] + v[2], b = v[0] - v[2], c = v[1] + v[3], d = mul_mi(v[1] - v[3]);
  v[0] = a + c; v[1] = b + d; v[2] = a - c; v[3] = b - d;
}
static inline void dft3(cd *v) {
  const double s = 0.86602540378443864676;  // sin(pi/3)
  cd t1 = v[1] + v[2];
  cd t2 = v[0] - 0.5 * t1;
  cd t3 = s * mul_mi(v[1] - v[2]);
  v[0] = v[0] + t1; v[1] = t2 + t3; v[2] = t2 - t3;
}
static inline void dft5(cd *v) {
  const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
  const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
  cd a1 = v[1] + v[4], a2 = v[2] + v[3];
  cd b1 = mul_mi(v[1] - v[4]), b2 = mul_mi(v[2] - v[3]);
  cd r1 = v[0] + c1 * a1 + c2 * a2;
  cd r2 = v[0] + c2 * a1 + c1 * a2;
  cd i1 = s1 * b1 + s2 * b2;
  cd i2 = s2 * b1 - s1 * b2;
  v[0] = v[0] + a1 + a2;
  v[1] = r1 + i1; v[4] = r1 - i1;
  v[2] = r2 + i2; v[3] = r2 - i2;
}
static void dftg(cd *v, int R, const cd *wm, int m) {
  std::vector<cd> o(R);
  for (int q = 0; q < R; ++q) {
    cd acc = 0.0;
    for (int s = 0; s < R; ++s) acc += v[s] * wm[(long long)(q * s % R) * (m / R)];
    o[q] = acc;
  }
  for (int q = 0; q < R; ++q) v[q] = o[q];
}

// forward complex FFT of length p.m (Stockham autosort), in x, scratch y
static void cfft(const FftPlan &p, cd *x, cd *y) {
  const int M = p.m;
  int Ns = 1;
  cd *in = x, *out = y;
  cd v[64];
  for (size_t s = 0; s < p.radices.size(); ++s) {
    const int R = p.radices[s];
    const int L = M / R;
    const int tws = M / (Ns * R);
    for (int j = 0; j < L; ++j) {
      const int k = j % Ns;
      const int j0 = (j - k) * R + k;
      const int tw = k * tws;
      v[0] = in[j];
      for (int r = 1; r < R; ++r) v[r] = in[j + r * L] * p.wm[(int)(((long long)r * tw) % M)];
      switch (R) {
        case 2: dft2(v); break;
        case 3: dft3(v); break;
        case 4: dft4(v); break;
        case 5: dft5(v); break;
        default: dftg(v, R, p.wm.data(), M); break;
      }
      for (int r = 0; r < R; ++r) out[j0 + r * Ns] = v[r];
    }
    cd *t = in; in = out; out = t;
    Ns *= R;
  }
  if (in != x) std::memcpy(x, in, sizeof(cd) * M);
}

void rfftf(const FftPlan &p, double *r, double *scratch) {
  const int N = p.n, M = p.m;
  cd *z = reinterpret_cast<cd *>(scratch);
  cd *y = z + M;  // scratch holds 2*n doubles = 2*M complex
  for (int k = 0; k < M; ++k) z[k] = cd(r[2 * k], r[2 * k + 1]);
  cfft(p, z, y);
  const cd z0 = z[0];
  r[0] = z0.real() + z0.imag();
  r[N - 1] = z0.real() - z0.imag();
  for (int k = 1; k < M; ++k) {
    cd a = z[k], b = std::conj(z[M - k]);
    cd e = 0.5 * (a + b);
    cd o = 0.5 * mul_mi(a - b) * p.wn[k];
    cd X = e + o;
    r[2 * k - 1] = X.real();
    r[2 * k] = X.imag();
  }
}

void rfftb(const FftPlan &p, double *r, double *scratch) {
  const int N = p.n, M = p.m;
  cd *z = reinterpret_cast<cd *>(scratch);
  cd *y = z + M;
  // Z_k = (X_k + conj X_{M-k}) + i e^{+2 pi i k/N} (X_k - conj X_{M-k}); feed conj(Z) to
  // the forward transform and conjugate the result (unnormalised inverse).
  auto X = [&](int k) -> cd {
    if (k == 0) return cd(r[0], 0.0);
    if (k == M) return cd(r[N - 1], 0.0);
    return cd(r[2 * k - 1], r[2 * k]);
  };
  for (int k = 0; k < M; ++k) {
    cd a = X(k), b = std::conj(X(M - k));
    cd e = a + b;
    cd o = cd(0.0, 1.0) * std::conj(p.wn[k]) * (a - b);
    z[k] = std::conj(e + o);
  }
  cfft(p, z, y);
  for (int k = 0; k < M; ++k) {
    r[2 * k] = z[k].real();
    r[2 * k + 1] = -z[k].imag();
  }
}

// dsint.f:17-43, 0-based restatement.  x[0..n-2] data (n-1 points), x[n-1] scratch.
void dsint(const FftPlan &p, double *x, double *scratch) {
  const int N = p.n;            // np1 in dsint.f
  const int nn = N - 1;         // n in dsint.f (odd, since N even)
  const int ns2 = nn / 2;
  double *work = scratch;       // 2*N doubles for rfftf
  // build t_0..t_{N-1} in place: x holds x_1..x_{N-1} at indices 0..N-2
  double x1 = x[0];
  // shift: t_k lives at index k; input x_k lives at index k-1.  Work right-to-left safe
  // order exactly as the Fortran: it reads x(np1-k) and x(k+1) before overwriting them.
  x[0] = 0.0;
  for (int k = 1; k <= ns2; ++k) {
    double xkc = x[N - k - 1];            // x(np1-k) 1-based
    double t1 = x1 - xkc;
    double t2 = p.sint_w[k] * (x1 + xkc);
    x1 = x[k];                            // x(k+1)
    x[k] = t1 + t2;
    x[N - k] = t2 - t1;                   // x(np1+1-k)
  }
  if (nn % 2 != 0) x[ns2 + 1] = 4.0 * x1;  // x(ns2+2)
  rfftf(p, x, work);
  x[0] = 0.5 * x[0];
  for (int i = 3; i <= nn; i += 2) {       // 1-based i
    double xim1 = x[i - 2];
    x[i - 2] = -x[i - 1];
    x[i - 1] = x[i - 3] + xim1;
  }
  if (nn % 2 == 0) x[nn - 1] = -x[nn];
}

}  // namespace orc

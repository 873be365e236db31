// TEST INFRASTRUCTURE ONLY -- CPU oracle for the Q-GCM hot path; never linked into
// or imported by the product (q-gcm_b200/).  See oracle/README.md.
//
// Real FFT and DST-I with the *definitions* of FFTPACK's drfftf/drfftb/dsint
// (reference: src/fftpack/newbihar/fft.doc:96-114 packed ordering, :334-342 sine
// transform; algorithm of dsint follows src/fftpack/newbihar/dsint.f:17-43 and
// dsinti.f:19-26).  The butterflies themselves are an independent Stockham
// mixed-radix implementation, not a restatement of dradf*/dradb* -- the transform
// definition is the contract, FFTPACK's rounding is not.
#pragma once
#include <complex>
#include <vector>

namespace orc {

struct FftPlan {
  int n = 0;                         // real length (even)
  int m = 0;                         // complex length n/2
  std::vector<int> radices;          // product = m
  std::vector<std::complex<double>> wm;   // exp(-2 pi i k / m), k < m
  std::vector<std::complex<double>> wn;   // exp(-2 pi i k / n), k <= m
  std::vector<double> sint_w;        // 2 sin(k pi / n), k = 1..n/2-1  (dsinti.f:23-26)
  void init(int n_real);
};

// r(1..n) in FFTPACK packed order, in place; scratch must hold 2*n doubles
void rfftf(const FftPlan &p, double *r, double *scratch);
void rfftb(const FftPlan &p, double *r, double *scratch);
// dsint(n-1, x, ..): x(1..n-1) data, x(n) scratch element (as in ocisubs.F:458-459)
void dsint(const FftPlan &p, double *x, double *scratch);

}  // namespace orc

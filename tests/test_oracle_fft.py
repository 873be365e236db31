"""Pin the oracle's transforms against an independent implementation (scipy/pocketfft)
using the FFTPACK definitions: packed real FFT (fft.doc:96-114) and DST-I (fft.doc:334-342)."""
import numpy as np
import pytest
import scipy.fft as sf

LENGTHS = [8, 12, 20, 30, 96, 288, 384, 480, 768, 960, 2400, 4608, 4800, 14, 22, 26]


def pack(X, n):
    """numpy rfft output -> FFTPACK packed order"""
    r = np.empty(n)
    r[0] = X[0].real
    r[1:n - 1:2] = X[1:n // 2].real
    r[2:n - 1:2] = X[1:n // 2].imag
    r[n - 1] = X[n // 2].real
    return r


@pytest.mark.parametrize("n", LENGTHS)
def test_rfftf_matches_scipy(pyorc, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    got = pyorc.rfftf(x)
    want = pack(sf.rfft(x), n)
    assert np.linalg.norm(got - want) <= 1e-14 * np.sqrt(n) * np.linalg.norm(want)


@pytest.mark.parametrize("n", LENGTHS)
def test_rfftb_inverts(pyorc, n):
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal(n)
    back = pyorc.rfftb(pyorc.rfftf(x))
    assert np.linalg.norm(back - n * x) <= 1e-14 * n * np.sqrt(n) * np.linalg.norm(x)


@pytest.mark.parametrize("n", LENGTHS)
def test_dsint_matches_scipy(pyorc, n):
    rng = np.random.default_rng(n + 2)
    x = rng.standard_normal(n - 1)
    got = pyorc.dsint(x)
    want = sf.dst(x, type=1)  # 2*sum x_i sin(pi (i+1)(k+1)/n): FFTPACK's unnormalised DST-I
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)
    # self-inverse up to 2(n+1) with n+1 -> n here (fft.doc:342)
    again = pyorc.dsint(got)
    assert np.linalg.norm(again - 2.0 * n * x) <= 1e-13 * 2 * n * np.linalg.norm(x)

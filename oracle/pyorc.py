"""TEST INFRASTRUCTURE ONLY -- Python handle on the CPU oracle (oracle/liborc.so).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Reuses the product's generic ctypes binding (CModel) with the
``orc_`` symbol prefix; the product never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _pkg  # noqa: E402

qg = _pkg.load()
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liborc.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
    return _LIB


class Oracle(qg.CModel):
    def __init__(self, cfg):
        super().__init__(lib(), "orc_", cfg)


def subsample(field, nsk):
    """the wrk vector of ocnc_out / atnc_out (src/nc_subs.F:869-890): wrk(i + iw*(j-1) + iw*jw*(k-1)) =
    f(1+(i-1)*nsk, 1+(j-1)*nsk, k) with iw = min(mod(nx,nsk),1) + (nx-mod(nx,nsk))/nsk; `field` is the
    Fortran-shaped array (nx, ny[, nl])"""
    f = np.asarray(field)
    if f.ndim == 2:
        f = f[:, :, None]
    nx, ny, nl = f.shape
    iw = min(nx % nsk, 1) + (nx - nx % nsk) // nsk
    jw = min(ny % nsk, 1) + (ny - ny % nsk) // nsk
    wrk = np.empty(iw * jw * nl)
    for k in range(nl):
        for j in range(jw):
            for i in range(iw):
                wrk[i + iw * j + iw * jw * k] = f[i * nsk, j * nsk, k]
    return wrk


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def rfftf(x):
    a = np.array(x, dtype=np.float64)
    assert lib().orc_rfftf(C.c_int(a.size), _dp(a)) == 0
    return a


def rfftb(x):
    a = np.array(x, dtype=np.float64)
    assert lib().orc_rfftb(C.c_int(a.size), _dp(a)) == 0
    return a


def dsint(x):
    """DST-I of x (n-1 points) with FFTPACK's unnormalised definition"""
    n = len(x) + 1
    a = np.zeros(n, dtype=np.float64)
    a[: n - 1] = x
    assert lib().orc_dsint(C.c_int(n), _dp(a)) == 0
    return a[: n - 1].copy()


def xintp(v):
    a = np.asfortranarray(v, dtype=np.float64)
    f = lib().orc_xintp
    f.restype = C.c_double
    return f(_dp(a), C.c_int(a.shape[0]), C.c_int(a.shape[1]))

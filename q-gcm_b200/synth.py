"""Seedable synthetic initial states and forcing.

The reference's own forcing files are missing from the repo (.MISSING_LARGE_BLOBS) and
its generators need netCDF-Fortran, so both the oracle and the CUDA library are fed the
same synthetic fields built here:

* ``eddy``   -- the fork's own initial state: a Gaussian SSH eddy in layer 1, zero
  forcing (k247_make_restart_q-gcm.F90:85,96-100,241-262; k247_make_forcing_q-gcm.F90:126-136).
* ``random`` -- band-limited random p (|k|,|l| <= 8), pom = po*(1-1e-3), smooth random
  SST about sstbar, double-gyre wind stress and a sinusoidal heat flux.

q is *not* produced here: call ``qcomp_ocean`` / ``qcomp_atmos`` on the model, as the
reference does at start-up (src/q-gcm.F:719-749).
"""
import math

import numpy as np

SEED = 20261018


def _smooth5(a, periodic_x):
    """two passes of a 5-point filter"""
    for _ in range(2):
        b = a.copy()
        up = np.roll(a, -1, axis=1); up[:, -1] = a[:, -1]
        dn = np.roll(a, 1, axis=1); dn[:, 0] = a[:, 0]
        ri = np.roll(a, -1, axis=0)
        le = np.roll(a, 1, axis=0)
        if not periodic_x:
            ri[-1, :] = a[-1, :]
            le[0, :] = a[0, :]
        a = 0.5 * b + 0.125 * (up + dn + ri + le)
    return a


def _bandlimited(rng, nx, ny, amp, periodic_x, nmode=8):
    """sum_{k,l<=nmode} a_kl X_k(x) sin(l pi y/Ly) on an (nx,ny) p grid, zero on N/S walls;
    X_k = sin(k pi x/Lx) (box: zero on W/E walls) or cos/sin(2 pi k x/Lx) (cyclic, with
    column nx equal to column 1)."""
    x = np.arange(nx, dtype=np.float64) / (nx - 1)
    y = np.arange(ny, dtype=np.float64) / (ny - 1)
    k = np.arange(1, nmode + 1, dtype=np.float64)
    sy = np.sin(math.pi * np.outer(k, y))
    sy[:, 0] = 0.0
    sy[:, -1] = 0.0
    if periodic_x:
        cx = np.cos(2.0 * math.pi * np.outer(k, x))
        sx = np.sin(2.0 * math.pi * np.outer(k, x))
        cx[:, -1] = cx[:, 0]
        sx[:, -1] = sx[:, 0]
        a = rng.standard_normal((nmode, nmode)) * amp
        b = rng.standard_normal((nmode, nmode)) * amp
        f = cx.T @ a @ sy + sx.T @ b @ sy
    else:
        sx = np.sin(math.pi * np.outer(k, x))
        sx[:, 0] = 0.0
        sx[:, -1] = 0.0
        a = rng.standard_normal((nmode, nmode)) * amp
        f = sx.T @ a @ sy
    return f / nmode


def ocean_state(p, cfg, kind="random", seed=SEED, amp=1.0):
    """dict of Fortran-shaped arrays for every ocean input field except q.  ``amp`` scales
    the pressure amplitude (reduced-size test grids keep the full-size Courant number by
    scaling p with the domain length)"""
    rng = np.random.default_rng(seed)
    nxp, nyp, nxt, nyt, nl = p.nxpo, p.nypo, p.nxto, p.nyto, p.nlo
    cyc = p.has("cyclic_ocean")
    rad = cfg._radiation
    st = {}
    po = np.zeros((nxp, nyp, nl))
    if kind == "eddy":
        g, amp, L = 9.8, 0.15, 80.0e3
        xx = (np.arange(nxp) - nxt // 2) * p.dxo
        yy = (np.arange(nyp) - nyt // 2) * p.dxo
        r2 = xx[:, None] ** 2 + yy[None, :] ** 2
        po[:, :, 0] = g * amp * np.exp(-r2 / L ** 2)
        if not cyc:
            po[0, :, 0] = po[-1, :, 0] = 0.0
        po[:, 0, 0] = po[:, -1, 0] = 0.0
        st["po"] = po
        st["pom"] = po.copy()
        st["sst"] = np.zeros((nxt, nyt))
        st["sstm"] = np.zeros((nxt, nyt))
        st["tauxo"] = np.zeros((nxp, nyp))
        st["tauyo"] = np.zeros((nxp, nyp))
        st["fnetoc"] = np.zeros((nxt, nyt))
    else:
        amps = [2.0, 1.0, 0.5] + [0.25] * max(0, nl - 3)
        for k in range(nl):
            po[:, :, k] = _bandlimited(rng, nxp, nyp, amp * amps[k], cyc)
        st["po"] = po
        st["pom"] = po * (1.0 - 1.0e-3)
        sstbar = np.asarray(rad["sstbar"])
        noise = _smooth5(0.5 * rng.standard_normal((nxt, nyt)), cyc)
        st["sst"] = sstbar[None, :] + noise
        st["sstm"] = sstbar[None, :] + 0.999 * noise
        ylo = p.nyto * p.dxo
        yp = np.arange(nyp) * p.dxo
        tau0 = 1.0e-4
        st["tauxo"] = np.repeat((-tau0 * np.cos(2.0 * math.pi * yp / ylo))[None, :], nxp, axis=0)
        st["tauyo"] = np.zeros((nxp, nyp))
        yt = (np.arange(nyt) + 0.5) * p.dxo
        st["fnetoc"] = np.repeat((-40.0 * np.sin(math.pi * (yt / ylo - 0.5)))[None, :], nxt, axis=0)
    st["ddynoc"] = np.zeros((nxp, nyp))
    st["sstbar"] = np.asarray(rad["sstbar"], dtype=np.float64)
    return st


def atmos_state(p, cfg, kind="random", seed=SEED + 1):
    """dict of Fortran-shaped arrays for every atmosphere input field except q"""
    rng = np.random.default_rng(seed)
    nxp, nyp, nxt, nyt, nl = p.nxpa, p.nypa, p.nxta, p.nyta, p.nla
    rad = cfg._radiation
    st = {}
    pa = np.zeros((nxp, nyp, nl))
    # keep the geostrophic wind of the full-size deck (30720 km channel) on reduced grids
    scale = min(1.0, (p.nxta * p.ndxr * p.dxo) / 3.072e7)
    amps = [a * scale for a in [2.0e3, 1.0e3, 0.5e3] + [0.25e3] * max(0, nl - 3)]
    if kind != "eddy":
        for k in range(nl):
            pa[:, :, k] = _bandlimited(rng, nxp, nyp, amps[k], True)
    st["pa"] = pa
    st["pam"] = pa * (1.0 - 1.0e-3)
    astbar = np.asarray(rad["astbar"])
    noise = _smooth5(rng.standard_normal((nxt, nyt)), True) if kind != "eddy" else np.zeros((nxt, nyt))
    st["ast"] = astbar[None, :] + noise
    st["astm"] = astbar[None, :] + 0.999 * noise
    hn = _smooth5(rng.standard_normal((nxt, nyt)), True) if kind != "eddy" else np.zeros((nxt, nyt))
    st["hmixa"] = p.hmat * (1.0 + 0.05 * hn)
    st["hmixam"] = p.hmat * (1.0 + 0.0499 * hn)
    st["ddynat"] = np.zeros((nxp, nyp))
    st["dtopat"] = np.zeros((nxp, nyp))
    st["xc1ast"] = np.zeros((nxt, nyt))
    st["astbar"] = np.asarray(rad["astbar"], dtype=np.float64)
    return st


def init_model(m, p, cfg, kind="random", seed=SEED, amp=None):
    """the start-up sequence of src/q-gcm.F:597-976 on a model (oracle or CUDA):
    load state, constr, q from p, first xforc, zero entrainment, homsol."""
    if amp is None:
        # keep the advective Courant number of the full-size decks (4800 km basins)
        amp = min(1.0, (p.nxto * p.dxo) / 4.8e6 * 4.0)
    if not p.has("atmos_only"):
        st = ocean_state(p, cfg, kind, seed, amp)
        for k, v in st.items():
            m.set_field(k, v)
    if not p.has("ocean_only"):
        st = atmos_state(p, cfg, kind, seed + 1)
        for k, v in st.items():
            m.set_field(k, v)
    m.constr()
    if not p.has("atmos_only"):
        m.qcomp_ocean()
    if not p.has("ocean_only"):
        m.qcomp_atmos()
    m.xforc()
    m.homsol()

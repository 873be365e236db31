"""Host-side restatement of ``radiat`` (src/radsubs.f:44-592) and ``trapin`` (:596-634).

Produces the scalar inputs the step kernels take from the radiation module
(src/radiate_data.F:34-38): relative layer temperatures ``toc``/``tat``, the A-D
linearised radiation coefficients, the entrainment coefficients ``aface..dface``, the
radiative-equilibrium profiles ``sstbar``/``astbar`` and ``tsbdy``/``tnbdy``.  In the
drop-in deployment this stays Fortran; it is restated here only so tests and bench.py
can build a qgcm_config.  The 3x3 LAPACK solve (DGETRF/DGETRS/DGERFS, :449-480) becomes
numpy.linalg.solve plus one refinement sweep.
"""
import math

import numpy as np

STEFAN = 5.67040e-8
SIGOV2 = 0.5 * STEFAN
NZ = 10001
NITMAX = 200
TMBTOL = 1.0e-13


def trapin(f, delz):
    """src/radsubs.f:596-634, Kahan-compensated extended trapezoid rule"""
    s = 0.5 * f[0]
    corr = 0.0
    for i in range(1, len(f) - 1):
        yadd = f[i] - corr
        sest = s + yadd
        corr = (sest - s) - yadd
        s = sest
    yadd = 0.5 * f[-1] - corr
    sest = s + yadd
    corr = (sest - s) - yadd
    s = sest - corr
    return delz * s


def fsprim(fspco, yrel, yla):
    """src/xfosubs.F:862-887"""
    return fspco * 0.5 * np.sin(math.pi * yrel / yla)


def radiat(p):
    nla, nlo = p.nla, p.nlo
    hat, tabsat, zopt = list(p.hat), list(p.tabsat), list(p.zopt)
    hmat, zm, gamma, fsbar, xlamda = p.hmat, p.zm, p.gamma, p.fsbar, p.xlamda
    hta = sum(hat)
    idx = np.arange(NZ, dtype=np.float64)
    # layer transmissivities, :91-97
    taum = math.exp(-hmat / zm)  # noqa: F841 (printed only in the reference)
    tauk = [0.0] * nla
    tauk[0] = math.exp(-(hat[0] - hmat) / zopt[0])
    tupmul = tauk[0]
    for k in range(1, nla):
        tauk[k] = math.exp(-hat[k] / zopt[k])
        tupmul *= tauk[k]
    uprad = [0.0] * nla
    dnrad = [0.0] * nla
    # layer 1, :101-121
    hbot, htop = hmat, hat[0]
    delz = (htop - hbot) / float(NZ - 1)
    zz = hbot + idx * delz
    fup = (tabsat[0] - gamma * zz) ** 4 * np.exp(-(htop - zz) / zopt[0])
    fdn = (tabsat[0] - gamma * zz) ** 4 * np.exp((hbot - zz) / zopt[0])
    uprad[0] = SIGOV2 * trapin(fup, delz) / zopt[0]
    dnrad[0] = SIGOV2 * trapin(fdn, delz) / zopt[0]
    rhstat = uprad[0]
    # upper layers, :125-147
    for k in range(1, nla):
        hbot = htop
        htop = hbot + hat[k]
        delz = hat[k] / float(NZ - 1)
        zz = hbot + idx * delz
        fup = (tabsat[k] - gamma * zz) ** 4 * np.exp(-(htop - zz) / zopt[k])
        fdn = (tabsat[k] - gamma * zz) ** 4 * np.exp((hbot - zz) / zopt[k])
        uprad[k] = SIGOV2 * trapin(fup, delz) / zopt[k]
        dnrad[k] = SIGOV2 * trapin(fdn, delz) / zopt[k]
        rhstat = rhstat * tauk[k] + uprad[k]
    # atmosphere mixed-layer mean temperature, :151-185
    rhstat = (-rhstat - fsbar) / tupmul
    rhstat = 2.0 * zm * rhstat / STEFAN
    tmbara = 300.0
    delz = hmat / float(NZ - 1)
    zz = idx * delz
    it = 0
    while True:
        fup = (tmbara - gamma * zz) ** 4 * np.exp(-(hmat - zz) / zm)
        upint = trapin(fup, delz)
        deltm = 0.25 * (rhstat - upint) * tmbara / upint
        tmbara = tmbara + 0.75 * deltm
        it += 1
        if it > NITMAX:
            raise RuntimeError("radiat: iteration for tmbara not converged")
        if not abs(deltm) > TMBTOL:
            break
    # ocean mixed-layer mean temperature, :188-204
    rhstoc = xlamda * tmbara + SIGOV2 * tmbara ** 4 - fsbar
    tmbaro = tmbara
    it = 0
    while True:
        tocold = tmbaro
        tmbaro = rhstoc / (xlamda + STEFAN * tocold ** 3)
        it += 1
        if it > NITMAX:
            raise RuntimeError("radiat: iteration for tmbaro not converged")
        if not abs(tmbaro - tocold) > TMBTOL:
            break
    toc = [p.tabsoc[k] - tmbaro for k in range(nlo)]
    tat = [tabsat[k] - tmbara for k in range(nla)]
    # mean-state fluxes, :214-233
    Fmupbar = SIGOV2 * upint / zm
    Fupbar = [0.0] * nla
    Fupbar[0] = Fmupbar * tauk[0] + uprad[0]
    for k in range(1, nla):
        Fupbar[k] = Fupbar[k - 1] * tauk[k] + uprad[k]
    Fdnbar = [0.0] * nla
    Fdnbar[nla - 1] = -dnrad[nla - 1]
    for k in range(nla - 2, -1, -1):
        Fdnbar[k] = Fdnbar[k + 1] * tauk[k] - dnrad[k]
    fspco = math.copysign(p.fspamp, p.fnot)
    # linearised coefficients, :286-372 (0-based [k][l] for the reference's (k+1,l+1))
    Aup = np.zeros((nla, max(nla - 1, 1)))
    Adown = np.zeros((nla, max(nla - 1, 1)))
    D0up = 4.0 * STEFAN * tmbaro ** 3
    Bmup = (SIGOV2 * (tmbara - gamma * hmat) ** 4 - Fmupbar) / zm
    Cmup = Bmup
    fup = (tmbara - gamma * zz) ** 3 * np.exp(-(hmat - zz) / zm)
    Dmup = 2.0 * STEFAN * trapin(fup, delz) / zm
    Bup = [0.0] * nla
    Cup = [0.0] * nla
    Dup = [0.0] * nla
    hbot, htop = hmat, hat[0]
    Aup[0, 0] = (-tauk[0] * Fmupbar - uprad[0] + SIGOV2 * (tabsat[0] - gamma * hat[0]) ** 4) / zopt[0]
    Bup[0] = tauk[0] * (Bmup + Fmupbar / zopt[0] - SIGOV2 * (tabsat[0] - gamma * hmat) ** 4 / zopt[0])
    Cup[0] = tauk[0] * (Cmup + Fmupbar / zopt[0] - SIGOV2 * (tabsat[0] - gamma * hmat) ** 4 / zopt[0])
    Dup[0] = Dmup * tauk[0]
    for k in range(1, nla):
        hbot = htop
        htop = hbot + hat[k]
        Bup[k] = Bup[k - 1] * tauk[k]
        Cup[k] = Cup[k - 1] * tauk[k]
        Dup[k] = Dup[k - 1] * tauk[k]
        for l in range(0, k - 1):
            Aup[k, l] = Aup[k - 1, l] * tauk[k]
        Aup[k, k - 1] = tauk[k] * (Aup[k - 1, k - 1] + Fupbar[k - 1] / zopt[k]
                                   - SIGOV2 * (tabsat[k] - gamma * hbot) ** 4 / zopt[k])
        if k < nla - 1:
            Aup[k, k] = (-tauk[k] * Fupbar[k - 1] - uprad[k] + SIGOV2 * (tabsat[k] - gamma * htop) ** 4) / zopt[k]
    htop = hta
    hbot = htop - hat[nla - 1]
    Adown[nla - 1, nla - 2] = (SIGOV2 * (tabsat[nla - 1] - gamma * hbot) ** 4 - dnrad[nla - 1]) / zopt[nla - 1]
    for k in range(nla - 2, 0, -1):
        htop = hbot
        hbot = htop - hat[k]
        for l in range(k + 1, nla - 1):
            Adown[k, l] = Adown[k + 1, l] * tauk[k]
        Adown[k, k - 1] = (Fdnbar[k + 1] * tauk[k] - dnrad[k] + SIGOV2 * (tabsat[k] - gamma * hbot) ** 4) / zopt[k]
        Adown[k, k] = tauk[k] * (Adown[k + 1, k] - Fdnbar[k + 1] / zopt[k]
                                 - SIGOV2 * (tabsat[k] - gamma * htop) ** 4 / zopt[k])
    for l in range(1, nla - 1):
        Adown[0, l] = Adown[1, l] * tauk[0]
    Adown[0, 0] = tauk[0] * (Adown[1, 0] - Fdnbar[1] / zopt[0] - SIGOV2 * (tabsat[0] - gamma * hat[0]) ** 4 / zopt[0])
    B1down = (Fdnbar[1] * tauk[0] - dnrad[0] + SIGOV2 * (tabsat[0] - gamma * hmat) ** 4) / zopt[0]
    C1down = B1down
    Dmdown = -2.0 * STEFAN * tmbara ** 3
    # radiation-balance initialisation coefficients, :412-492
    rbalar = np.zeros((nla, nla))
    for i in range(nla - 1):
        rbalar[0, i] = Adown[0, i]
    rbalar[0, nla - 1] = Dmup
    for k in range(1, nla - 1):
        for i in range(nla - 1):
            rbalar[k, i] = Adown[k + 1, i] + Aup[k, i]
        rbalar[k, nla - 1] = Dup[k]
    for i in range(nla - 1):
        rbalar[nla - 1, i] = Aup[nla - 1, i]
    rbalar[nla - 1, nla - 1] = Dup[nla - 1]
    balrhs = -np.ones(nla)
    rbafac = np.linalg.solve(rbalar, balrhs)
    rbafac = rbafac + np.linalg.solve(rbalar, balrhs - rbalar @ rbafac)
    rbetat = list(rbafac[: nla - 1])
    rbtmat = float(rbafac[nla - 1])
    rbtmoc = ((xlamda - Dmdown) * rbtmat - 1.0) / (xlamda + D0up)
    # perturbed-state balance, :506-548 (grids as src/q-gcm.F:395-431)
    dxa = p.ndxr * p.dxo
    yla = p.nyta * dxa
    ytarel = (np.arange(p.nyta) * dxa + 0.5 * dxa) - 0.5 * yla
    ypo = (p.ny1 - 1) * dxa + np.arange(p.nypo) * p.dxo
    ytorel = (ypo[: p.nyto] + 0.5 * p.dxo) - 0.5 * yla
    astbar = rbtmat * fsprim(fspco, ytarel, yla)
    sstbar = rbtmoc * fsprim(fspco, ytorel, yla)
    tnbdy = float(sstbar[-1])
    tsbdy = float(sstbar[0])
    rrcpat = 1.0 / (p.rhoat * p.cpat)
    rrcpdt = rrcpat / (tat[1] - tat[0])
    aface = [rrcpdt * (Adown[0, l] - Aup[nla - 1, l]) for l in range(nla - 1)]
    bface = rrcpdt * (B1down + Bmup - Bup[nla - 1])
    cface = rrcpdt * (C1down + Cmup - Cup[nla - 1])
    dface = rrcpdt * (Dmup - Dup[nla - 1])
    # Aup/Adown are (nla, nla-1) in the reference; store column-major with ld = nla
    return dict(toc=toc, tat=tat, tmbara=tmbara, tmbaro=tmbaro, tsbdy=tsbdy, tnbdy=tnbdy, fspco=fspco,
                Bmup=Bmup, B1down=B1down, Cmup=Cmup, C1down=C1down, D0up=D0up, Dmup=Dmup, Dmdown=Dmdown,
                bface=bface, cface=cface, dface=dface, Aup=Aup, Adown=Adown, Bup=Bup, Cup=Cup, Dup=Dup,
                rbetat=rbetat, aface=aface, rbtmat=rbtmat, rbtmoc=rbtmoc, sstbar=sstbar, astbar=astbar,
                Fupbar=Fupbar, fsbar=fsbar)

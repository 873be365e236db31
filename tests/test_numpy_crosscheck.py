"""Second, independent implementation of the step routines, written from the Fortran with
vectorised numpy/scipy (scipy.fft.dst / rfft, scipy.linalg.solve_banded), compared with the
oracle's loop restatement.  Interior points only where the
boundary zoo would make the re-derivation as long as the oracle itself."""
import numpy as np
import pytest
import scipy.fft as sf
import scipy.linalg as sla

from util import small_configs, rel_l2


def lap5(f, dxm2):
    return (f[1:-1, :-2] + f[:-2, 1:-1] + f[2:, 1:-1] + f[1:-1, 2:] - 4.0 * f[1:-1, 1:-1]) * dxm2


def jac9(q, p):
    """Arakawa 9-point J(q,p) numerator at interior points (src/qgosubs.F:376-388)"""
    c = (slice(1, -1), slice(1, -1))
    def s(a, di, dj):
        return a[1 + di: a.shape[0] - 1 + di, 1 + dj: a.shape[1] - 1 + dj]
    return ((s(q, 1, 0) - s(q, -1, 0)) * (s(p, 0, 1) - s(p, 0, -1)) + (s(q, 0, -1) - s(q, 0, 1)) * (s(p, 1, 0) - s(p, -1, 0))
            + s(q, 1, 0) * (s(p, 1, 1) - s(p, 1, -1)) - s(q, -1, 0) * (s(p, -1, 1) - s(p, -1, -1))
            - s(q, 0, 1) * (s(p, 1, 1) - s(p, -1, 1)) + s(q, 0, -1) * (s(p, 1, -1) - s(p, -1, -1))
            + s(p, 0, 1) * (s(q, 1, 1) - s(q, -1, 1)) - s(p, 0, -1) * (s(q, 1, -1) - s(q, -1, -1))
            - s(p, 1, 0) * (s(q, 1, 1) - s(q, 1, -1)) + s(p, -1, 0) * (s(q, -1, 1) - s(q, -1, -1)))


def test_qgostep_box_interior(qg, pyorc):
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.oml()
    sh = (p.nxpo, p.nypo, p.nlo)
    po, pom, qo, qom = (m.get_field(n, sh) for n in ("po", "pom", "qo", "qom"))
    wek, ent = m.get_field("wekpo", sh[:2]), m.get_field("entoc", sh[:2])
    m.qgostep()
    qnew = m.get_field("qo", sh)
    dxm2 = 1.0 / p.dxo ** 2
    bcf = p.bccooc * dxm2 / (0.5 * p.bccooc + 1.0)
    adf = 1.0 / (12.0 * p.dxo ** 2 * p.fnot)
    tdt = 2.0 * p.dto
    for k in range(p.nlo):
        d2 = np.zeros(sh[:2])
        d2[1:-1, 1:-1] = lap5(pom[:, :, k], dxm2)
        d2[:, 0] = bcf * (pom[:, 1, k] - pom[:, 0, k]); d2[:, -1] = bcf * (pom[:, -2, k] - pom[:, -1, k])
        d2[0, 1:-1] = bcf * (pom[1, 1:-1, k] - pom[0, 1:-1, k]); d2[-1, 1:-1] = bcf * (pom[-2, 1:-1, k] - pom[-1, 1:-1, k])
        d4 = np.zeros(sh[:2])
        d4[1:-1, 1:-1] = lap5(d2, dxm2)
        d4[:, 0] = bcf * (d2[:, 1] - d2[:, 0]); d4[:, -1] = bcf * (d2[:, -2] - d2[:, -1])
        d4[0, 1:-1] = bcf * (d2[1, 1:-1] - d2[0, 1:-1]); d4[-1, 1:-1] = bcf * (d2[-2, 1:-1] - d2[-1, 1:-1])
        d6 = lap5(d4, dxm2)
        dq = adf * jac9(qo[:, :, k], po[:, :, k]) + (p.ah2oc[k] / p.fnot) * d4[1:-1, 1:-1] - (p.ah4oc[k] / p.fnot) * d6
        if k == 0:
            dq = dq + (p.fnot / p.hoc[0]) * (wek[1:-1, 1:-1] - ent[1:-1, 1:-1])
        if k == 1:
            dq = dq + (p.fnot / p.hoc[1]) * ent[1:-1, 1:-1]
        if k == p.nlo - 1:
            dq = dq - 0.5 * np.sign(p.fnot) * p.delek / p.hoc[-1] * d2[1:-1, 1:-1]
        want = qom[1:-1, 1:-1, k] + tdt * dq
        assert rel_l2(qnew[1:-1, 1:-1, k], want) <= 1e-13, k


def test_qgostep_channel_interior(qg, pyorc):
    p = small_configs(qg)["chan_so"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    sh = (p.nxpo, p.nypo, p.nlo)
    po, pom, qo, qom = (m.get_field(n, sh) for n in ("po", "pom", "qo", "qom"))
    wek, ent = m.get_field("wekpo", sh[:2]), m.get_field("entoc", sh[:2])
    m.qgostep()
    qnew = m.get_field("qo", sh)
    dxm2 = 1.0 / p.dxo ** 2
    bcf = p.bccooc * dxm2 / (0.5 * p.bccooc + 1.0)
    adf = 1.0 / (12.0 * p.dxo ** 2 * p.fnot)
    tdt = 2.0 * p.dto

    def per(f):   # periodic extension by one column each side (column nxp == column 1)
        return np.vstack([f[-2:-1], f, f[1:2]])

    for k in range(p.nlo):
        pm = per(pom[:, :, k])
        d2 = np.zeros(sh[:2])
        d2[:, 1:-1] = lap5(pm, dxm2)
        d2[:, 0] = bcf * (pom[:, 1, k] - pom[:, 0, k]); d2[:, -1] = bcf * (pom[:, -2, k] - pom[:, -1, k])
        d4 = np.zeros(sh[:2])
        d4[:, 1:-1] = lap5(per(d2), dxm2)
        d4[:, 0] = bcf * (d2[:, 1] - d2[:, 0]); d4[:, -1] = bcf * (d2[:, -2] - d2[:, -1])
        d6 = lap5(per(d4), dxm2)
        dq = adf * jac9(per(qo[:, :, k]), per(po[:, :, k])) + (p.ah2oc[k] / p.fnot) * d4[:, 1:-1] - (p.ah4oc[k] / p.fnot) * d6
        if k == 0:
            dq = dq + (p.fnot / p.hoc[0]) * (wek[:, 1:-1] - ent[:, 1:-1])
        if k == 1:
            dq = dq + (p.fnot / p.hoc[1]) * ent[:, 1:-1]
        if k == p.nlo - 1:
            dq = dq - 0.5 * np.sign(p.fnot) * p.delek / p.hoc[-1] * d2[:, 1:-1]
        want = qom[:, 1:-1, k] + tdt * dq
        assert rel_l2(qnew[:, 1:-1, k], want) <= 1e-13, k


def test_hsbxoc_against_scipy(qg, pyorc):
    """box solver = DST-I rows, banded solve per wavenumber, DST-I rows (src/ocisubs.F:461-509)"""
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    rng = np.random.default_rng(11)
    nxp, nyp, nxt = p.nxpo, p.nypo, p.nxto
    rhs = rng.standard_normal((nxp, nyp))
    a = 1.0 / p.dxo ** 2
    rd = cfg.rdm2oc[2]
    k = np.arange(1, nxt)
    b = -2 * a + 2 * a * (np.cos(k * np.pi / nxt) - 1.0) - rd
    bfull = np.zeros(nxt); bfull[: nxt - 1] = b
    got = m.helmholtz(0, rhs, bfull)
    spec = sf.dst(rhs[1:-1, 1:-1], type=1, axis=0)
    sol = np.empty_like(spec)
    n = nyp - 2
    for i in range(nxt - 1):
        ab = np.zeros((3, n)); ab[0, 1:] = a; ab[1, :] = b[i]; ab[2, :-1] = a
        sol[i] = sla.solve_banded((1, 1), ab, spec[i])
    want = sf.dst(sol * (0.5 / nxt), type=1, axis=0)
    assert rel_l2(got[1:-1, 1:-1], want) <= 1e-13


def test_hscyoc_against_scipy(qg, pyorc):
    """channel solver = rfft rows, banded solve per wavenumber, irfft rows (src/ocisubs.F:566-604)"""
    p = small_configs(qg)["chan_so"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    rng = np.random.default_rng(12)
    nxp, nyp, nxt = p.nxpo, p.nypo, p.nxto
    rhs = rng.standard_normal((nxp, nyp)); rhs[-1] = rhs[0]
    a = 1.0 / p.dxo ** 2
    rd = cfg.rdm2oc[1]
    bd2 = np.zeros(nxt)
    for i in range(2, nxt // 2 + 1):
        bd2[2 * i - 3] = -2 * a + 2 * a * (np.cos((i - 1) * 2 * np.pi / nxt) - 1.0)
        bd2[2 * i - 2] = bd2[2 * i - 3]
    bd2[0] = -2 * a; bd2[nxt - 1] = -6 * a
    got = m.helmholtz(0, rhs, bd2 - rd)
    spec = sf.rfft(rhs[:nxt, 1:-1], axis=0)
    kk = np.arange(nxt // 2 + 1)
    bk = -2 * a + 2 * a * (np.cos(kk * 2 * np.pi / nxt) - 1.0) - rd
    n = nyp - 2
    sol = np.empty_like(spec)
    for i in range(nxt // 2 + 1):
        ab = np.zeros((3, n)); ab[0, 1:] = a; ab[1, :] = bk[i]; ab[2, :-1] = a
        sol[i] = sla.solve_banded((1, 1), ab, spec[i].real) + 1j * sla.solve_banded((1, 1), ab, spec[i].imag)
    want = sf.irfft(sol, n=nxt, axis=0)
    assert rel_l2(got[:nxt, 1:-1], want) <= 1e-13
    assert np.array_equal(got[-1], got[0])


def test_ocinvq_box_against_numpy(qg, pyorc):
    """layer->mode, invert, constrain, mode->layer with numpy linear algebra"""
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.oml(); m.qgostep()
    sh = (p.nxpo, p.nypo, p.nlo)
    qo = m.get_field("qo", sh)
    ochom = m.get_field("ochom", (p.nxpo, p.nypo, p.nlo - 1))
    s0 = m.get_scalars()
    nl = p.nlo
    l2m = np.array(cfg.ctl2moc[: nl * nl]).reshape(nl, nl, order="F")
    m2l = np.array(cfg.ctm2loc[: nl * nl]).reshape(nl, nl, order="F")
    yrel = (p.ny1 - 1) * p.dxa + np.arange(p.nypo) * p.dxo - 0.5 * p.nyta * p.dxa
    ql = qo - (p.beta * yrel)[None, :, None]
    wrk = p.fnot * np.einsum("km,ijk->ijm", l2m, ql)
    a = 1.0 / p.dxo ** 2
    nxt = p.nxto
    kk = np.arange(1, nxt)
    xin = np.zeros(nl)
    pm = np.zeros(sh)
    for mo in range(nl):
        b = -2 * a + 2 * a * (np.cos(kk * np.pi / nxt) - 1.0) - cfg.rdm2oc[mo]
        spec = sf.dst(wrk[1:-1, 1:-1, mo], type=1, axis=0)
        n = p.nypo - 2
        sol = np.empty_like(spec)
        for i in range(nxt - 1):
            ab = np.zeros((3, n)); ab[0, 1:] = a; ab[1, :] = b[i]; ab[2, :-1] = a
            sol[i] = sla.solve_banded((1, 1), ab, spec[i])
        pm[1:-1, 1:-1, mo] = sf.dst(sol * (0.5 / nxt), type=1, axis=0)
        xin[mo] = pm[:, :, mo].sum() * p.dxo ** 2
    tdt = 2 * p.dto
    dpi_new = np.array(s0.dpiocp[: nl - 1]) - tdt * np.array(p.gpoc) * np.array([s0.xon[0]] + [0.0] * (nl - 2))
    cdiffo = np.array(s0.cdiffo[: nl * (nl - 1)]).reshape(nl, nl - 1, order="F")
    cdhoc = np.array(s0.cdhoc[: (nl - 1) ** 2]).reshape(nl - 1, nl - 1, order="F")
    hcl = np.linalg.solve(cdhoc, dpi_new - cdiffo.T @ xin)
    for mo in range(1, nl):
        pm[:, :, mo] += hcl[mo - 1] * ochom[:, :, mo - 1]
    want = np.einsum("mk,ijm->ijk", m2l, pm)
    m.ocinvq()
    got = m.get_field("po", sh)
    assert rel_l2(got, want) <= 1e-12


def test_oml_interior_against_numpy(qg, pyorc):
    """C-grid advection + del2/del4 diffusion + sst update at points two cells from any wall"""
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    nxt, nyt = p.nxto, p.nyto
    po = m.get_field("po", (p.nxpo, p.nypo, p.nlo))[:, :, 0]
    tx, ty = m.get_field("tauxo", (p.nxpo, p.nypo)), m.get_field("tauyo", (p.nxpo, p.nypo))
    sst, sstm = m.get_field("sst", (nxt, nyt)), m.get_field("sstm", (nxt, nyt))
    wek, fnet = m.get_field("wekto", (nxt, nyt)), m.get_field("fnetoc", (nxt, nyt))
    m.oml()
    got = m.get_field("sst", (nxt, nyt))
    uvg = p.ycexp / (p.dxo * p.fnot)
    rh = 0.5 / (p.fnot * p.hmoc)
    hdx = 0.5 / p.dxo
    # face velocities on the whole grid: u(i,j) west face of cell i (p column i), v(i,j) south face
    u = -uvg * (po[:, 1:] - po[:, :-1]) + rh * (ty[:, 1:] + ty[:, :-1])          # (nxp, nyt)
    v = uvg * (po[1:, :] - po[:-1, :]) - rh * (tx[1:, :] + tx[:-1, :])            # (nxt, nyp)
    I = slice(2, nxt - 2); J = slice(2, nyt - 2)
    def sh_(a, di, dj):
        return a[2 + di: nxt - 2 + di, 2 + dj: nyt - 2 + dj]
    um, up = u[2:nxt - 2, J], u[3:nxt - 1, J]
    vm, vp = v[I, 2:nyt - 2], v[I, 3:nyt - 1]
    hx = hdx * (up * (sh_(sst, 0, 0) + sh_(sst, 1, 0)) - um * (sh_(sst, -1, 0) + sh_(sst, 0, 0)))
    hy = hdx * (vp * (sh_(sst, 0, 1) + sh_(sst, 0, 0)) - vm * (sh_(sst, 0, 0) + sh_(sst, 0, -1)))
    d2 = np.zeros((nxt, nyt))
    d2[1:-1, 1:-1] = sstm[1:-1, :-2] + sstm[:-2, 1:-1] + sstm[2:, 1:-1] + sstm[1:-1, 2:] - 4 * sstm[1:-1, 1:-1]
    dxm2 = 1.0 / p.dxo ** 2
    d4 = sh_(d2, 0, -1) + sh_(d2, -1, 0) + sh_(d2, 1, 0) + sh_(d2, 0, 1) - 4 * sh_(d2, 0, 0)
    rhs = -(hx + hy) + p.st2d * dxm2 * sh_(d2, 0, 0) - p.st4d * dxm2 ** 2 * d4
    toc1 = cfg.toc[0]
    rrcp = 1.0 / (p.rhooc * p.cpoc)
    new = sh_(sstm, 0, 0) + 2 * p.dto * (rhs + (rrcp * sh_(fnet, 0, 0) + 0.5 * sh_(wek, 0, 0) * (sh_(sstm, 0, 0) + toc1)) / p.hmoc)
    new = new + np.maximum(0.0, toc1 - new)
    assert rel_l2(got[I, J], new) <= 1e-13


# ---- boundary potential vorticity: ocqbdy (src/vorsubs.F:245-388) and atqzbd (:396-480) ----
def _amat(flat, n):
    return np.array(flat[: n * n]).reshape((n, n), order="F")


@pytest.mark.parametrize("case", ["box_dg", "chan_so"])
def test_ocqbdy_matches_numpy(qg, pyorc, case):
    p = small_configs(qg)[case]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.ocean_step()
    sh = (p.nxpo, p.nypo, p.nlo)
    po = m.get_field("po", sh)
    ddyn = m.get_field("ddynoc", sh[:2]) + 1e-9 * np.cos(np.arange(p.nxpo))[:, None]     # exercise the topography term
    m.set_field("ddynoc", ddyn)
    m.set_field("qo", np.full(sh, 7.0))            # every boundary value must be overwritten
    m.ocqbdy()
    q = m.get_field("qo", sh)
    A = _amat(cfg.amatoc, p.nlo)
    bcf = p.bccooc / p.dxo ** 2 / (0.5 * p.bccooc + 1.0) / p.fnot
    ypo = (p.ny1 - 1) * p.ndxr * p.dxo + np.arange(p.nypo) * p.dxo
    yrel = ypo - 0.5 * p.nyta * p.ndxr * p.dxo
    fAp = p.fnot * np.einsum("kl,ijl->ijk", A, po)          # f0 (A p)_k at every point
    want = np.full(sh, 7.0)
    top = np.zeros(sh[:2] + (p.nlo,)); top[:, :, -1] = ddyn
    want[:, 0, :] = bcf * (po[:, 1, :] - po[:, 0, :]) - fAp[:, 0, :] + p.beta * yrel[0] + top[:, 0, :]
    want[:, -1, :] = bcf * (po[:, -2, :] - po[:, -1, :]) - fAp[:, -1, :] + p.beta * yrel[-1] + top[:, -1, :]
    if not p.has("cyclic_ocean"):
        by = (p.beta * yrel)[1:-1, None]
        want[0, 1:-1, :] = bcf * (po[1, 1:-1, :] - po[0, 1:-1, :]) - fAp[0, 1:-1, :] + by + top[0, 1:-1, :]
        want[-1, 1:-1, :] = bcf * (po[-2, 1:-1, :] - po[-1, 1:-1, :]) - fAp[-1, 1:-1, :] + by + top[-1, 1:-1, :]
    assert rel_l2(q, want) <= 1e-13


def test_atqzbd_matches_numpy_with_its_quirk(qg, pyorc):
    """the top layer's southern row takes pa(i,2,nla), not pa(i,1,nla), in the stretching term
    (src/vorsubs.F:470)"""
    p = qg.named_config("dg_coupled").scaled(6, 5, ndxr=16, name="cpl_dg")
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.run(1, 2)
    nxp, nyp, nl = p.nxta + 1, p.nyta + 1, p.nla
    sh = (nxp, nyp, nl)
    pa = m.get_field("pa", sh)
    ddyn = m.get_field("ddynat", sh[:2])
    m.set_field("qa", np.full(sh, 7.0))
    m.atqzbd()
    q = m.get_field("qa", sh)
    A = _amat(cfg.amatat, nl)
    dxa = p.ndxr * p.dxo
    zbf = p.bccoat / dxa ** 2 / (0.5 * p.bccoat + 1.0) / p.fnot
    yrel = np.arange(nyp) * dxa - 0.5 * p.nyta * dxa
    fAp = p.fnot * np.einsum("kl,ijl->ijk", A, pa)
    want = np.full(sh, 7.0)
    bot = np.zeros(sh); bot[:, :, 0] = ddyn
    want[:, 0, :] = zbf * (pa[:, 1, :] - pa[:, 0, :]) - fAp[:, 0, :] + p.beta * yrel[0] + bot[:, 0, :]
    want[:, -1, :] = zbf * (pa[:, -2, :] - pa[:, -1, :]) - fAp[:, -1, :] + p.beta * yrel[-1] + bot[:, -1, :]
    # the quirk: f0Ac multiplies pa(i,2,nla) on the southern row of the top layer
    want[:, 0, -1] = (zbf * (pa[:, 1, -1] - pa[:, 0, -1])
                      - p.fnot * (A[-1, -2] * pa[:, 0, -2] + A[-1, -1] * pa[:, 1, -1]) + p.beta * yrel[0])
    assert rel_l2(q, want) <= 1e-13
    plain = fAp[:, 0, -1]
    assert rel_l2(q[:, 0, -1], zbf * (pa[:, 1, -1] - pa[:, 0, -1]) - plain + p.beta * yrel[0]) > 1e-9      # and it matters


def test_qgastep_matches_numpy(qg, pyorc):
    """atmosphere vorticity step (src/qgasubs.F:45-317): periodic channel, del-6th friction only,
    entrainment/Ekman forcing with the atmosphere's signs (:122-124)"""
    p = qg.named_config("dg_coupled").scaled(6, 5, ndxr=16, name="cpl_dg")
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.run(1, 2)
    m.aml()
    nxp, nyp, nl = p.nxta + 1, p.nyta + 1, p.nla
    sh = (nxp, nyp, nl)
    pa, pam, qa, qam = (m.get_field(n, sh) for n in ("pa", "pam", "qa", "qam"))
    wek, ent = m.get_field("wekpa", sh[:2]), m.get_field("entat", sh[:2])
    m.qgastep()
    qnew = m.get_field("qa", sh)
    dxa = p.ndxr * p.dxo
    dxm2 = 1.0 / dxa ** 2
    zbf = p.bccoat * dxm2 / (0.5 * p.bccoat + 1.0)
    adf = 1.0 / (12.0 * dxa * dxa * p.fnot)

    def lap(f):            # f on the nxta distinct columns; mixed condition on the zonal walls
        out = np.empty_like(f)
        out[:, 1:-1] = (f[:, :-2] + np.roll(f, 1, axis=0)[:, 1:-1] + np.roll(f, -1, axis=0)[:, 1:-1] + f[:, 2:] - 4.0 * f[:, 1:-1]) * dxm2
        out[:, 0] = zbf * (f[:, 1] - f[:, 0])
        out[:, -1] = zbf * (f[:, -2] - f[:, -1])
        return out

    def sx(a, d):
        return np.roll(a, -d, axis=0)

    for k in range(nl):
        P, Q = pa[:-1, :, k], qa[:-1, :, k]            # drop the repeated column
        d6 = lap(lap(lap(pam[:-1, :, k])))
        c = slice(1, -1)
        jac = ((sx(Q, 1)[:, c] - sx(Q, -1)[:, c]) * (P[:, 2:] - P[:, :-2]) + (Q[:, :-2] - Q[:, 2:]) * (sx(P, 1)[:, c] - sx(P, -1)[:, c])
               + sx(Q, 1)[:, c] * (sx(P, 1)[:, 2:] - sx(P, 1)[:, :-2]) - sx(Q, -1)[:, c] * (sx(P, -1)[:, 2:] - sx(P, -1)[:, :-2])
               - Q[:, 2:] * (sx(P, 1)[:, 2:] - sx(P, -1)[:, 2:]) + Q[:, :-2] * (sx(P, 1)[:, :-2] - sx(P, -1)[:, :-2])
               + P[:, 2:] * (sx(Q, 1)[:, 2:] - sx(Q, -1)[:, 2:]) - P[:, :-2] * (sx(Q, 1)[:, :-2] - sx(Q, -1)[:, :-2])
               - sx(P, 1)[:, c] * (sx(Q, 1)[:, 2:] - sx(Q, 1)[:, :-2]) + sx(P, -1)[:, c] * (sx(Q, -1)[:, 2:] - sx(Q, -1)[:, :-2]))
        dq = adf * jac - (p.ah4at[k] / p.fnot) * d6[:, c]
        if k == 0:
            dq = dq + (p.fnot / p.hat[0]) * (ent[:-1, c] - wek[:-1, c])
        if k == 1:
            dq = dq - (p.fnot / p.hat[1]) * ent[:-1, c]
        want = qam[:-1, c, k] + 2.0 * p.dta * dq
        assert rel_l2(qnew[:-1, c, k], want) <= 1e-12, k
        assert np.array_equal(qnew[-1, c, k], qnew[0, c, k])              # the periodic column is a copy
    # zonal boundary rows are left for atqzbd; qam takes the old qa there (:138-145)
    assert np.array_equal(m.get_field("qam", sh)[:, 0, :], qa[:, 0, :])


@pytest.mark.parametrize("case", ["box_dg", "chan_so"])
def test_time_level_average_matches_numpy(qg, pyorc, case):
    """src/q-gcm.F:1328-1366: x <- (x + xm)/2 for qo, po, sst and the constraint scalars; the
    lagged levels are left alone"""
    p = small_configs(qg)[case]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.ocean_step()
    m.ocean_step()
    before = {n: m.get_field(n) for n in ("po", "pom", "qo", "qom", "sst", "sstm", "entoc")}
    s0 = m.get_scalars().as_dict()
    m.tlavg_ocean()
    s1 = m.get_scalars().as_dict()
    for new, old in (("po", "pom"), ("qo", "qom"), ("sst", "sstm")):
        assert np.array_equal(m.get_field(new), 0.5 * (before[new] + before[old])), new
        assert np.array_equal(m.get_field(old), before[old]), old
    assert np.array_equal(m.get_field("entoc"), before["entoc"])
    for k in range(p.nlo - 1):
        assert s1["dpioc"][k] == 0.5 * (s0["dpioc"][k] + s0["dpiocp"][k])
        assert s1["dpiocp"][k] == s0["dpiocp"][k]
    if p.has("cyclic_ocean"):
        for k in range(p.nlo):
            assert s1["ocncs"][k] == 0.5 * (s0["ocncs"][k] + s0["ocncsp"][k])
            assert s1["ocncn"][k] == 0.5 * (s0["ocncn"][k] + s0["ocncnp"][k])
    else:
        assert s1["ocncs"] == s0["ocncs"]


def test_aml_interior_against_numpy(qg, pyorc):
    """atmosphere mixed layer (src/amlsubs.F:47-563): C-grid advection of ast and hmixa by the
    layer-1 geostrophic wind plus uekat/vekat, del-sqd/del-4th diffusion of ast, diffusion and
    relaxation of hmixa with its floor, diabatic term, convective adjustment, entrainment
    averaged to p points with the eta and topography terms.  Rows away from the zonal walls."""
    p = qg.named_config("dg_coupled").scaled(6, 5, ndxr=16, name="cpl_dg")
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.run(1, 2)
    nxt, nyt, nxp, nyp, nl = p.nxta, p.nyta, p.nxta + 1, p.nyta + 1, p.nla
    g = lambda n, sh: m.get_field(n, sh)
    ast, astm, hm, hmm = g("ast", (nxt, nyt)), g("astm", (nxt, nyt)), g("hmixa", (nxt, nyt)), g("hmixam", (nxt, nyt))
    fnet, wekta, xc1 = g("fnetat", (nxt, nyt)), g("wekta", (nxt, nyt)), g("xc1ast", (nxt, nyt))
    uek, vek = g("uekat", (nxp, nyt)), g("vekat", (nxt, nyp))
    pa, pam, dtop = g("pa", (nxp, nyp, nl)), g("pam", (nxp, nyp, nl)), g("dtopat", (nxp, nyp))
    m.aml()
    ast_n, hm_n, ent_n = g("ast", (nxt, nyt)), g("hmixa", (nxt, nyt)), g("entat", (nxp, nyp))
    assert np.array_equal(g("astm", (nxt, nyt)), ast) and np.array_equal(g("hmixam", (nxt, nyt)), hm)
    dxa = p.ndxr * p.dxo
    rdxf0, hdxm1, dxm2 = 1.0 / (dxa * p.fnot), 0.5 / dxa, 1.0 / dxa ** 2
    tdt = 2.0 * p.dta
    rrcpat = 1.0 / (cfg.rhoat * cfg.cpat)
    p1 = pa[:, :, 0]
    u = -rdxf0 * (p1[:, 1:] - p1[:, :-1]) + uek          # (nxp, nyt): x faces
    v = rdxf0 * (p1[1:, :] - p1[:-1, :]) + vek           # (nxt, nyp): y faces
    rows = slice(2, nyt - 2)                             # two T rows clear of the walls: plain stencils only
    W = lambda a: np.roll(a, 1, axis=0)                  # value at i-1 (periodic)
    E = lambda a: np.roll(a, -1, axis=0)

    def flux_div(f):
        um, up = u[:-1], u[1:]
        x = hdxm1 * (up * (f + E(f)) - um * (W(f) + f))
        y = np.zeros_like(f)
        y[:, 1:-1] = hdxm1 * (v[:, 2:-1] * (f[:, 2:] + f[:, 1:-1]) - v[:, 1:-2] * (f[:, 1:-1] + f[:, :-2]))
        return x + y

    def lap(f):
        out = np.zeros_like(f)
        out[:, 1:-1] = f[:, :-2] + W(f)[:, 1:-1] + E(f)[:, 1:-1] + f[:, 2:] - 4.0 * f[:, 1:-1]
        return out

    d2t = lap(astm)
    tmrhs = -flux_div(ast) + p.at2d * dxm2 * d2t - p.at4d * dxm2 ** 2 * lap(d2t)
    hmrhs = -flux_div(hm) + p.ahmd * dxm2 * lap(hmm)
    tat1, tat2 = cfg.tat[0], cfg.tat[1]
    hdrcdt = p.hmadmp * rrcpat * tdt
    diabcr = tat1 - 2.0 * hdrcdt
    cold = astm <= diabcr
    with np.errstate(divide="ignore", invalid="ignore"):
        hnew = hmm + tdt * hmrhs - hdrcdt * (hmm - p.hmat) / (tat1 - astm)
    dhfix = np.maximum(p.hmamin - hnew, 0.0)
    hnew = np.where(cold, hnew + dhfix, p.hmat)
    dtfix = np.where(cold, dhfix * (tat1 - astm) / hmm, 0.0)
    trhtot = tmrhs + rrcpat * fnet / hmm - wekta * astm / p.hmat
    astnew = astm + tdt * trhtot + dtfix
    dtanew = tat1 - astnew
    entfac = 1.0 / (tdt * (tat2 - tat1))
    conena = entfac * hm * np.minimum(0.0, dtanew)
    xfa = p.xcexp * cfg.bface * (hmm - p.hmat) + cfg.dface * (p.xcexp * astm + xc1) - p.xcexp * conena
    astnew = astnew + np.minimum(0.0, dtanew)
    assert rel_l2(ast_n[:, rows], astnew[:, rows]) <= 1e-12
    assert rel_l2(hm_n[:, rows], hnew[:, rows]) <= 1e-12
    # entrainment at p points i = 2..nxpa-1 of the same rows: four-point average plus the p-point terms
    avg = 0.25 * (xfa[:-1, :-1] + xfa[1:, :-1] + xfa[:-1, 1:] + xfa[1:, 1:])       # p point (i+1, j+1), 0-based
    adp = sum(cfg.aface[l] / cfg.gpat[l] * (pam[:, :, l] - pam[:, :, l + 1]) for l in range(nl - 1))
    want = avg + (adp + cfg.cface * dtop)[1:-1, 1:-1]
    prow = slice(3, nyt - 2)
    assert rel_l2(ent_n[1:-1, 1:-1][:, prow], want[:, prow]) <= 1e-12


def test_valids_matches_numpy(qg, pyorc):
    """extreme-value scan and perturbed layer thicknesses of valids (src/valsubs.F:43-630): a
    state with an interface displaced far enough for the thin-layer percentages to be non-zero"""
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    sh = (p.nxpo, p.nypo, p.nlo)
    po = m.get_field("po", sh)
    x = np.linspace(-1.0, 1.0, p.nxpo)[:, None]
    y = np.linspace(-1.0, 1.0, p.nypo)[None, :]
    po[:, :, 0] -= 330.0 * cfg.gpoc[0] * np.exp(-(x ** 2 + y ** 2) / 0.1)      # lifts interface 1 by ~330 m into the 350 m top layer
    m.set_field("po", po)
    r = m.valids().as_dict()
    for name, f in (("poc", po), ("qoc", m.get_field("qo")), ("sst", m.get_field("sst")), ("wto", m.get_field("wekto"))):
        assert r[name + "min"] == f.min() and r[name + "max"] == f.max(), name
    eta = [(po[:, :, k + 1] - po[:, :, k]) / cfg.gpoc[k] for k in range(p.nlo - 1)]
    dtop = cfg.hoc[p.nlo - 1] / p.fnot * m.get_field("ddynoc", sh[:2])
    hf = [cfg.hoc[0] - eta[0]] + [cfg.hoc[k] - eta[k] + eta[k - 1] for k in range(1, p.nlo - 1)] + [cfg.hoc[p.nlo - 1] + eta[-1] - dtop]
    assert np.isclose(r["hfmint"], hf[0].min(), rtol=1e-14) and np.isclose(r["hfmaxt"], hf[0].max(), rtol=1e-14)
    assert np.isclose(r["hfminb"], hf[-1].min(), rtol=1e-14) and np.isclose(r["hfmaxb"], hf[-1].max(), rtol=1e-14)
    mid = np.array(hf[1:-1])
    assert np.isclose(r["hfmini"], mid.min(), rtol=1e-14) and np.isclose(r["hfmaxi"], mid.max(), rtol=1e-14)
    w = np.ones(p.nxpo)[:, None] * np.ones(p.nypo)[None, :]
    w[0] *= 0.5; w[-1] *= 0.5; w[:, 0] *= 0.5; w[:, -1] *= 0.5
    for k in range(p.nlo):
        pct = 100.0 * (w * (hf[k] < 100.0)).sum() / (p.nxto * p.nyto)
        assert np.isclose(r["hfbad"][k], pct, rtol=1e-12, atol=1e-12), (k, r["hfbad"][k], pct)
    assert r["hfbad"][0] > 0.0                      # the bump did thin the top layer below 100 m somewhere
    extreme = abs(po).max() >= 1.0e4
    assert r["solnok"] == int(not extreme and max(r["hfbad"][:p.nlo]) <= 20.0)


def test_ocinvq_channel_against_numpy(qg, pyorc):
    """channel ocean (src/ocisubs.F:174-327): constraint right-hand sides from the boundary
    integrals, leapfrog of ocncs/ocncn, line integrals of the inhomogeneous modes, c1/c2/c3,
    the homogeneous corrections pbhoc / pch1oc / pch2oc, continuity update and monitors"""
    p = small_configs(qg)["chan_so"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.ocean_step()
    m.oml(); m.qgostep()
    nl, nxp, nyp, nxt = p.nlo, p.nxpo, p.nypo, p.nxto
    sh = (nxp, nyp, nl)
    qo, po_old = m.get_field("qo", sh), m.get_field("po", sh)
    pch1, pch2 = m.get_field("pch1oc", (nyp, nl - 1)), m.get_field("pch2oc", (nyp, nl - 1))
    pbh = m.get_field("pbhoc")
    s0 = m.get_scalars().as_dict()
    l2m = np.array(cfg.ctl2moc[: nl * nl]).reshape(nl, nl, order="F")
    m2l = np.array(cfg.ctm2loc[: nl * nl]).reshape(nl, nl, order="F")
    yrel = (p.ny1 - 1) * p.dxa + np.arange(nyp) * p.dxo - 0.5 * p.nyta * p.dxa
    wrk = np.zeros(sh)
    wrk[:, 1:-1, :] = p.fnot * np.einsum("km,ijk->ijm", l2m, (qo - (p.beta * yrel)[None, :, None])[:, 1:-1, :])
    a = 1.0 / p.dxo ** 2
    kk = np.arange(nxt // 2 + 1)
    pm = np.zeros(sh)
    xin = np.zeros(nl)
    n = nyp - 2
    wts = np.ones(nxp); wts[0] = wts[-1] = 0.5
    for mo in range(nl):
        bk = -2 * a + 2 * a * (np.cos(kk * 2 * np.pi / nxt) - 1.0) - cfg.rdm2oc[mo]
        spec = sf.rfft(wrk[:nxt, 1:-1, mo], axis=0)
        sol = np.empty_like(spec)
        for i in range(nxt // 2 + 1):
            ab = np.zeros((3, n)); ab[0, 1:] = a; ab[1, :] = bk[i]; ab[2, :-1] = a
            sol[i] = sla.solve_banded((1, 1), ab, spec[i].real) + 1j * sla.solve_banded((1, 1), ab, spec[i].imag)
        pm[:nxt, 1:-1, mo] = sf.irfft(sol, n=nxt, axis=0)
        pm[-1, :, mo] = pm[0, :, mo]
        xin[mo] = (wts @ pm[:, 1:-1, mo]).sum() * p.dxo ** 2         # xintp: the wall rows are zero
    tdt = 2 * p.dto
    H = np.array(cfg.hoc[:nl])
    entfac = 0.5 * p.dxo * p.fnot ** 2
    A = {k: np.array(s0[k][:nl]) for k in ("enisoc", "eninoc", "ajisoc", "ajinoc", "ap3soc", "ap3noc", "ap5soc", "ap5noc",
                                           "ocncs", "ocncn", "ocncsp", "ocncnp")}
    en_s = np.concatenate([[0.0], A["enisoc"][: nl - 1], [0.0]])     # enisoc(0) = enisoc(nlo) = 0 in the layer differences
    en_n = np.concatenate([[0.0], A["eninoc"][: nl - 1], [0.0]])
    rhss = entfac / H * (en_s[1:] - en_s[:-1]) + A["ajisoc"] - A["ap3soc"] + A["ap5soc"]
    rhsn = entfac / H * (en_n[1:] - en_n[:-1]) + A["ajinoc"] + A["ap3noc"] - A["ap5noc"]
    rhss[0] += p.fnot / H[0] * s0["txisoc"]
    rhsn[0] -= p.fnot / H[0] * s0["txinoc"]
    rhss[-1] += p.fnot / H[-1] * s0["bdrins"]
    rhsn[-1] -= p.fnot / H[-1] * s0["bdrinn"]
    ocs_new, ocn_new = A["ocncsp"] + tdt * rhss, A["ocncnp"] + tdt * rhsn
    ayis = np.array([wts @ pm[:, 1, mo] for mo in range(nl)])         # dx/dy = 1
    ayin = np.array([-(wts @ pm[:, -2, mo]) for mo in range(nl)])
    clhss = l2m.T @ ocs_new + ayis
    clhsn = l2m.T @ ocn_new - ayin
    c3 = clhss[0] * s0["hbsioc"]
    hc1s, hc2s, hc1n, hc2n = (np.array(s0[k][: nl - 1]) for k in ("hc1soc", "hc2soc", "hc1noc", "hc2noc"))
    c1 = hc2n * clhss[1:] - hc2s * clhsn[1:]
    c2 = hc1s * clhsn[1:] - hc1n * clhss[1:]
    aipmod = np.concatenate([[xin[0] + c3 * s0["aipbho"]], xin[1:] + (c1 + c2) * np.array(s0["aipcho"][: nl - 1])])
    aiplay = m2l.T @ aipmod
    pmode = pm.copy()
    pmode[:, :, 0] += c3 * pbh[None, :]
    for mo in range(1, nl):
        pmode[:, :, mo] += (c1[mo - 1] * pch1[:, mo - 1] + c2[mo - 1] * pch2[:, mo - 1])[None, :]
    want = np.einsum("mk,ijm->ijk", m2l, pmode)
    m.ocinvq()
    s1 = m.get_scalars().as_dict()
    assert rel_l2(m.get_field("po", sh), want) <= 1e-12
    assert np.array_equal(m.get_field("pom", sh), po_old)
    assert np.allclose(s1["ocncs"][:nl], ocs_new, rtol=1e-12, atol=0.0) and np.allclose(s1["ocncn"][:nl], ocn_new, rtol=1e-12, atol=0.0)
    assert s1["ocncsp"][:nl] == list(A["ocncs"]) and s1["ocncnp"][:nl] == list(A["ocncn"])
    scale = np.abs(want).sum() * p.dxo ** 2
    assert np.abs(np.array(s1["dpioc"][: nl - 1]) - (aiplay[1:] - aiplay[:-1])).max() <= 1e-12 * scale
    assert s1["dpiocp"][: nl - 1] == s0["dpioc"][: nl - 1]
    est2 = np.array(s0["dpiocp"][: nl - 1]) - tdt * np.array(cfg.gpoc[: nl - 1]) * np.array(s0["xon"][: nl - 1])
    assert np.abs(np.array(s1["ermaso"][: nl - 1]) - ((aiplay[1:] - aiplay[:-1]) - est2)).max() <= 1e-12 * scale


def test_atinvq_against_numpy(qg, pyorc):
    """atmosphere inversion (src/atisubs.F:60-293): topography enters layer 1, the constraint
    right-hand sides and dpiat carry the atmosphere's signs, no del-cubed terms"""
    p = qg.named_config("dg_coupled").scaled(6, 5, ndxr=16, name="cpl_dg")
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    dxa = p.ndxr * p.dxo
    nl, nxp, nyp, nxt = p.nla, p.nxta + 1, p.nyta + 1, p.nxta
    sh = (nxp, nyp, nl)
    x = np.arange(nxp)[:, None] * 2 * np.pi / nxt
    ddyn = 1e-6 * np.cos(x) * np.sin(np.linspace(0, np.pi, nyp))[None, :]       # exercise the topography term
    m.set_field("ddynat", ddyn)
    m.run(1, 2)
    m.aml(); m.qgastep()
    qa, pa_old = m.get_field("qa", sh), m.get_field("pa", sh)
    pch1, pch2 = m.get_field("pch1at", (nyp, nl - 1)), m.get_field("pch2at", (nyp, nl - 1))
    pbh = m.get_field("pbhat")
    s0 = m.get_scalars().as_dict()
    l2m = np.array(cfg.ctl2mat[: nl * nl]).reshape(nl, nl, order="F")
    m2l = np.array(cfg.ctm2lat[: nl * nl]).reshape(nl, nl, order="F")
    yrel = np.arange(nyp) * dxa - 0.5 * p.nyta * dxa
    ql = qa - (p.beta * yrel)[None, :, None]
    ql[:, :, 0] -= ddyn
    wrk = np.zeros(sh)
    wrk[:, 1:-1, :] = p.fnot * np.einsum("km,ijk->ijm", l2m, ql[:, 1:-1, :])
    a = 1.0 / dxa ** 2
    kk = np.arange(nxt // 2 + 1)
    pm = np.zeros(sh)
    xin = np.zeros(nl)
    n = nyp - 2
    wts = np.ones(nxp); wts[0] = wts[-1] = 0.5
    for mo in range(nl):
        bk = -2 * a + 2 * a * (np.cos(kk * 2 * np.pi / nxt) - 1.0) - cfg.rdm2at[mo]
        spec = sf.rfft(wrk[:nxt, 1:-1, mo], axis=0)
        sol = np.empty_like(spec)
        for i in range(nxt // 2 + 1):
            ab = np.zeros((3, n)); ab[0, 1:] = a; ab[1, :] = bk[i]; ab[2, :-1] = a
            sol[i] = sla.solve_banded((1, 1), ab, spec[i].real) + 1j * sla.solve_banded((1, 1), ab, spec[i].imag)
        pm[:nxt, 1:-1, mo] = sf.irfft(sol, n=nxt, axis=0)
        pm[-1, :, mo] = pm[0, :, mo]
        xin[mo] = (wts @ pm[:, 1:-1, mo]).sum() * dxa ** 2
    tdt = 2 * p.dta
    H = np.array(cfg.hat[:nl])
    entfac = 0.5 * dxa * p.fnot ** 2
    A = {k: np.array(s0[k][:nl]) for k in ("enisat", "eninat", "ajisat", "ajinat", "ap5sat", "ap5nat", "atmcs", "atmcn", "atmcsp", "atmcnp")}
    en_s = np.concatenate([[0.0], A["enisat"][: nl - 1], [0.0]])
    en_n = np.concatenate([[0.0], A["eninat"][: nl - 1], [0.0]])
    rhss = -entfac / H * (en_s[1:] - en_s[:-1]) + A["ajisat"] + A["ap5sat"]
    rhsn = -entfac / H * (en_n[1:] - en_n[:-1]) + A["ajinat"] - A["ap5nat"]
    rhss[0] -= p.fnot / H[0] * s0["txisat"]
    rhsn[0] += p.fnot / H[0] * s0["txinat"]
    ats_new, atn_new = A["atmcsp"] + tdt * rhss, A["atmcnp"] + tdt * rhsn
    ayis = np.array([wts @ pm[:, 1, mo] for mo in range(nl)])
    ayin = np.array([-(wts @ pm[:, -2, mo]) for mo in range(nl)])
    clhss = l2m.T @ ats_new + ayis
    clhsn = l2m.T @ atn_new - ayin
    c3 = clhss[0] * s0["hbsiat"]
    hc1s, hc2s, hc1n, hc2n = (np.array(s0[k][: nl - 1]) for k in ("hc1sat", "hc2sat", "hc1nat", "hc2nat"))
    c1 = hc2n * clhss[1:] - hc2s * clhsn[1:]
    c2 = hc1s * clhsn[1:] - hc1n * clhss[1:]
    aipmod = np.concatenate([[xin[0] + c3 * s0["aipbha"]], xin[1:] + (c1 + c2) * np.array(s0["aipcha"][: nl - 1])])
    aiplay = m2l.T @ aipmod
    pmode = pm.copy()
    pmode[:, :, 0] += c3 * pbh[None, :]
    for mo in range(1, nl):
        pmode[:, :, mo] += (c1[mo - 1] * pch1[:, mo - 1] + c2[mo - 1] * pch2[:, mo - 1])[None, :]
    want = np.einsum("mk,ijm->ijk", m2l, pmode)
    m.atinvq()
    s1 = m.get_scalars().as_dict()
    assert rel_l2(m.get_field("pa", sh), want) <= 1e-12
    assert np.array_equal(m.get_field("pam", sh), pa_old)
    assert np.allclose(s1["atmcs"][:nl], ats_new, rtol=1e-12, atol=0.0) and np.allclose(s1["atmcn"][:nl], atn_new, rtol=1e-12, atol=0.0)
    scale = np.abs(want).sum() * dxa ** 2
    assert np.abs(np.array(s1["dpiat"][: nl - 1]) - (aiplay[:-1] - aiplay[1:])).max() <= 1e-12 * scale       # sign: layer k minus k+1
    assert s1["dpiatp"][: nl - 1] == s0["dpiat"][: nl - 1]


def test_homsol_box_against_numpy(qg, pyorc):
    """homsol, finite box (src/conhoms.F:549-640): ochom_m = 1 + rdm2 * sol0 with
    (del-sqd - rdm2) sol0 = 1, sol0 = 0 on the walls; aipohs, cdiffo, cdhoc from it"""
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")        # ends with homsol
    nl, nxp, nyp, nxt = p.nlo, p.nxpo, p.nypo, p.nxto
    ochom = m.get_field("ochom", (nxp, nyp, nl - 1))
    s = m.get_scalars().as_dict()
    m2l = np.array(cfg.ctm2loc[: nl * nl]).reshape(nl, nl, order="F")
    a = 1.0 / p.dxo ** 2
    kk = np.arange(1, nxt)
    n = nyp - 2
    w = np.ones(nxp); w[0] = w[-1] = 0.5
    wy = np.ones(nyp); wy[0] = wy[-1] = 0.5
    aip = np.zeros(nl - 1)
    for mo in range(nl - 1):
        rd = cfg.rdm2oc[mo + 1]
        b = -2 * a + 2 * a * (np.cos(kk * np.pi / nxt) - 1.0) - rd
        spec = sf.dst(np.ones((nxt - 1, n)), type=1, axis=0)
        sol = np.empty_like(spec)
        for i in range(nxt - 1):
            ab = np.zeros((3, n)); ab[0, 1:] = a; ab[1, :] = b[i]; ab[2, :-1] = a
            sol[i] = sla.solve_banded((1, 1), ab, spec[i])
        want = np.ones((nxp, nyp))
        want[1:-1, 1:-1] = 1.0 + rd * sf.dst(sol * (0.5 / nxt), type=1, axis=0)
        assert rel_l2(ochom[:, :, mo], want) <= 1e-12, mo
        # the homogeneous problem itself: (del-sqd - rdm2) ochom = 0 inside, ochom = 1 on the walls
        h = ochom[:, :, mo]
        res = lap5(h, a) - rd * h[1:-1, 1:-1]
        assert np.abs(res).max() <= 1e-9 * rd
        aip[mo] = (w @ h @ wy) * p.dxo ** 2
        assert np.isclose(s["aipohs"][mo], aip[mo], rtol=1e-12)
    cdiffo = np.array(s["cdiffo"][: nl * (nl - 1)]).reshape(nl, nl - 1, order="F")
    cdhoc = np.array(s["cdhoc"][: (nl - 1) ** 2]).reshape(nl - 1, nl - 1, order="F")
    for k in range(nl - 1):
        assert np.allclose(cdiffo[:, k], m2l[:, k + 1] - m2l[:, k], rtol=1e-14, atol=0.0)
        assert np.allclose(cdhoc[k, :], (m2l[1:, k + 1] - m2l[1:, k]) * aip, rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("case", ["box_dg", "chan_so"])
def test_constr_matches_numpy(qg, pyorc, case):
    """constr (src/conhoms.F:44-200): area integrals of the interface pressure differences on
    both time levels; in a channel the zonal line integrals of p_y and of A p"""
    p = small_configs(qg)[case]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.ocean_step()
    m.constr()
    s = m.get_scalars().as_dict()
    nl, nxp, nyp = p.nlo, p.nxpo, p.nypo
    po, pom = m.get_field("po", (nxp, nyp, nl)), m.get_field("pom", (nxp, nyp, nl))
    w = np.ones(nxp); w[0] = w[-1] = 0.5
    wy = np.ones(nyp); wy[0] = wy[-1] = 0.5
    dA = p.dxo ** 2
    for k in range(nl - 1):
        for name, f in (("dpioc", po), ("dpiocp", pom)):
            d = f[:, :, k + 1] - f[:, :, k]
            want, scale = (w @ d @ wy) * dA, (w @ np.abs(d) @ wy) * dA
            assert abs(s[name][k] - want) <= 1e-12 * scale, (name, k)
    if p.has("cyclic_ocean"):
        A = _amat(cfg.amatoc, nl)
        half = 0.5 * p.dxo * p.fnot ** 2
        # the synthetic channel modes have vanishing zonal means: compare against the integrals of the magnitudes
        def scale(f, j0, j1):
            return (w @ np.abs(f[:, j1, :] - f[:, j0, :])).max() + half * (np.abs(A) @ (p.dxo * (w @ np.abs(f[:, j0, :])))).max()
        for name, f in (("ocncs", po), ("ocncsp", pom)):
            line = -(w @ (f[:, 1, :] - f[:, 0, :])) + half * (A @ (p.dxo * (w @ f[:, 0, :])))
            assert np.abs(np.array(s[name][:nl]) - line).max() <= 1e-12 * scale(f, 0, 1), name
        for name, f in (("ocncn", po), ("ocncnp", pom)):
            line = (w @ (f[:, -1, :] - f[:, -2, :])) + half * (A @ (p.dxo * (w @ f[:, -1, :])))
            assert np.abs(np.array(s[name][:nl]) - line).max() <= 1e-12 * scale(f, -1, -2), name


def _oml_ghost_cells(qg, pyorc, p):
    """the whole of oml/omladf (src/omlsubs.F:47-763) as a ghost-cell scheme: every boundary
    variant of the reference (no-flux walls, periodic channel, Ekman outflow at tsbdy / tnbdy,
    the four corners) is 'flux through a face = face velocity x sum of the two cell values'
    and 'Laplacian with a ghost value outside', with the ghost chosen per boundary"""
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.ocean_step()                                  # so that sst and sstm differ
    nxt, nyt, nxp, nyp = p.nxto, p.nyto, p.nxpo, p.nypo
    po = m.get_field("po", (nxp, nyp, p.nlo))[:, :, 0]
    tx, ty = m.get_field("tauxo", (nxp, nyp)), m.get_field("tauyo", (nxp, nyp))
    sst, sstm = m.get_field("sst", (nxt, nyt)), m.get_field("sstm", (nxt, nyt))
    wek, fnet = m.get_field("wekto", (nxt, nyt)), m.get_field("fnetoc", (nxt, nyt))
    m.oml()
    got_sst, got_ent = m.get_field("sst", (nxt, nyt)), m.get_field("entoc", (nxp, nyp))
    s = m.get_scalars().as_dict()
    cyc, sb, nb = p.has("cyclic_ocean"), p.has("sb_hflux"), p.has("nb_hflux")
    uvg, rh, hdx = p.ycexp / (p.dxo * p.fnot), 0.5 / (p.fnot * p.hmoc), 0.5 / p.dxo
    # face velocities: u on the nxp x faces of every T row, v on the nyp y faces of every T column
    u = -uvg * (po[:, 1:] - po[:, :-1]) + rh * (ty[:, 1:] + ty[:, :-1])
    v = uvg * (po[1:, :] - po[:-1, :]) - rh * (tx[1:, :] + tx[:-1, :])
    if not cyc:
        u[0] = u[-1] = 0.0
    v[:, 0] = -rh * (tx[1:, 0] + tx[:-1, 0]) if sb else 0.0
    v[:, -1] = -rh * (tx[1:, -1] + tx[:-1, -1]) if nb else 0.0

    def pad(f, south, north, xmode):
        """one ghost ring: x periodic or copy (no flux); y copy or a prescribed temperature"""
        g = np.empty((f.shape[0] + 2, f.shape[1] + 2))
        g[1:-1, 1:-1] = f
        g[1:-1, 0] = f[:, 0] if south is None else south
        g[1:-1, -1] = f[:, -1] if north is None else north
        if xmode == "periodic":
            g[0], g[-1] = g[-2], g[1]
        else:
            g[0], g[-1] = g[1], g[-2]
        return g

    xm = "periodic" if cyc else "copy"
    T = pad(sst, cfg.tsbdy if sb else None, cfg.tnbdy if nb else None, xm)        # advected temperature
    hx = hdx * (u[1:] * (T[1:-1, 1:-1] + T[2:, 1:-1]) - u[:-1] * (T[:-2, 1:-1] + T[1:-1, 1:-1]))
    hy = hdx * (v[:, 1:] * (T[1:-1, 1:-1] + T[1:-1, 2:]) - v[:, :-1] * (T[1:-1, :-2] + T[1:-1, 1:-1]))
    Tm = pad(sstm, cfg.tsbdy if sb else None, cfg.tnbdy if nb else None, xm)
    lap = lambda g: g[1:-1, :-2] + g[:-2, 1:-1] + g[2:, 1:-1] + g[1:-1, 2:] - 4.0 * g[1:-1, 1:-1]
    d2 = lap(Tm)
    d4 = lap(pad(d2, None, None, xm))                # no diffusive flux of del-sqd T through any wall
    dxm2 = 1.0 / p.dxo ** 2
    rhs = -(hx + hy) + p.st2d * dxm2 * d2 - p.st4d * dxm2 ** 2 * d4
    toc1, toc2 = cfg.toc[0], cfg.toc[1]
    tdt = 2.0 * p.dto
    rrcp = 1.0 / (p.rhooc * p.cpoc)
    new = sstm + tdt * (rhs + (rrcp * fnet + 0.5 * wek * (sstm + toc1)) / p.hmoc)
    dtonew = toc1 - new
    dtoinv = 1.0 / (toc1 - toc2)
    xfo = -(0.5 * dtoinv) * wek * (sstm - toc1) - (p.hmoc * dtoinv / tdt) * np.maximum(0.0, dtonew)
    new = new + np.maximum(0.0, dtonew)
    xfo = xfo - xfo.sum() / (nxt * nyt)
    X = pad(xfo, None, None, xm)                     # edge and corner rules = averaging with copied ghosts
    ent = 0.25 * (X[:-1, :-1] + X[1:, :-1] + X[:-1, 1:] + X[1:, 1:])
    return (got_sst, new), (got_ent, ent), s, m, p


@pytest.mark.parametrize("case", ["box_dg", "box_natl1km", "chan_so", "box_plain"])
def test_oml_whole_domain_against_ghost_cell_scheme(qg, pyorc, case):
    if case == "box_plain":
        p = small_configs(qg)["box_dg"]
        p.flags = ["ocean_only"]                     # neither sb_hflux nor nb_hflux
    else:
        p = small_configs(qg)[case]
    (got_sst, want_sst), (got_ent, want_ent), s, m, p = _oml_ghost_cells(qg, pyorc, p)
    assert rel_l2(got_sst, want_sst) <= 1e-13, case
    scale = np.abs(want_ent).max()
    assert np.abs(got_ent - want_ent).max() <= 1e-12 * scale, case
    w = np.ones(p.nxpo); w[0] = w[-1] = 0.5
    wy = np.ones(p.nypo); wy[0] = wy[-1] = 0.5
    assert abs(s["xon"][0]) <= 1e-10 * (w @ np.abs(want_ent) @ wy) * p.dxo ** 2      # zero net entrainment


@pytest.mark.parametrize("deck", ["dg_coupled", "so_coupled"])
def test_aml_whole_domain_against_ghost_cell_scheme(qg, pyorc, deck):
    """aml/amladf (src/amlsubs.F:47-563) on the whole atmosphere grid: periodic in x; on the zonal
    walls no heat flux (temperature ghost = inside value, no advective flux), while the
    thickness sees hmat outside and is carried through the wall by vekat"""
    if deck == "dg_coupled":
        p = qg.named_config(deck).scaled(6, 5, ndxr=16, name="cpl_dg")
    else:
        p = qg.named_config(deck).scaled(12, 3, nxta=12, nyta=9, ndxr=16, name="cpl_so")
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.run(1, 2)
    nxt, nyt, nxp, nyp, nl = p.nxta, p.nyta, p.nxta + 1, p.nyta + 1, p.nla
    g = lambda n, sh: m.get_field(n, sh)
    ast, astm, hm, hmm = g("ast", (nxt, nyt)), g("astm", (nxt, nyt)), g("hmixa", (nxt, nyt)), g("hmixam", (nxt, nyt))
    fnet, wekta, xc1 = g("fnetat", (nxt, nyt)), g("wekta", (nxt, nyt)), g("xc1ast", (nxt, nyt))
    uek, vek = g("uekat", (nxp, nyt)), g("vekat", (nxt, nyp))
    pa, pam, dtop = g("pa", (nxp, nyp, nl)), g("pam", (nxp, nyp, nl)), g("dtopat", (nxp, nyp))
    m.aml()
    ast_n, hm_n, ent_n = g("ast", (nxt, nyt)), g("hmixa", (nxt, nyt)), g("entat", (nxp, nyp))
    dxa = p.ndxr * p.dxo
    rdxf0, hdx, dxm2 = 1.0 / (dxa * p.fnot), 0.5 / dxa, 1.0 / dxa ** 2
    tdt = 2.0 * p.dta
    rrcpat = 1.0 / (cfg.rhoat * cfg.cpat)
    p1 = pa[:, :, 0]
    u = -rdxf0 * (p1[:, 1:] - p1[:, :-1]) + uek
    v = rdxf0 * (p1[1:, :] - p1[:-1, :]) + vek
    v[:, 0], v[:, -1] = vek[:, 0], vek[:, -1]            # p is uniform along the zonal walls

    def pad(f, south, north):
        gg = np.empty((f.shape[0] + 2, f.shape[1] + 2))
        gg[1:-1, 1:-1] = f
        gg[1:-1, 0] = f[:, 0] if south is None else south
        gg[1:-1, -1] = f[:, -1] if north is None else north
        gg[0], gg[-1] = gg[-2], gg[1]
        return gg

    lap = lambda gg: gg[1:-1, :-2] + gg[:-2, 1:-1] + gg[2:, 1:-1] + gg[1:-1, 2:] - 4.0 * gg[1:-1, 1:-1]

    def flux_div(gg, vv):
        x = hdx * (u[1:] * (gg[1:-1, 1:-1] + gg[2:, 1:-1]) - u[:-1] * (gg[:-2, 1:-1] + gg[1:-1, 1:-1]))
        y = hdx * (vv[:, 1:] * (gg[1:-1, 1:-1] + gg[1:-1, 2:]) - vv[:, :-1] * (gg[1:-1, :-2] + gg[1:-1, 1:-1]))
        return x + y

    vT = v.copy(); vT[:, 0] = vT[:, -1] = 0.0            # no heat flux through the walls
    d2 = lap(pad(astm, None, None))
    tmrhs = -flux_div(pad(ast, None, None), vT) + p.at2d * dxm2 * d2 - p.at4d * dxm2 ** 2 * lap(pad(d2, None, None))
    hmrhs = -flux_div(pad(hm, p.hmat, p.hmat), v) + p.ahmd * dxm2 * lap(pad(hmm, p.hmat, p.hmat))
    tat1, tat2 = cfg.tat[0], cfg.tat[1]
    hdrcdt = p.hmadmp * rrcpat * tdt
    cold = astm <= tat1 - 2.0 * hdrcdt
    with np.errstate(divide="ignore", invalid="ignore"):
        hnew = hmm + tdt * hmrhs - hdrcdt * (hmm - p.hmat) / (tat1 - astm)
    dhfix = np.maximum(p.hmamin - hnew, 0.0)
    hnew = np.where(cold, hnew + dhfix, p.hmat)
    dtfix = np.where(cold, dhfix * (tat1 - astm) / hmm, 0.0)
    astnew = astm + tdt * (tmrhs + rrcpat * fnet / hmm - wekta * astm / p.hmat) + dtfix
    dtanew = tat1 - astnew
    conena = hm * np.minimum(0.0, dtanew) / (tdt * (tat2 - tat1))
    xfa = p.xcexp * cfg.bface * (hmm - p.hmat) + cfg.dface * (p.xcexp * astm + xc1) - p.xcexp * conena
    astnew = astnew + np.minimum(0.0, dtanew)
    assert rel_l2(ast_n, astnew) <= 1e-13
    assert rel_l2(hm_n, hnew) <= 1e-13
    X = pad(xfa, None, None)
    ent = 0.25 * (X[:-1, :-1] + X[1:, :-1] + X[:-1, 1:] + X[1:, 1:])
    ent = ent + sum(cfg.aface[l] / cfg.gpat[l] * (pam[:, :, l] - pam[:, :, l + 1]) for l in range(nl - 1)) + cfg.cface * dtop
    assert np.abs(ent_n - ent).max() <= 1e-12 * np.abs(ent).max()


def test_qgostep_box_walls_and_time_levels(qg, pyorc):
    """the parts of qgostep the interior check leaves out (src/qgosubs.F:184-219, SURVEY.md
    quirk 5): the W/E wall columns are stepped with the forcing alone (dqdt = 0 there,
    :371, :397), every interior row of qom takes the old qo, the zonal boundary rows of qo stay
    and qom copies them"""
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.ocean_step()
    m.oml()
    sh = (p.nxpo, p.nypo, p.nlo)
    pom, qo, qom = (m.get_field(n, sh) for n in ("pom", "qo", "qom"))
    wek, ent = m.get_field("wekpo", sh[:2]), m.get_field("entoc", sh[:2])
    m.qgostep()
    qnew, qmnew = m.get_field("qo", sh), m.get_field("qom", sh)
    dxm2 = 1.0 / p.dxo ** 2
    bcf = p.bccooc * dxm2 / (0.5 * p.bccooc + 1.0)
    tdt = 2.0 * p.dto
    rows = slice(1, -1)
    for wall, inner in ((0, 1), (-1, -2)):
        d2w = bcf * (pom[inner, rows, -1] - pom[wall, rows, -1])        # del-sqd of the bottom layer on the wall
        for k in range(p.nlo):
            f = np.zeros(p.nypo - 2)
            if k == 0:
                f = f + (p.fnot / p.hoc[0]) * (wek[wall, rows] - ent[wall, rows])
            if k == 1:
                f = f + (p.fnot / p.hoc[1]) * ent[wall, rows]
            if k == p.nlo - 1:
                f = f - 0.5 * np.sign(p.fnot) * p.delek / p.hoc[-1] * d2w
            assert rel_l2(qnew[wall, rows, k], qom[wall, rows, k] + tdt * f) <= 1e-14, (wall, k)
    assert np.array_equal(qmnew[:, rows, :], qo[:, rows, :])
    for j in (0, -1):
        assert np.array_equal(qnew[:, j, :], qo[:, j, :]) and np.array_equal(qmnew[:, j, :], qo[:, j, :])


def test_qgostep_channel_boundary_integrals(qg, pyorc):
    """the zonal-boundary sums a channel feeds into its momentum constraints (src/qgosubs.F:
    155-162 bdrins/bdrinn, :284-296 and :409-423 Jacobian strips, :429-443 third and fifth
    derivative strips)"""
    p = small_configs(qg)["chan_so"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.ocean_step()
    m.oml()
    nl, nxt = p.nlo, p.nxto
    sh = (p.nxpo, p.nypo, nl)
    po, pom, qo = (m.get_field(n, sh)[:-1] for n in ("po", "pom", "qo"))       # the nxto distinct columns
    m.qgostep()
    s = m.get_scalars().as_dict()
    dxm2 = 1.0 / p.dxo ** 2
    bcf = p.bccooc * dxm2 / (0.5 * p.bccooc + 1.0)
    adf = 1.0 / (12.0 * p.dxo ** 2 * p.fnot)

    def lap(f):
        out = np.empty_like(f)
        out[:, 1:-1] = (f[:, :-2] + np.roll(f, 1, axis=0)[:, 1:-1] + np.roll(f, -1, axis=0)[:, 1:-1] + f[:, 2:] - 4.0 * f[:, 1:-1]) * dxm2
        out[:, 0] = bcf * (f[:, 1] - f[:, 0])
        out[:, -1] = bcf * (f[:, -2] - f[:, -1])
        return out

    dpx = lambda f, j: np.roll(f[:, j], -1) - np.roll(f[:, j], 1)           # p(i+1,j) - p(i-1,j), periodic
    for k in range(nl):
        P, Q = po[:, :, k], qo[:, :, k]
        d2 = lap(pom[:, :, k]); d4 = lap(d2)
        ajis = p.dxo ** 2 * p.fnot * adf * ((Q[:, 0] * dpx(P, 1)).sum() + 2.0 * (Q[:, 1] * dpx(P, 1)).sum())
        ajin = -p.dxo ** 2 * p.fnot * adf * ((Q[:, -1] * dpx(P, -2)).sum() + 2.0 * (Q[:, -2] * dpx(P, -2)).sum())
        sc = p.dxo ** 2 * abs(p.fnot * adf) * 3.0 * (np.abs(Q[:, 0]) * np.abs(dpx(P, 1))).sum()
        assert abs(s["ajisoc"][k] - ajis) <= 1e-11 * sc and abs(s["ajinoc"][k] - ajin) <= 1e-11 * sc, k
        for name, fld, coef, rows in (("ap3soc", d2, p.ah2oc[k], (1, 0)), ("ap3noc", d2, p.ah2oc[k], (-1, -2)),
                                      ("ap5soc", d4, p.ah4oc[k], (1, 0)), ("ap5noc", d4, p.ah4oc[k], (-1, -2))):
            diff = fld[:, rows[0]] - fld[:, rows[1]]
            assert abs(s[name][k] - coef * diff.sum()) <= 1e-11 * max(coef, 1e-300) * np.abs(diff).sum(), (name, k)
    bd = 0.5 * np.sign(p.fnot) * p.delek
    ds, dn = pom[:, 1, -1] - pom[:, 0, -1], pom[:, -1, -1] - pom[:, -2, -1]
    assert abs(s["bdrins"] - bd * ds.sum()) <= 1e-12 * abs(bd) * np.abs(ds).sum()
    assert abs(s["bdrinn"] - bd * dn.sum()) <= 1e-12 * abs(bd) * np.abs(dn).sum()


def _channel_homsol(ny, dx, rdm2, xl, yl):
    """1-D version of the channel's homogeneous baroclinic solutions (src/conhoms.F:400-543):
    pch = L(y) + rdm2 * sol0 with sol0'' - rdm2 sol0 = L(y), sol0 = 0 on both walls"""
    a = 1.0 / dx ** 2
    y = np.arange(ny) * dx
    out = []
    for L in ((y[-1] - y) / yl, (y - y[0]) / yl):
        n = ny - 2
        ab = np.zeros((3, n)); ab[0, 1:] = a; ab[1, :] = -2.0 * a - rdm2; ab[2, :-1] = a
        sol = np.zeros(ny)
        sol[1:-1] = sla.solve_banded((1, 1), ab, L[1:-1])
        out.append(L + rdm2 * sol)
    p1, p2 = out
    w = np.ones(ny); w[0] = w[-1] = 0.5
    aip = 0.5 * ((w @ p1) + (w @ p2)) * xl * dx
    ys = lambda q: xl * (-(q[1] - q[0]) / dx + 0.5 * dx * rdm2 * q[0])
    yn = lambda q: xl * ((q[-1] - q[-2]) / dx + 0.5 * dx * rdm2 * q[-1])
    det = ys(p1) * yn(p2) - ys(p2) * yn(p1)
    return p1, p2, aip, ys(p1) / det, ys(p2) / det, yn(p1) / det, yn(p2) / det


def test_homsol_channel_against_numpy(qg, pyorc):
    p = small_configs(qg)["chan_so"]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    nl, nyp = p.nlo, p.nypo
    s = m.get_scalars().as_dict()
    pch1, pch2 = m.get_field("pch1oc", (nyp, nl - 1)), m.get_field("pch2oc", (nyp, nl - 1))
    xl, yl = p.nxto * p.dxo, p.nyto * p.dxo
    assert np.allclose(m.get_field("pbhoc"), (nyp - 1 - np.arange(nyp)) / (nyp - 1), rtol=1e-15, atol=0.0)
    assert np.isclose(s["hbsioc"], yl / xl, rtol=1e-15) and np.isclose(s["aipbho"], 0.5 * xl * yl, rtol=1e-15)
    for mo in range(nl - 1):
        p1, p2, aip, h1s, h2s, h1n, h2n = _channel_homsol(nyp, p.dxo, cfg.rdm2oc[mo + 1], xl, yl)
        assert rel_l2(pch1[:, mo], p1) <= 1e-12 and rel_l2(pch2[:, mo], p2) <= 1e-12, mo
        assert np.isclose(s["aipcho"][mo], aip, rtol=1e-12)
        for name, v in (("hc1soc", h1s), ("hc2soc", h2s), ("hc1noc", h1n), ("hc2noc", h2n)):
            assert np.isclose(s[name][mo], v, rtol=1e-10), (name, mo)


def test_homsol_atmosphere_against_numpy(qg, pyorc):
    """the atmosphere's twin (src/conhoms.F:650-818)"""
    p = qg.named_config("dg_coupled").scaled(6, 5, ndxr=16, name="cpl_dg")
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    nl, nyp = p.nla, p.nyta + 1
    dxa = p.ndxr * p.dxo
    s = m.get_scalars().as_dict()
    pch1, pch2 = m.get_field("pch1at", (nyp, nl - 1)), m.get_field("pch2at", (nyp, nl - 1))
    xl, yl = p.nxta * dxa, p.nyta * dxa
    for mo in range(nl - 1):
        p1, p2, aip, h1s, h2s, h1n, h2n = _channel_homsol(nyp, dxa, cfg.rdm2at[mo + 1], xl, yl)
        assert rel_l2(pch1[:, mo], p1) <= 1e-12 and rel_l2(pch2[:, mo], p2) <= 1e-12, mo
        assert np.isclose(s["aipcha"][mo], aip, rtol=1e-12)
        for name, v in (("hc1sat", h1s), ("hc2sat", h2s), ("hc1nat", h1n), ("hc2nat", h2n)):
            assert np.isclose(s[name][mo], v, rtol=1e-10), (name, mo)

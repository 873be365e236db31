"""The reference ships no golden vectors, so the oracle is pinned by
(1) the analytic identities the reference itself prints or relies on, and (2) an
independent numpy/scipy re-derivation of the same Fortran (test_numpy_crosscheck.py)."""
import numpy as np
import pytest

from util import small_configs


def _bd2(p, cyc):
    nxt, a = p.nxto, 1.0 / p.dxo ** 2
    bd2 = np.zeros(nxt)
    if cyc:
        for i in range(2, nxt // 2 + 1):
            bd2[2 * i - 3] = -2 * a + 2 * a * (np.cos((i - 1) * 2 * np.pi / nxt) - 1.0)
            bd2[2 * i - 2] = bd2[2 * i - 3]
        bd2[0] = -2 * a
        bd2[nxt - 1] = -6 * a
    else:
        bd2[: nxt - 1] = -2 * a + 2 * a * (np.cos(np.arange(1, nxt) * np.pi / nxt) - 1.0)
    return bd2


@pytest.fixture(scope="module")
def models(qg, pyorc):
    out = {}
    for name, p in small_configs(qg).items():
        cfg = qg.build_config(p)
        m = pyorc.Oracle(cfg)
        qg.synth.init_model(m, p, cfg, "random")
        out[name] = (p, cfg, m)
    return out


def test_mode_matrices_are_inverse(qg):
    """eigmod prints cm2l*cl2m 'should be the identity matrix' (src/eigmode.f:430-438)"""
    cfg = qg.build_config(qg.named_config("dg_oo").scaled(6, 5))
    for l2m, m2l, n in ((cfg.ctl2moc, cfg.ctm2loc, cfg.nlo), (cfg.ctl2mat, cfg.ctm2lat, cfg.nla)):
        a = np.array(l2m[: n * n]).reshape(n, n, order="F")   # ctl2m(k,m)
        b = np.array(m2l[: n * n]).reshape(n, n, order="F")   # ctm2l(m,k)
        assert np.allclose(b.T @ a.T, np.eye(n), atol=1e-13)
    assert cfg.rdm2oc[0] == 0.0 and cfg.rdm2oc[1] > 0.0 and cfg.rdm2oc[2] > cfg.rdm2oc[1]


@pytest.mark.parametrize("case", ["box_dg", "chan_so"])
def test_helmholtz_residual(models, case):
    """apply the 5-point modified-Helmholtz operator to the solver output"""
    p, cfg, m = models[case]
    cyc = p.has("cyclic_ocean")
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal((p.nxpo, p.nypo))
    if cyc:
        rhs[-1] = rhs[0]
    a = 1.0 / p.dxo ** 2
    for mode in (1, 2):
        rd = cfg.rdm2oc[mode]
        s = m.helmholtz(0, rhs, _bd2(p, cyc) - rd)
        assert np.all(s[:, 0] == 0) and np.all(s[:, -1] == 0)
        if cyc:
            assert np.array_equal(s[-1], s[0])
            ext = np.vstack([s[-2:-1], s, s[1:2]])
            lap = (ext[2:, 1:-1] + ext[:-2, 1:-1] + s[:, 2:] + s[:, :-2] - 4 * s[:, 1:-1]) * a - rd * s[:, 1:-1]
            want = rhs[:, 1:-1]
        else:
            assert np.all(s[0] == 0) and np.all(s[-1] == 0)
            lap = (s[2:, 1:-1] + s[:-2, 1:-1] + s[1:-1, 2:] + s[1:-1, :-2] - 4 * s[1:-1, 1:-1]) * a - rd * s[1:-1, 1:-1]
            want = rhs[1:-1, 1:-1]
        assert np.linalg.norm(lap - want) <= 1e-11 * np.linalg.norm(want)


@pytest.mark.parametrize("case", ["box_dg", "box_natl1km", "chan_so"])
def test_entrainment_has_zero_mean(models, case):
    """xon(1) is zero by construction of entoc (src/ocisubs.F:334, src/omlsubs.F:146-155)"""
    p, cfg, m = models[case]
    m.oml()
    s = m.get_scalars()
    ent = m.get_field("entoc")
    scale = np.abs(ent).sum() * p.dxo ** 2
    assert abs(s.xon[0]) <= 1e-12 * max(scale, 1.0)


def test_box_mass_constraint_holds(models):
    """after ocinvq the area integral of p(k+1)-p(k) equals the stepped dpioc (the
    constraint the homogeneous solutions enforce, src/ocisubs.F:333-370)"""
    p, cfg, m = models["box_dg"]
    import pyorc
    m.run(1, 7)
    s = m.get_scalars()
    po = m.get_field("po", (p.nxpo, p.nypo, p.nlo))
    for k in range(p.nlo - 1):
        integ = pyorc.xintp(po[:, :, k + 1] - po[:, :, k]) * p.dxo ** 2
        assert abs(integ - s.dpioc[k]) <= 1e-9 * abs(s.dpioc[k])


def test_channel_mass_continuity_monitor(models):
    """ermaso ~ 0: the two estimates of dpioc agree (src/ocisubs.F:268-284)"""
    p, cfg, m = models["chan_so"]
    m.run(1, 9)
    s = m.get_scalars()
    for k in range(p.nlo - 1):
        assert abs(s.ermaso[k]) <= 1e-6 * max(abs(s.dpioc[k]), p.nxto * p.nyto * p.dxo ** 2 * 1e-6)


def test_atqzbd_quirk(qg, pyorc):
    """src/vorsubs.F:470: southern row of the top layer uses pa(i,2,nla)"""
    p = qg.named_config("dg_coupled").scaled(6, 5, nxta=12, nyta=8)
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    rng = np.random.default_rng(1)
    pa = rng.standard_normal((p.nxpa, p.nypa, p.nla))
    m.set_field("pa", pa)
    m.atqzbd()
    qa = m.get_field("qa", pa.shape)
    nla = p.nla
    A = np.array(cfg.amatat[: nla * nla]).reshape(nla, nla, order="F")
    dxa = p.dxa
    zb = p.bccoat / dxa ** 2 / (0.5 * p.bccoat + 1.0) / p.fnot
    yla = p.nyta * dxa
    betays = p.beta * (0.0 - 0.5 * yla)
    want = zb * (pa[:, 1, -1] - pa[:, 0, -1]) - (p.fnot * A[-1, -2] * pa[:, 0, -2] + p.fnot * A[-1, -1] * pa[:, 1, -1]) + betays
    assert np.allclose(qa[:, 0, -1], want, rtol=1e-13, atol=0)


def test_modes_diagonalise_the_stretching_operator(qg):
    """what the mode matrices are for (src/eigmode.f:41-440, src/ocisubs.F:117-139): with
    q_k = del-sqd p_k - f0^2 (A p)_k, projecting layers onto modes turns f0^2 A into
    diag(rdm2), i.e. ctl2m^T (f0^2 A) ctm2l^T = diag(rdm2); the barotropic mode is uniform in
    the vertical, and the ocean's modes carry Flierl's normalisation sum_k H_k phi_k^2 = H"""
    p = qg.named_config("dg_coupled")
    cfg = qg.build_config(p)
    for amat, l2m, m2l, rd, n, h, ocean in ((cfg.amatoc, cfg.ctl2moc, cfg.ctm2loc, cfg.rdm2oc, cfg.nlo, cfg.hoc, True),
                                            (cfg.amatat, cfg.ctl2mat, cfg.ctm2lat, cfg.rdm2at, cfg.nla, cfg.hat, False)):
        A = np.array(amat[: n * n]).reshape(n, n, order="F")
        L = np.array(l2m[: n * n]).reshape(n, n, order="F")       # ctl2m(k,m)
        M = np.array(m2l[: n * n]).reshape(n, n, order="F")       # ctm2l(m,k)
        D = L.T @ (p.fnot ** 2 * A) @ M.T
        want = np.diag(np.array(rd[:n]))
        assert np.abs(D - want).max() <= 1e-10 * np.abs(want).max()
        assert np.abs(A.sum(axis=1)).max() <= 1e-15 * np.abs(A).max()   # A annihilates a depth-independent p
        assert np.allclose(M[0] / M[0, 0], 1.0, rtol=1e-12)       # so the barotropic mode is uniform
        if ocean:
            H = np.array(h[:n])
            for m in range(n):
                assert np.isclose((H * M[m] ** 2).sum(), H.sum(), rtol=1e-12) and M[m, 0] > 0.0

"""The oracle's restatement of qocdiag_out (src/qocdiag.F:303-683) pinned two ways: its dqdt is
the tendency the oracle's qgostep (a separate restatement, src/qgosubs.F) applies in the
leapfrog step, and its terms match a vectorised numpy evaluation in the interior."""
import numpy as np
import pytest

from util import small_configs, rel_l2
from test_numpy_crosscheck import lap5, jac9


def _setup(qg, pyorc, case):
    p = small_configs(qg)[case]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.ocean_step()
    m.oml()                    # qocdiag_out runs between oml and qgostep (src/q-gcm.F:1232-1243)
    return p, cfg, m


@pytest.mark.parametrize("case", ["box_dg", "chan_so"])
def test_dqdt_is_the_leapfrog_tendency(qg, pyorc, case):
    p, cfg, m = _setup(qg, pyorc, case)
    sh = (p.nxpo, p.nypo, p.nlo)
    d = m.qocdiag(1).reshape(sh + (5,), order="F")
    qom = m.get_field("qom", sh)
    m.qgostep()
    qnew = m.get_field("qo", sh)
    cols = slice(None) if p.has("cyclic_ocean") else slice(1, -1)
    tend = (qnew - qom)[cols, 1:-1, :] / (2.0 * p.dto)
    assert rel_l2(d[cols, 1:-1, :, 0], tend) <= 1e-9      # the difference of two close numbers
    # the four terms add up to dqdt wherever they are defined
    assert rel_l2(d[cols, 1:-1, :, 1:].sum(axis=3), d[cols, 1:-1, :, 0]) <= 1e-14


def test_terms_match_numpy_interior(qg, pyorc):
    p, cfg, m = _setup(qg, pyorc, "box_dg")
    sh = (p.nxpo, p.nypo, p.nlo)
    d = m.qocdiag(1).reshape(sh + (5,), order="F")
    po, pom, qo = (m.get_field(n, sh) for n in ("po", "pom", "qo"))
    wek, ent = m.get_field("wekpo", sh[:2]), m.get_field("entoc", sh[:2])
    dxm2 = 1.0 / p.dxo ** 2
    adf = 1.0 / (12.0 * p.dxo ** 2 * p.fnot)
    inner = (slice(3, -3), slice(3, -3))          # three points from the walls: no boundary condition enters
    for k in range(p.nlo):
        d2 = np.zeros(sh[:2]); d4 = np.zeros(sh[:2]); d6 = np.zeros(sh[:2])
        d2[1:-1, 1:-1] = lap5(pom[:, :, k], dxm2)
        d4[1:-1, 1:-1] = lap5(d2, dxm2)
        d6[1:-1, 1:-1] = lap5(d4, dxm2)
        jac = np.zeros(sh[:2])
        jac[1:-1, 1:-1] = adf * jac9(qo[:, :, k], po[:, :, k])
        frc = np.zeros(sh[:2])
        if k == 0:
            frc = (p.fnot / p.hoc[0]) * (wek - ent)
        if k == 1:
            frc = (p.fnot / p.hoc[1]) * ent
        if k == p.nlo - 1:
            frc = frc - 0.5 * np.sign(p.fnot) * p.delek / p.hoc[-1] * d2
        assert rel_l2(d[:, :, k, 1][inner], jac[inner]) <= 1e-12, k
        assert rel_l2(d[:, :, k, 2][inner], (p.ah2oc[k] / p.fnot * d4)[inner]) <= 1e-12 or p.ah2oc[k] == 0.0, k
        assert rel_l2(d[:, :, k, 3][inner], (-p.ah4oc[k] / p.fnot * d6)[inner]) <= 1e-12, k
        assert rel_l2(d[:, :, k, 4][inner], frc[inner]) <= 1e-12, k
    # box walls: only dqdt is filled there, by time differencing (src/qocdiag.F:526-531, :607-612)
    assert not d[0, :, :, 1:].any() and not d[:, 0, :, 1:].any()
    qom = m.get_field("qom", sh)
    assert rel_l2(d[0, :, :, 0], (qo - qom)[0] / p.dto) <= 1e-14


@pytest.mark.parametrize("nsk", [2, 5])
def test_subsampled_terms(qg, pyorc, nsk):
    p, cfg, m = _setup(qg, pyorc, "box_dg")
    full = m.qocdiag(1).reshape((p.nxpo, p.nypo, p.nlo, 5), order="F")
    sub = m.qocdiag(nsk)
    want = full[::nsk, ::nsk]
    assert np.array_equal(sub.reshape(want.shape, order="F"), want)

"""The oracle's restatement of the running sums of src/timavge.F (tavocn :425-617, tavatm
:278-419, avg_ocn_k247 :624-660) against an independent vectorised numpy derivation written
from the Fortran, and the sub-sampling of ocnc_out (src/nc_subs.F:869-890) against slicing."""
import numpy as np
import pytest

from util import small_configs, rel_l2


def _oracle(qg, pyorc, p):
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    return cfg, m


def ocean_terms(m, p, cfg):
    """one contribution of tavocn to every sum, vectorised"""
    nxp, nyp, nxt, nyt = p.nxpo, p.nypo, p.nxto, p.nyto
    po = m.get_field("po", (nxp, nyp, p.nlo))
    qo = m.get_field("qo", (nxp, nyp, p.nlo))
    sst = m.get_field("sst", (nxt, nyt))
    tx, ty = m.get_field("tauxo", (nxp, nyp)), m.get_field("tauyo", (nxp, nyp))
    ug = p.ycexp / (p.dxo * p.fnot)
    rh = 0.5 / (p.fnot * p.hmoc)
    cyc = p.has("cyclic_ocean")
    p1 = po[:, :, 0]
    uu = -ug * (p1[:, 1:] - p1[:, :-1]) + rh * (ty[:, 1:] + ty[:, :-1])          # (nxp, nyt)
    tu = np.empty((nxp, nyt))
    tu[1:-1] = 0.5 * (sst[1:] + sst[:-1])
    if cyc:
        tu[0] = tu[-1] = 0.5 * (sst[0] + sst[-1])
        uu[-1] = uu[0]
    else:
        tu[0], tu[-1] = sst[0], sst[-1]
        uu[0] = uu[-1] = 0.0
    vv = np.zeros((nxt, nyp))
    tv = np.empty((nxt, nyp))
    vv[:, 1:-1] = ug * (p1[1:, 1:-1] - p1[:-1, 1:-1]) - rh * (tx[1:, 1:-1] + tx[:-1, 1:-1])
    tv[:, 1:-1] = 0.5 * (sst[:, 1:] + sst[:, :-1])
    if p.has("sb_hflux"):
        vv[:, 0] = -rh * (tx[1:, 0] + tx[:-1, 0])
        tv[:, 0] = 0.5 * (sst[:, 0] + cfg.tsbdy)
    else:
        tv[:, 0] = sst[:, 0]
    if p.has("nb_hflux"):
        vv[:, -1] = -rh * (tx[1:, -1] + tx[:-1, -1])
        tv[:, -1] = 0.5 * (sst[:, -1] + cfg.tnbdy)
    else:
        tv[:, -1] = sst[:, -1]
    return {"txocav": tx, "tyocav": ty, "wpocav": m.get_field("wekpo", (nxp, nyp)),
            "wtocav": m.get_field("wekto", (nxt, nyt)), "fmocav": m.get_field("fnetoc", (nxt, nyt)), "sstav": sst,
            "uufo": uu, "tufo": tu, "utufo": uu * tu, "vvfo": vv, "tvfo": tv, "vtvfo": vv * tv,
            "pocav": po, "qocav": qo}


def atmos_terms(m, p, cfg):
    nxp, nyp, nxt, nyt = p.nxta + 1, p.nyta + 1, p.nxta, p.nyta
    pa = m.get_field("pa", (nxp, nyp, p.nla))
    qa = m.get_field("qa", (nxp, nyp, p.nla))
    ast = m.get_field("ast", (nxt, nyt))
    tx, ty = m.get_field("tauxa", (nxp, nyp)), m.get_field("tauya", (nxp, nyp))
    ug = 1.0 / (p.ndxr * p.dxo * p.fnot)
    rh = 0.5 / (p.fnot * p.hmat)
    p1 = pa[:, :, 0]
    uu = -ug * (p1[:, 1:] - p1[:, :-1]) - rh * (ty[:, 1:] + ty[:, :-1])
    tu = np.empty((nxp, nyt))
    tu[1:-1] = 0.5 * (ast[1:] + ast[:-1])
    tu[0] = tu[-1] = 0.5 * (ast[0] + ast[-1])
    vv = np.zeros((nxt, nyp))
    tv = np.empty((nxt, nyp))
    vv[:, 1:-1] = ug * (p1[1:, 1:-1] - p1[:-1, 1:-1]) + rh * (tx[1:, 1:-1] + tx[:-1, 1:-1])
    tv[:, 1:-1] = 0.5 * (ast[:, 1:] + ast[:, :-1])
    tv[:, 0], tv[:, -1] = ast[:, 0], ast[:, -1]
    return {"txatav": tx, "tyatav": ty, "wtatav": m.get_field("wekta", (nxt, nyt)),
            "fmatav": m.get_field("fnetat", (nxt, nyt)), "astav": ast,
            "uufa": uu, "tufa": tu, "utufa": uu * tu, "vvfa": vv, "tvfa": tv, "vtvfa": vv * tv,
            "patav": pa, "qatav": qa}


@pytest.mark.parametrize("case", ["box_dg", "chan_so"])
def test_tavocn_matches_numpy(qg, pyorc, case):
    p = small_configs(qg)[case]
    cfg, m = _oracle(qg, pyorc, p)
    m.tavini()
    want = {}
    for rep in range(2):        # two contributions, the state advanced in between
        for k, v in ocean_terms(m, p, cfg).items():
            want[k] = want.get(k, 0.0) + v
        m.tavocn()
        m.avg_ocn_k247()
        if rep == 0:
            m.ocean_step()
    nsumat, nsumoc, nsum_ocavg = m.tav_counts()
    assert (nsumoc, nsum_ocavg) == (2, 2)
    for k, v in want.items():
        got = m.get_field(k, v.shape)
        assert rel_l2(got, v) <= 1e-14, (case, k)
    # po_avg holds po before and after the step, like pocav
    assert rel_l2(m.get_field("po_avg"), m.get_field("pocav")) == 0.0
    m.tavini()
    assert not m.get_field("pocav").any() and m.tav_counts() == (0, 0, 0)


def test_tavatm_matches_numpy(qg, pyorc):
    p = qg.named_config("dg_coupled").scaled(6, 5, ndxr=16, name="cpl_dg")
    cfg, m = _oracle(qg, pyorc, p)
    want = {}
    for rep in range(2):
        for k, v in atmos_terms(m, p, cfg).items():
            want[k] = want.get(k, 0.0) + v
        m.tavatm()              # allocates its sums on first use without touching the ocean's
        if rep == 0:
            m.run(1, 1)
    assert m.tav_counts()[0] == 2
    for k, v in want.items():
        assert rel_l2(m.get_field(k, v.shape), v) <= 1e-14, k


@pytest.mark.parametrize("nsk", [1, 2, 3, 7])
def test_subsample_is_strided_slicing(pyorc, nsk):
    rng = np.random.default_rng(3)
    f = rng.standard_normal((23, 17, 3))
    want = f[::nsk, ::nsk, :]
    got = pyorc.subsample(f, nsk)
    assert got.size == want.size
    assert np.array_equal(got.reshape(want.shape, order="F"), want)

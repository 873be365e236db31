"""-m gpu: device-side ocean section of monnc_comp (src/monitor_diag.F:480-840) against the CPU
oracle through the C ABI.  Area integrals of signed fields are compared against the integral of
the magnitudes they are rounded against (the reference's own OpenMP reduction order is not
reproducible either)."""
import numpy as np
import pytest

from util import small_configs, make_pair
from test_gpu_parity import coupled_configs

pytestmark = pytest.mark.gpu

# quantity -> the always-positive companion that sets its rounding scale (None: itself)
SCALE = {"wetmoc": "watmoc", "wepmoc": "wapmoc", "entmoc": "enamoc", "etamoc": "et2moc_sqrt", "pkenoc": "pke_scale",
         "ddtpeoc": "ddtpe_scale", "ddtkeoc": "ddtke_scale", "utauoc": "utau_scale", "qavgoc": "q_scale", "pavgoc": "p_scale",
         "hfmloc": "hf_scale", "tmlmoc": "t_scale", "occirc": "psi_scale", "occtot": "psi_scale", "osfmin": "psi_scale",
         "osfmax": "psi_scale", "ah2doc": "ah2_scale", "ah4doc": "ah4_scale", "ocjval": "u_scale"}


def scales(m, p, cfg, rep):
    po = m.get_field("po", (p.nxpo, p.nypo, p.nlo))
    pom = m.get_field("pom", (p.nxpo, p.nypo, p.nlo))
    sst, wek = m.get_field("sst"), m.get_field("wekto")
    s = {"et2moc_sqrt": float(np.sqrt(max(rep["et2moc"][:p.nlo - 1]))),
         "p_scale": float(np.abs(po).mean()), "q_scale": float(np.abs(m.get_field("qo")).mean()),
         "t_scale": float(np.abs(sst).mean()), "hf_scale": cfg.rhooc * cfg.cpoc * float(np.abs(sst * wek).mean()),
         "psi_scale": 1e-6 * max(cfg.hoc[:p.nlo]) * float(np.abs(po).max()) / abs(p.fnot),
         "ddtke_scale": max(rep["kealoc"][:p.nlo]) / p.dto, "utau_scale": cfg.rhooc * float(np.abs(m.get_field("tauxo")).max())
         * float(np.abs(po).max()) / (p.dxo * abs(p.fnot))}
    eta1 = np.abs(po[:, :, 1] - po[:, :, 0]).mean() / cfg.gpoc[0]
    s["pke_scale"] = cfg.rhooc * cfg.gpoc[0] * eta1 * float(np.abs(m.get_field("entoc")).mean())
    s["ddtpe_scale"] = cfg.rhooc * max(cfg.gpoc[:p.nlo - 1]) * max(rep["et2moc"][:p.nlo - 1]) * p.nxto * p.nyto / p.dto
    # dissipation integrals: u * del^n(u) summed with alternating signs
    ug = np.abs(np.diff(pom, axis=1)).mean() / (p.dxo * abs(p.fnot))
    s["u_scale"] = float(ug)
    s["ah2_scale"] = cfg.rhooc * max(max(cfg.ah2oc[:p.nlo]), 1e-300) * max(cfg.hoc[:p.nlo]) * ug * ug * 8.0 / p.dxo ** 2
    s["ah4_scale"] = cfg.rhooc * max(cfg.ah4oc[:p.nlo]) * max(cfg.hoc[:p.nlo]) * ug * ug * 64.0 / p.dxo ** 4
    return s


def check(gpu, cpu, p, cfg, label, tol=1e-10):
    a, b = gpu.monnc_ocean().as_dict(), cpu.monnc_ocean().as_dict()
    sc = scales(cpu, p, cfg, b)
    bad = []
    for name, vb in b.items():
        va = a[name]
        if name.startswith("reserved"):
            continue
        if name == "ocjpos":
            # the row of the largest zonal-mean u; a state whose zonal means vanish identically (the
            # synthetic channel modes) leaves rounding noise to pick the row
            for k in range(p.nlo):
                if b["ocjval"][k] > 1e-6 * sc["u_scale"] and va[k] != vb[k]:
                    bad.append((name, k, va[k], vb[k]))
            continue
        xa, xb = np.atleast_1d(va).astype(float), np.atleast_1d(vb).astype(float)
        ref = SCALE.get(name)
        scale = sc[ref] if ref in sc else (max(np.abs(np.atleast_1d(b[ref])).max(), 1e-300) if ref else max(np.abs(xb).max(), 1e-300))
        if not np.abs(xa - xb).max() <= tol * scale:
            bad.append((name, xa.tolist(), xb.tolist(), scale))
    assert not bad, "%s: %s" % (label, bad)


@pytest.mark.parametrize("case", ["box_dg", "box_natl1km", "chan_so", "box_fast"])
def test_monnc_ocean(qg, pyorc, case):
    p = small_configs(qg)[case]
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    check(gpu, cpu, p, cfg, case + " initial state")
    for m in (gpu, cpu):
        m.run(1, 3 * p.nstr)
    check(gpu, cpu, p, cfg, case + " after 3 ocean steps")
    # the diagnostic must not disturb the step that follows
    for m in (gpu, cpu):
        m.run(3 * p.nstr + 1, 4 * p.nstr)
    for name in ("po", "qo", "sst"):
        e = np.linalg.norm(gpu.get_field(name) - cpu.get_field(name)) / np.linalg.norm(cpu.get_field(name))
        assert e <= 1e-11, (case, name, e)


def test_monnc_ocean_coupled(qg, pyorc):
    p = coupled_configs(qg)["cpl_dg"]
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    for m in (gpu, cpu):
        m.run(1, p.nstr)
    check(gpu, cpu, p, cfg, "cpl_dg")


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_monnc_ocean_over_slabs(qg, pyorc, nranks):
    """every rank sums the rows it owns; the shares are added across the ranks"""
    p = small_configs(qg)["box_natl1km"] if nranks == 8 else small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    grp = qg.SlabGroup(cfg, nranks)
    cpu = pyorc.Oracle(cfg)
    for m in (grp, cpu):
        qg.synth.init_model(m, p, cfg, "random")
        m.run(1, 2 * p.nstr)
    check(grp, cpu, p, cfg, "%d slabs" % nranks)
    for m in (grp, cpu):
        m.run(2 * p.nstr + 1, 3 * p.nstr)
    for name in ("po", "qo", "sst"):
        e = np.linalg.norm(grp.get_field(name) - cpu.get_field(name)) / np.linalg.norm(cpu.get_field(name))
        assert e <= 1e-11, (nranks, name, e)


def atmos_scales(m, p, cfg, rep):
    nxp, nyp, nl = p.nxta + 1, p.nyta + 1, p.nla
    pa = m.get_field("pa", (nxp, nyp, nl))
    pam = m.get_field("pam", (nxp, nyp, nl))
    dxa = p.ndxr * p.dxo
    ug = np.abs(np.diff(pam, axis=1)).mean() / (dxa * abs(p.fnot))
    ast, hm = m.get_field("ast"), m.get_field("hmixa")
    et2 = max(rep["et2mat"][:nl - 1])
    s = {"wetmat": rep["watmat"], "wepmat": rep["wapmat"], "entmat": max(rep["enamat"]), "etamat": float(np.sqrt(et2)),
         "pavgat": float(np.abs(pa).mean()), "qavgat": float(np.abs(m.get_field("qa")).mean()),
         "tmlmat": float(np.abs(ast).mean()), "hcmlat": cfg.rhoat * cfg.cpat * float(np.abs(ast * hm).mean()),
         "tmaooc": float(np.abs(ast).mean()),
         "utauat": cfg.rhoat * float(np.abs(m.get_field("tauxa")).max()) * float(np.abs(pa).max()) / (dxa * abs(p.fnot)),
         "ddtkeat": max(rep["kealat"][:nl]) / p.dta + cfg.rhoat * max(cfg.hat[:nl]) * ug * 8.0 / dxa ** 2,
         "ddtpeat": cfg.rhoat * max(cfg.gpat[:nl - 1]) * et2 * p.nxta * p.nyta / p.dta,
         "pkenat": cfg.rhoat * cfg.gpat[0] * float(np.sqrt(et2)) * max(rep["enamat"]),
         "ah4dat": cfg.rhoat * max(cfg.ah4at[:nl]) * max(cfg.hat[:nl]) * ug * ug * 64.0 / dxa ** 4,
         "olrtop": abs(cfg.Dup[nl - 1]) * float(np.abs(ast).mean()) + abs(cfg.Bup[nl - 1]) * p.hmat,
         "davgat": max(float(np.abs(m.get_field("dtopat")).mean()), 1e-300), "u_scale": float(ug), "atstval": float(ug)}
    return s


@pytest.mark.parametrize("case", ["cpl_dg", "cpl_so", "cpl_dg_udiff"])
def test_monnc_atmos(qg, pyorc, case):
    p = coupled_configs(qg)[case]
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    for stage in ("initial state", "after two ocean steps"):
        a, b = gpu.monnc_atmos().as_dict(), cpu.monnc_atmos().as_dict()
        sc = atmos_scales(cpu, p, cfg, b)
        bad = []
        for name, vb in b.items():
            if name.startswith("reserved"):
                continue
            va = a[name]
            if name == "atstpos":
                for k in range(p.nla):
                    if b["atstval"][k] > 1e-6 * sc["u_scale"] and va[k] != vb[k]:
                        bad.append((name, k, va[k], vb[k]))
                continue
            xa, xb = np.atleast_1d(va).astype(float), np.atleast_1d(vb).astype(float)
            scale = max(sc.get(name, 0.0), np.abs(xb).max(), 1e-300)
            if not np.abs(xa - xb).max() <= 1e-10 * scale:
                bad.append((name, xa.tolist(), xb.tolist(), scale))
        assert not bad, "%s %s: %s" % (case, stage, bad)
        for m in (gpu, cpu):
            m.run(1, 2 * p.nstr) if stage == "initial state" else None
    # the diagnostics leave the state alone
    for m in (gpu, cpu):
        m.run(2 * p.nstr + 1, 3 * p.nstr)
    for name in ("pa", "qa", "ast", "po"):
        e = np.linalg.norm(gpu.get_field(name) - cpu.get_field(name)) / np.linalg.norm(cpu.get_field(name))
        assert e <= 1e-10, (case, name, e)


def test_monnc_atmos_absent_in_ocean_only(qg):
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    m = qg.Model(cfg)
    r = m.monnc_atmos().as_dict()
    assert all(not np.any(np.atleast_1d(v)) for v in r.values())

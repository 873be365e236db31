"""The oracle's restatement of the ocean section of monnc_comp (src/monitor_diag.F:480-840, with
del4bx/del4ch/genint) against an independent vectorised numpy evaluation written from the
Fortran."""
import numpy as np
import pytest

from util import small_configs


def genint(v, fw, fs):
    """src/monitor_diag.F:1160: area sum with weights fw on the W/E columns, fs on the S/N rows"""
    w = np.ones(v.shape[0]); w[0] = w[-1] = fw
    s = np.ones(v.shape[1]); s[0] = s[-1] = fs
    return float(w @ v @ s)


class V(float):
    """a float that remembers the sum of magnitudes it was rounded against"""
    def __new__(cls, x, scale):
        o = float.__new__(cls, x)
        o.scale = scale
        return o

    def _s(self, o):
        return getattr(o, "scale", abs(float(o)))

    def __add__(self, o): return V(float(self) + float(o), self.scale + self._s(o))
    __radd__ = __add__
    def __sub__(self, o): return V(float(self) - float(o), self.scale + self._s(o))
    def __mul__(self, o): return V(float(self) * float(o), self.scale * abs(float(o)))
    __rmul__ = __mul__
    def __neg__(self): return V(-float(self), self.scale)


def gint(v, fw, fs):
    return V(genint(v, fw, fs), genint(np.abs(v), fw, fs))


def lap1(a, dxm2, cyclic):
    """one Laplacian of del4bx / del4ch: centred inside, one-sided second differences on walls,
    periodic in x for the channel"""
    if cyclic:
        dxx = np.roll(a, 1, axis=0) - 2 * a + np.roll(a, -1, axis=0)
    else:
        dxx = np.empty_like(a)
        dxx[1:-1] = a[:-2] - 2 * a[1:-1] + a[2:]
        dxx[0] = a[2] - 2 * a[1] + a[0]
        dxx[-1] = a[-1] - 2 * a[-2] + a[-3]
    dyy = np.empty_like(a)
    dyy[:, 1:-1] = a[:, :-2] - 2 * a[:, 1:-1] + a[:, 2:]
    dyy[:, 0] = a[:, 2] - 2 * a[:, 1] + a[:, 0]
    dyy[:, -1] = a[:, -1] - 2 * a[:, -2] + a[:, -3]
    return dxm2 * (dxx + dyy)


@pytest.mark.parametrize("case", ["box_dg", "chan_so"])
def test_monnc_ocean_matches_numpy(qg, pyorc, case):
    p = small_configs(qg)[case]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.run(1, 2 * p.nstr)
    r = m.monnc_ocean().as_dict()
    nxp, nyp, nxt, nyt, nl = p.nxpo, p.nypo, p.nxto, p.nyto, p.nlo
    po, pom, qo = (m.get_field(n, (nxp, nyp, nl)) for n in ("po", "pom", "qo"))
    sst, wekto = m.get_field("sst", (nxt, nyt)), m.get_field("wekto", (nxt, nyt))
    wekpo, entoc = m.get_field("wekpo", (nxp, nyp)), m.get_field("entoc", (nxp, nyp))
    tx, ty = m.get_field("tauxo", (nxp, nyp)), m.get_field("tauyo", (nxp, nyp))
    norm = 1.0 / (nxt * nyt)
    rdxf0 = 1.0 / (p.dxo * p.fnot)
    dxm2 = 1.0 / p.dxo ** 2
    cyc = p.has("cyclic_ocean")
    want = {
        "wetmoc": gint(wekto, 1, 1) * norm, "watmoc": gint(np.abs(wekto), 1, 1) * norm,
        "wepmoc": gint(wekpo, .5, .5) * norm, "wapmoc": gint(np.abs(wekpo), .5, .5) * norm,
        "entmoc": gint(entoc, .5, .5) * norm, "enamoc": gint(np.abs(entoc), .5, .5) * norm,
        "tmlmoc": gint(sst, 1, 1) * norm, "hfmloc": cfg.rhooc * cfg.cpoc * gint(sst * wekto, 1, 1) * norm,
        "sstmin": sst.min(), "sstmax": sst.max(),
    }
    ug1 = -rdxf0 * (po[:, 1:, 0] - po[:, :-1, 0])
    vg1 = rdxf0 * (po[1:, :, 0] - po[:-1, :, 0])
    want["utauoc"] = cfg.rhooc * (gint(vg1 * 0.5 * (ty[1:] + ty[:-1]), 1, .5)
                                  + gint(ug1 * 0.5 * (tx[:, 1:] + tx[:, :-1]), .5, 1)) * norm
    vec = {k: [] for k in ("etamoc", "et2moc", "ddtpeoc", "pavgoc", "qavgoc", "kealoc", "ddtkeoc", "ah2doc", "ah4doc",
                           "osfmin", "osfmax", "occirc", "ocjval", "ocjpos")}
    for k in range(nl - 1):
        eta = (po[:, :, k + 1] - po[:, :, k]) / cfg.gpoc[k]
        etadot = ((po[:, :, k] - po[:, :, k + 1]) - (pom[:, :, k] - pom[:, :, k + 1])) / (cfg.gpoc[k] * p.dto)
        vec["etamoc"].append(gint(eta, .5, .5) * norm)
        vec["et2moc"].append(gint(eta * eta, .5, .5) * norm)
        vec["ddtpeoc"].append(cfg.rhooc * cfg.gpoc[k] * gint(eta * etadot, .5, .5))
        if k == 0:
            want["pkenoc"] = cfg.rhooc * cfg.gpoc[0] * gint(eta * entoc, .5, .5) * norm
    pref = po[0, 0, :] if p.fnot > 0 else po[0, -1, :]
    for k in range(nl):
        ugm = -rdxf0 * (pom[:, 1:, k] - pom[:, :-1, k])
        vgm = rdxf0 * (pom[1:, :, k] - pom[:-1, :, k])
        u2, v2 = lap1(ugm, dxm2, cyc), lap1(vgm, dxm2, cyc)
        u4, v4 = lap1(u2, dxm2, cyc), lap1(v2, dxm2, cyc)
        ug = -rdxf0 * (po[:, 1:, k] - po[:, :-1, k])
        vg = rdxf0 * (po[1:, :, k] - po[:-1, :, k])
        # the reference's ugdot drops po(i,j,k): its two pom(i,j,k) terms cancel (src/monitor_diag.F:676-677)
        ugdot = -(rdxf0 / p.dto) * (po[:, 1:, k] - pom[:, 1:, k])
        vgdot = (rdxf0 / p.dto) * ((po[1:, :, k] - po[:-1, :, k]) - (pom[1:, :, k] - pom[:-1, :, k]))
        h = cfg.hoc[k]
        vec["pavgoc"].append(gint(po[:, :, k], .5, .5) * norm)
        vec["qavgoc"].append(gint(qo[:, :, k], .5, .5) * norm)
        vec["kealoc"].append(0.5 * cfg.rhooc * h * (gint(ug * ug, .5, 1) + gint(vg * vg, 1, .5)) * norm)
        vec["ddtkeoc"].append(cfg.rhooc * h * (gint(ug * ugdot, .5, 1) + gint(vg * vgdot, 1, .5)) * norm)
        vec["ah2doc"].append(-cfg.rhooc * cfg.ah2oc[k] * h * (gint(ug * u2, .5, 1) + gint(vg * v2, 1, .5)) * norm)
        vec["ah4doc"].append(cfg.rhooc * cfg.ah4oc[k] * h * (gint(ug * u4, .5, 1) + gint(vg * v4, 1, .5)) * norm)
        lo, hi = po[:, :, k].min() / p.fnot, po[:, :, k].max() / p.fnot
        vec["osfmin"].append(1e-6 * h * (min(lo, hi) - pref[k] / p.fnot))
        vec["osfmax"].append(1e-6 * h * (max(lo, hi) - pref[k] / p.fnot))
        vec["occirc"].append(1e-6 * h * (po[0, 0, k] - po[0, -1, k]) / p.fnot)
        ujet = np.abs(ug[:-1].sum(axis=0)) / nxt        # the end columns are counted once
        vec["ocjval"].append(ujet.max())
        vec["ocjpos"].append(int(ujet.argmax()) + 1)
    ugb = -rdxf0 * (pom[:, 1:, -1] - pom[:, :-1, -1])
    vgb = rdxf0 * (pom[1:, :, -1] - pom[:-1, :, -1])
    want["btdgoc"] = 0.5 * cfg.rhooc * p.delek * abs(p.fnot) * (gint(ugb * ugb, .5, 1) + gint(vgb * vgb, 1, .5)) * norm
    want["occtot"] = float(sum(vec["occirc"]))

    def close(a, b, name, tol=1e-10):
        assert abs(a - float(b)) <= tol * max(getattr(b, "scale", abs(b)), 1e-300), (case, name, a, float(b))

    for k, v in want.items():
        # sums of signed fields: compare against the sum of magnitudes' rounding
        close(r[k], v, k, tol=1e-9)
    for name, vals in vec.items():
        for k, v in enumerate(vals):
            if name == "ocjpos":
                assert r[name][k] == v, (case, name, k)
            else:
                close(r[name][k], v, "%s[%d]" % (name, k), tol=1e-8 if name in ("ddtkeoc", "ddtpeoc", "ah2doc", "ah4doc") else 1e-9)


@pytest.mark.parametrize("case", ["box_dg", "chan_so"])
def test_couroc_matches_numpy(qg, pyorc, case):
    """face velocities, their extrema and the Courant numbers of couroc (src/monitor_diag.F:1450-1925)"""
    p = small_configs(qg)[case]
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.run(1, 2 * p.nstr)
    r = m.monnc_ocean().as_dict()
    nxp, nyp, nl = p.nxpo, p.nypo, p.nlo
    po = m.get_field("po", (nxp, nyp, nl))
    tx, ty = m.get_field("tauxo", (nxp, nyp)), m.get_field("tauyo", (nxp, nyp))
    rdxf0 = 1.0 / (p.dxo * p.fnot)
    rh = 0.5 / (p.fnot * p.hmoc)
    cyc = p.has("cyclic_ocean")

    def faces(pk, ug, ekman):
        u = -ug * (pk[:, 1:] - pk[:, :-1])                    # (nxp, nyt): every x face of every T row
        v = ug * (pk[1:, :] - pk[:-1, :])                     # (nxt, nyp): every y face of every T column
        if ekman:
            u = u + rh * (ty[:, 1:] + ty[:, :-1])
            v = v - rh * (tx[1:, :] + tx[:-1, :])
        if not cyc:
            u[0] = u[-1] = 0.0
        v[:, 0] = -rh * (tx[1:, 0] + tx[:-1, 0]) if (ekman and p.has("sb_hflux")) else 0.0
        v[:, -1] = -rh * (tx[1:, -1] + tx[:-1, -1]) if (ekman and p.has("nb_hflux")) else 0.0
        vsq = (u[:-1] + u[1:]) ** 2 + (v[:, :-1] + v[:, 1:]) ** 2
        return u.min(), u.max(), v.min(), v.max(), 0.5 / p.dxo * p.dto * np.sqrt(vsq.max())

    want = faces(po[:, :, 0], p.ycexp * rdxf0, True)
    got = (r["umminoc"], r["ummaxoc"], r["vmminoc"], r["vmmaxoc"], r["cnmloc"])
    assert np.allclose(got, want, rtol=1e-12, atol=0.0), (case, got, want)
    for k in range(nl):
        want = faces(po[:, :, k], rdxf0, False)
        got = (r["ugminoc"][k], r["ugmaxoc"][k], r["vgminoc"][k], r["vgmaxoc"][k], r["cnqgoc"][k])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-300), (case, k, got, want)


def test_monnc_atmos_matches_numpy(qg, pyorc):
    """atmosphere section (src/monitor_diag.F:186-478) and courat (:1215-1445), including the
    reference's vkedot slip: it integrates del-sqd of the lagged v (the stale workspace attwk3)"""
    p = qg.named_config("dg_coupled").scaled(6, 5, ndxr=16, name="cpl_dg")
    cfg = qg.build_config(p)
    m = pyorc.Oracle(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.run(1, 2 * p.nstr)
    r = m.monnc_atmos().as_dict()
    nxp, nyp, nxt, nyt, nl = p.nxta + 1, p.nyta + 1, p.nxta, p.nyta, p.nla
    pa, pam, qa = (m.get_field(n, (nxp, nyp, nl)) for n in ("pa", "pam", "qa"))
    ast, hm, wekta = (m.get_field(n, (nxt, nyt)) for n in ("ast", "hmixa", "wekta"))
    wekpa, entat = m.get_field("wekpa", (nxp, nyp)), m.get_field("entat", (nxp, nyp))
    tx, ty = m.get_field("tauxa", (nxp, nyp)), m.get_field("tauya", (nxp, nyp))
    uek, vek = m.get_field("uekat", (nxp, nyt)), m.get_field("vekat", (nxt, nyp))
    dxa = p.ndxr * p.dxo
    norm, rdxf0, dxm2 = 1.0 / (nxt * nyt), 1.0 / (dxa * p.fnot), 1.0 / dxa ** 2
    want = {
        "wetmat": gint(wekta, 1, 1) * norm, "watmat": gint(np.abs(wekta), 1, 1) * norm,
        "wepmat": gint(wekpa, .5, .5) * norm, "wapmat": gint(np.abs(wekpa), .5, .5) * norm,
        "tmlmat": gint(ast, 1, 1) * norm, "hmlmat": gint(hm, 1, 1) * norm,
        "hcmlat": cfg.rhoat * cfg.cpat * gint(ast * hm, 1, 1) * norm,
        "astmin": ast.min(), "astmax": ast.max(),
        "davgat": gint(m.get_field("dtopat", (nxp, nyp)), .5, .5) * norm,
    }
    nxa, nya = p.nxto // p.ndxr, p.nyto // p.ndxr
    want["tmaooc"] = ast[p.nx1 - 1:p.nx1 - 1 + nxa, p.ny1 - 1:p.ny1 - 1 + nya].mean()
    ug1 = -rdxf0 * (pa[:, 1:, 0] - pa[:, :-1, 0])
    vg1 = rdxf0 * (pa[1:, :, 0] - pa[:-1, :, 0])
    want["utauat"] = cfg.rhoat * (gint(vg1 * 0.5 * (ty[1:] + ty[:-1]), 1, .5) + gint(ug1 * 0.5 * (tx[:, 1:] + tx[:, :-1]), .5, 1)) * norm
    vec = {k: [] for k in ("entmat", "enamat", "etamat", "et2mat", "ddtpeat", "pkenat", "pavgat", "qavgat", "kealat", "ddtkeat",
                           "ah4dat", "atstval", "atstpos")}
    vec["entmat"].append(gint(entat, .5, .5) * norm)
    vec["enamat"].append(gint(np.abs(entat), .5, .5) * norm)
    for k in range(nl - 1):
        eta = (pa[:, :, k] - pa[:, :, k + 1]) / cfg.gpat[k]
        etadot = ((pa[:, :, k] - pa[:, :, k + 1]) - (pam[:, :, k] - pam[:, :, k + 1])) / (cfg.gpat[k] * p.dta)
        vec["etamat"].append(gint(eta, .5, .5) * norm)
        vec["et2mat"].append(gint(eta * eta, .5, .5) * norm)
        vec["ddtpeat"].append(cfg.rhoat * cfg.gpat[k] * gint(eta * etadot, .5, .5))
        vec["pkenat"].append(cfg.rhoat * cfg.gpat[0] * gint(eta * entat, .5, .5) * norm if k == 0 else V(0.0, 1.0))
    for k in range(nl):
        ugm = -rdxf0 * (pam[:, 1:, k] - pam[:, :-1, k])
        vgm = rdxf0 * (pam[1:, :, k] - pam[:-1, :, k])
        u4 = lap1(lap1(ugm, dxm2, True), dxm2, True)
        v2 = lap1(vgm, dxm2, True)
        v4 = lap1(v2, dxm2, True)
        ug = -rdxf0 * (pa[:, 1:, k] - pa[:, :-1, k])
        vg = rdxf0 * (pa[1:, :, k] - pa[:-1, :, k])
        ugdot = -(rdxf0 / p.dta) * ((pa[:, 1:, k] - pa[:, :-1, k]) - (pam[:, 1:, k] - pam[:, :-1, k]))
        h = cfg.hat[k]
        vec["pavgat"].append(gint(pa[:, :, k], .5, .5) * norm)
        vec["qavgat"].append(gint(qa[:, :, k], .5, .5) * norm)
        vec["kealat"].append(0.5 * cfg.rhoat * h * (gint(ug * ug, .5, 1) + gint(vg * vg, 1, .5)) * norm)
        vec["ddtkeat"].append(cfg.rhoat * h * (gint(ug * ugdot, .5, 1) + gint(v2, 1, .5)) * norm)      # the vkedot slip
        vec["ah4dat"].append(cfg.rhoat * cfg.ah4at[k] * h * (gint(ug * u4, .5, 1) + gint(vg * v4, 1, .5)) * norm)
        ujet = np.abs(ug[:-1].sum(axis=0)) / nxt
        vec["atstval"].append(ujet.max())
        vec["atstpos"].append(int(ujet.argmax()) + 1)
    olr = cfg.Bup[nl - 1] * (float(want["hmlmat"]) - p.hmat) + cfg.Cup[nl - 1] * float(want["davgat"]) + cfg.Dup[nl - 1] * float(want["tmlmat"])
    for i in range(nl - 1):
        olr += cfg.Aup[(nl - 1) + nl * i] * float(vec["etamat"][i])
    want["olrtop"] = V(olr, abs(cfg.Dup[nl - 1]) * float(np.abs(ast).mean()) + abs(cfg.Bup[nl - 1]) * p.hmat)

    def close(a, b, name, tol):
        assert abs(a - float(b)) <= tol * max(getattr(b, "scale", abs(b)), 1e-300), (name, a, float(b))

    for k, v in want.items():
        close(r[k], v, k, 1e-9)
    for name, vals in vec.items():
        for k, v in enumerate(vals):
            if name == "atstpos":
                assert r[name][k] == v, (name, k)
            else:
                close(r[name][k], v, "%s[%d]" % (name, k), 1e-8)

    def faces(pk, ekman):
        u = -rdxf0 * (pk[:, 1:] - pk[:, :-1])
        v = rdxf0 * (pk[1:, :] - pk[:-1, :])
        if ekman:
            u, v = u + uek, v + vek
            v[:, 0], v[:, -1] = vek[:, 0], vek[:, -1]
        else:
            v[:, 0] = v[:, -1] = 0.0
        vsq = (u[:-1] + u[1:]) ** 2 + (v[:, :-1] + v[:, 1:]) ** 2
        return u.min(), u.max(), v.min(), v.max(), 0.5 / dxa * p.dta * np.sqrt(vsq.max())

    got = (r["umminat"], r["ummaxat"], r["vmminat"], r["vmmaxat"], r["cnmlat"])
    assert np.allclose(got, faces(pa[:, :, 0], True), rtol=1e-12, atol=0.0), got
    for k in range(nl):
        got = (r["ugminat"][k], r["ugmaxat"][k], r["vgminat"][k], r["vgmaxat"][k], r["cnqgat"][k])
        assert np.allclose(got, faces(pa[:, :, k], False), rtol=1e-12, atol=1e-300), (k, got)

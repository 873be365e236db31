"""Pins the oracle's restatement of the coupled forcing xforc (src/xfosubs.F:52-858) with an
independent numpy derivation.  The reference ships no vectors for it, so the check uses
closed forms that do not share code with oracle/orc_xforc.cpp:

* a bicubic Hermite patch whose derivatives are centred differences is the separable
  Catmull-Rom spline, so inside the channel the interpolated wind is a 4x4 convolution;
* the ocean stress is the quadratic drag law of that wind times rhoat/rhooc;
* vekat / wekpa over the ocean are trapezoid sums / box means of quantities that can be
  rebuilt from tauxo, tauyo;
* fnetoc, fnetat use a vectorised bilinear interpolation of astm.
"""
import numpy as np
import pytest


def cr_weights(t):
    """Catmull-Rom weights for the points -1, 0, 1, 2 at fractional position t"""
    return np.array([-0.5 * t + t * t - 0.5 * t ** 3, 1.0 - 2.5 * t * t + 1.5 * t ** 3,
                     0.5 * t + 2.0 * t * t - 1.5 * t ** 3, -0.5 * t * t + 0.5 * t ** 3])


def coupled_case(qg, pyorc, tau_udiff=False, ndxr=16, cyc=False):
    if cyc:
        p = qg.named_config("so_coupled").scaled(12, 3, nxta=12, nyta=9, ndxr=ndxr, name="xf_chan")
    else:
        p = qg.named_config("dg_coupled").scaled(6, 5, ndxr=ndxr, name="xf_box")
    if tau_udiff:
        p.flags = list(p.flags) + ["tau_udiff"]
    cfg = qg.build_config(p)
    o = pyorc.Oracle(cfg)
    qg.synth.init_model(o, p, cfg, "random")    # ends with xforc + homsol
    return p, cfg, o


def wind(p, cfg, pam1):
    """u1at, v1at at atmosphere p points (src/xfosubs.F:186-214), numpy slices"""
    nxp, nyp = p.nxpa, p.nypa
    dxa = p.ndxr * p.dxo
    rdxaf0 = 1.0 / (dxa * p.fnot)
    hx = 0.5 * rdxaf0
    zb = rdxaf0 / (0.5 * p.bccoat + 1.0)
    u = np.zeros((nxp, nyp)); v = np.zeros((nxp, nyp))
    u[:, 0] = -zb * (pam1[:, 1] - pam1[:, 0])
    u[:, -1] = -zb * (pam1[:, -1] - pam1[:, -2])
    u[:, 1:-1] = -hx * (pam1[:, 2:] - pam1[:, :-2])
    pe = np.vstack([pam1[-2:-1, :], pam1])     # periodic: columns -1 .. nxta
    v[:-1, 1:-1] = hx * (pe[2:, 1:-1] - pe[:-2, 1:-1])
    v[-1, :] = v[0, :]
    u[-1, 1:-1] = u[0, 1:-1]
    return u, v


def drag(p, u, v, ab=False):
    raoro = p.rhoat / p.rhooc
    cdh = (p.cdat / p.fnot) * ((1.0 / p.hmat + raoro / p.hmoc) if ab else 1.0 / p.hmat)
    cdr, qu2 = p.cdat / abs(cdh), 4.0 * cdh * cdh
    sq = -0.5 + 0.5 * np.sqrt(1.0 + qu2 * (u * u + v * v))
    sh = np.sqrt(sq)
    cd = cdr * sh / (1.0 + sq)
    return cd * (u - sh * v), cd * (v + sh * u)


@pytest.mark.parametrize("cyc", [False, True])
def test_stress_sampled_on_coarse_grid_is_drag_of_coarse_wind(qg, pyorc, cyc):
    """the interpolant passes through its data, so tauxa/tauya are the drag law of u1at, v1at;
    the northern row is zero because bcuini never fills jj = ndxr"""
    p, cfg, o = coupled_case(qg, pyorc, cyc=cyc)
    pam1 = o.get_field("pam", (p.nxpa, p.nypa, p.nla))[:, :, 0]
    u, v = wind(p, cfg, pam1)
    tx, ty = drag(p, u, v)
    gx, gy = o.get_field("tauxa", (p.nxpa, p.nypa)), o.get_field("tauya", (p.nxpa, p.nypa))
    assert np.abs(gx[:, :-1] - tx[:, :-1]).max() <= 1e-13 * np.abs(tx).max()
    assert np.abs(gy[:, :-1] - ty[:, :-1]).max() <= 1e-13 * np.abs(tx).max()
    assert np.all(gx[:, -1] == 0.0) and np.all(gy[:, -1] == 0.0)


@pytest.mark.parametrize("tau_udiff", [False, True])
def test_ocean_stress_is_catmull_rom_wind_through_drag_law(qg, pyorc, tau_udiff):
    p, cfg, o = coupled_case(qg, pyorc, tau_udiff=tau_udiff)
    n = p.ndxr
    assert p.ny1 >= 3 and p.ny1 - 1 + p.nyto // n <= p.nyta - 2, "ocean window must avoid the wall cells"
    pam1 = o.get_field("pam", (p.nxpa, p.nypa, p.nla))[:, :, 0]
    u, v = wind(p, cfg, pam1)
    nxp, nyp = p.nxpo, p.nypo
    uo = np.zeros((nxp, nyp)); vo = np.zeros((nxp, nyp))
    for io in range(nxp):
        fi = (p.nx1 - 1) * n + io
        ic, ii = divmod(fi, n)
        wx = cr_weights(ii / n)
        cols = [(ic - 1) % p.nxta, ic, ic + 1, (ic + 2) % p.nxta]
        for jo in range(nyp):
            fj = (p.ny1 - 1) * n + jo
            jc, jj = divmod(fj, n)
            wy = cr_weights(jj / n)
            uo[io, jo] = wx @ u[cols][:, jc - 1:jc + 3] @ wy
            vo[io, jo] = wx @ v[cols][:, jc - 1:jc + 3] @ wy
    if tau_udiff:
        po1 = o.get_field("pom", (nxp, nyp, p.nlo))[:, :, 0]
        rdx = 1.0 / (p.dxo * p.fnot)
        zb = rdx / (0.5 * p.bccooc + 1.0)
        uoc = np.zeros((nxp, nyp)); voc = np.zeros((nxp, nyp))
        uoc[:, 0] = -zb * (po1[:, 1] - po1[:, 0]); uoc[:, -1] = -zb * (po1[:, -1] - po1[:, -2])
        uoc[1:-1, 1:-1] = -0.5 * rdx * (po1[1:-1, 2:] - po1[1:-1, :-2])
        voc[1:-1, 1:-1] = 0.5 * rdx * (po1[2:, 1:-1] - po1[:-2, 1:-1])
        voc[0, 1:-1] = zb * (po1[1, 1:-1] - po1[0, 1:-1]); voc[-1, 1:-1] = zb * (po1[-1, 1:-1] - po1[-2, 1:-1])
        uo -= uoc; vo -= voc
    tx, ty = drag(p, uo, vo, ab=tau_udiff)
    raoro = p.rhoat / p.rhooc
    gx, gy = o.get_field("tauxo", (nxp, nyp)), o.get_field("tauyo", (nxp, nyp))
    scale = np.abs(raoro * tx).max()
    assert np.abs(gx - raoro * tx).max() <= 1e-12 * scale
    assert np.abs(gy - raoro * ty).max() <= 1e-12 * scale


def test_ekman_fields_over_the_ocean_follow_from_the_ocean_stress(qg, pyorc):
    p, cfg, o = coupled_case(qg, pyorc)
    n = p.ndxr
    raoro = p.rhoat / p.rhooc
    dxa = n * p.dxo
    tx = o.get_field("tauxo", (p.nxpo, p.nypo)) / raoro
    ty = o.get_field("tauyo", (p.nxpo, p.nypo)) / raoro
    uvekfc = 1.0 / (p.hmat * p.fnot * n)
    vek = o.get_field("vekat", (p.nxta, p.nypa)); uek = o.get_field("uekat", (p.nxpa, p.nyta))
    nxa, nya = p.nxto // n, p.nyto // n
    for ca in range(nxa):
        for cb in range(nya + 1):
            seg = tx[ca * n:ca * n + n + 1, cb * n]
            want = uvekfc * (seg.sum() - 0.5 * (seg[0] + seg[-1]))
            assert abs(vek[p.nx1 - 1 + ca, p.ny1 - 1 + cb] - want) <= 1e-12 * abs(vek).max()
    for ca in range(nxa + 1):
        for cb in range(nya):
            seg = ty[ca * n, cb * n:cb * n + n + 1]
            want = -uvekfc * (seg.sum() - 0.5 * (seg[0] + seg[-1]))
            assert abs(uek[p.nx1 - 1 + ca, p.ny1 - 1 + cb] - want) <= 1e-12 * abs(uek).max()
    # wekta is the divergence of the Ekman transport
    wekta = o.get_field("wekta", (p.nxta, p.nyta))
    want = -(p.hmat / dxa) * (uek[1:, :] - uek[:-1, :] + vek[:, 1:] - vek[:, :-1])
    assert np.abs(wekta - want).max() <= 1e-12 * np.abs(wekta).max()
    # wekpa at p points whose box lies inside the ocean: mean of wekto / raoro (even ndxr)
    wekto = o.get_field("wekto", (p.nxto, p.nyto)) / raoro
    wekpa = o.get_field("wekpa", (p.nxpa, p.nypa))
    h = n // 2
    for ca in range(1, nxa):
        for cb in range(1, nya):
            box = wekto[ca * n - h:ca * n + h, cb * n - h:cb * n + h]
            assert abs(wekpa[p.nx1 - 1 + ca, p.ny1 - 1 + cb] - box.mean()) <= 1e-11 * abs(wekpa).max()


def test_diabatic_forcing_against_vectorised_bilinear(qg, pyorc):
    p, cfg, o = coupled_case(qg, pyorc)
    n = p.ndxr
    dxa = n * p.dxo
    astm = o.get_field("astm", (p.nxta, p.nyta)); sstm = o.get_field("sstm", (p.nxto, p.nyto))
    xo = (p.nx1 - 1) * dxa + (np.arange(p.nxto) + 0.5) * p.dxo
    yo = (p.ny1 - 1) * dxa + (np.arange(p.nyto) + 0.5) * p.dxo
    fx = xo / dxa - 0.5; fy = yo / dxa - 0.5            # position in units of atmosphere T cells
    i0 = np.floor(fx).astype(int); j0 = np.floor(fy).astype(int)
    wx = fx - i0; wy = fy - j0
    ia, ib = i0 % p.nxta, (i0 + 1) % p.nxta
    ja, jb = np.clip(j0, 0, p.nyta - 1), np.clip(j0 + 1, 0, p.nyta - 1)
    asto = ((1 - wx)[:, None] * (1 - wy)[None, :] * astm[ia][:, ja] + wx[:, None] * (1 - wy)[None, :] * astm[ib][:, ja] +
            (1 - wx)[:, None] * wy[None, :] * astm[ia][:, jb] + wx[:, None] * wy[None, :] * astm[ib][:, jb])
    yla = p.nyta * dxa
    fsp_o = cfg.fspco * 0.5 * np.sin(np.pi * (yo - 0.5 * yla) / yla)
    ocnrad = cfg.D0up * sstm; slhf = cfg.xlamda * (sstm - asto)
    want_oc = -fsp_o[None, :] - cfg.Dmdown * asto - ocnrad - slhf
    got_oc = o.get_field("fnetoc", (p.nxto, p.nyto))
    assert np.abs(got_oc - want_oc).max() <= 1e-12 * np.abs(want_oc).max()
    # fnetat: land value, ocean exchange summed per atmosphere cell, then the p-grid terms
    yta = (np.arange(p.nyta) + 0.5) * dxa
    fsp_a = cfg.fspco * 0.5 * np.sin(np.pi * (yta - 0.5 * yla) / yla)
    fa = -fsp_a[None, :] - cfg.Dmup * astm
    ex = (p.dxo / dxa) ** 2 * (ocnrad + (cfg.Dmdown - cfg.Dmup) * asto + slhf)
    nxa, nya = p.nxto // n, p.nyto // n
    fa[p.nx1 - 1:p.nx1 - 1 + nxa, p.ny1 - 1:p.ny1 - 1 + nya] = ex.reshape(nxa, n, nya, n).sum(axis=(1, 3))
    pam = o.get_field("pam", (p.nxpa, p.nypa, p.nla)); dp = pam[:, :, 0] - pam[:, :, 1]
    dtop = o.get_field("dtopat", (p.nxpa, p.nypa)); hmm = o.get_field("hmixam", (p.nxta, p.nyta))
    corner = lambda f: f[:-1, :-1] + f[1:, :-1] + f[:-1, 1:] + f[1:, 1:]
    fa = fa - cfg.Adown[0] * 0.25 / cfg.gpat[0] * corner(dp) - 0.25 * (cfg.Cmup + cfg.C1down) * corner(dtop) \
        + (-cfg.hmadmp - cfg.Bmup - cfg.B1down) * (hmm - cfg.hmat)
    got_at = o.get_field("fnetat", (p.nxta, p.nyta))
    assert np.abs(got_at - fa).max() <= 1e-11 * np.abs(fa).max()
    s = o.get_scalars().as_dict()
    assert abs(s["slhfav"] - slhf.mean()) <= 1e-11 * abs(slhf).max()
    assert abs(s["oradav"] - ocnrad.mean()) <= 1e-11 * abs(ocnrad).max()
    land = np.ones((p.nxta, p.nyta), bool); land[p.nx1 - 1:p.nx1 - 1 + nxa, p.ny1 - 1:p.ny1 - 1 + nya] = False
    assert abs(s["arlaav"] - cfg.Dmup * astm[land].mean()) <= 1e-11 * abs(cfg.Dmup * astm).max()

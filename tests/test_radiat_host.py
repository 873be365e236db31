"""The host-side restatement of radiat (src/radsubs.f:44-592, q-gcm_b200/radiat.py), which only
builds test and bench configurations (it stays Fortran in deployment), against the balances
the reference itself prints and against independent quadrature / finite differences."""
import math

import numpy as np
import pytest
from scipy.integrate import quad

STEFAN = 5.67040e-8


def _radiat(qg, p):
    import importlib
    return importlib.import_module(qg.__name__ + ".radiat").radiat(p)


@pytest.mark.parametrize("deck", ["dg_coupled", "so_coupled"])
def test_radiative_equilibrium_balances(qg, deck):
    p = qg.named_config(deck)
    r = _radiat(qg, p)
    # "Fractional error in OLR" (src/radsubs.f:282-283): outgoing long wave balances the mean forcing
    assert abs(r["Fupbar"][p.nla - 1] + r["fsbar"]) / abs(r["fsbar"]) <= 1e-10
    # ocean mixed layer (:188-204): lambda To + sigma To^4 = lambda Ta + sigma/2 Ta^4 - fsbar
    ta, to = r["tmbara"], r["tmbaro"]
    lhs = p.xlamda * to + STEFAN * to ** 4
    rhs = p.xlamda * ta + 0.5 * STEFAN * ta ** 4 - p.fsbar
    assert abs(lhs - rhs) <= 1e-9 * abs(rhs)
    assert np.allclose(r["toc"], np.array(p.tabsoc[: p.nlo]) - to) and np.allclose(r["tat"], np.array(p.tabsat[: p.nla]) - ta)
    # boundary temperatures are the ends of the equilibrium SST profile (:544-548)
    assert r["tsbdy"] == r["sstbar"][0] and r["tnbdy"] == r["sstbar"][-1]
    assert math.copysign(1.0, r["fspco"]) == math.copysign(1.0, p.fnot)


def test_mixed_layer_flux_and_its_temperature_derivative(qg):
    """Fm-up = (sigma/2 zm) int_0^hm (Tm - gamma z)^4 exp(-(hm - z)/zm) dz by adaptive quadrature, and
    Dmup = dFm-up/dTm by central differences; D0up is the black-body derivative 4 sigma To^3"""
    p = qg.named_config("dg_coupled")
    r = _radiat(qg, p)

    def fm_up(tm):
        f = lambda z: (tm - p.gamma * z) ** 4 * math.exp(-(p.hmat - z) / p.zm)
        return 0.5 * STEFAN * quad(f, 0.0, p.hmat, epsabs=0.0, epsrel=1e-13)[0] / p.zm

    ta = r["tmbara"]
    h = 1e-3
    dmup = (fm_up(ta + h) - fm_up(ta - h)) / (2 * h)
    assert abs(r["Dmup"] - dmup) <= 1e-7 * abs(dmup)
    assert abs(r["D0up"] - 4.0 * STEFAN * r["tmbaro"] ** 3) <= 1e-14 * r["D0up"]
    assert abs(r["Dmdown"] + 2.0 * STEFAN * ta ** 3) <= 1e-14 * abs(r["Dmdown"])
    # the temperature perturbation is attenuated by every layer above: Dup(k) = Dmup * prod tau
    tau = [math.exp(-(p.hat[0] - p.hmat) / p.zopt[0])] + [math.exp(-p.hat[k] / p.zopt[k]) for k in range(1, p.nla)]
    assert np.allclose(r["Dup"], r["Dmup"] * np.cumprod(tau), rtol=1e-13)
    # and the upward flux at the top of the atmosphere, rebuilt layer by layer with quadrature
    f_up = fm_up(ta)
    hbot = p.hmat
    for k in range(p.nla):
        htop = p.hat[0] if k == 0 else hbot + p.hat[k]
        g = lambda z, k=k, htop=htop: (p.tabsat[k] - p.gamma * z) ** 4 * math.exp(-(htop - z) / p.zopt[k])
        up = 0.5 * STEFAN * quad(g, hbot, htop, epsabs=0.0, epsrel=1e-13)[0] / p.zopt[k]
        f_up = f_up * tau[k] + up
        assert abs(r["Fupbar"][k] - f_up) <= 2e-7 * abs(f_up), k      # 10001-point trapezoid rule vs adaptive quadrature
        hbot = htop

"""Host-side preparation of a qgcm_config: what the Fortran main program has computed
before the time loop starts.

* ``Params`` holds the positional values of ``input.params`` (src/in_param.f:31-142) and
  the compile-time PARAMETERs of ``parameters_data.F`` (src/parameters_data.F:41-119);
  ``named_config`` returns the five benchmark decks of BASELINE.json.
* ``eigmod`` restates src/eigmode.f:130-144 (A matrix), :310-327 (Flierl normalisation)
  and :386-428 (mode/layer conversion matrices) with numpy/scipy in place of the six
  LAPACK routines the reference calls (LAPACK is not vendored in the reference).
* ``build_config`` fills the C struct.

In the drop-in deployment all of this stays Fortran (BASELINE.json north_star); this
module exists so tests and bench.py can drive the library without a Fortran compiler.
"""
from dataclasses import dataclass, field, replace
from typing import List

import numpy as np

from .abi import QgcmConfig, FLAGS, NLMAX, ABI_VERSION
import ctypes as C


@dataclass
class Params:
    name: str = "custom"
    # parameters_data.F
    nxta: int = 384
    nyta: int = 96
    nla: int = 3
    nxaooc: int = 60
    nyaooc: int = 60
    ndxr: int = 16
    nlo: int = 3
    fnot: float = 9.37456e-05
    beta: float = 1.75360e-11
    # make.config flags
    flags: List[str] = field(default_factory=lambda: ["ocean_only", "sb_hflux"])
    # input.params
    dta: float = 180.0
    nstr: int = 3
    dxo: float = 5.0e3
    delek: float = 2.0
    cdat: float = 1.3e-3
    rhoat: float = 1.0
    rhooc: float = 1.0e3
    cpat: float = 1.0e3
    cpoc: float = 4.0e3
    bccoat: float = 1.0
    bccooc: float = 0.2
    xcexp: float = 1.0
    ycexp: float = 1.0
    xlamda: float = 35.0
    hmoc: float = 100.0
    st2d: float = 100.0
    st4d: float = 2.0e9
    hmat: float = 1000.0
    hmamin: float = 100.0
    ahmd: float = 2.0e5
    at2d: float = 2.5e4
    at4d: float = 2.0e14
    hmadmp: float = 0.15
    fsbar: float = -210.0
    fspamp: float = 80.0
    zm: float = 200.0
    zopt: List[float] = field(default_factory=lambda: [2.0e4, 2.0e4, 3.0e4])
    gamma: float = 1.0e-2
    ah2oc: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    ah4oc: List[float] = field(default_factory=lambda: [2.0e9, 2.0e9, 2.0e9])
    tabsoc: List[float] = field(default_factory=lambda: [287.0, 282.0, 276.0])
    hoc: List[float] = field(default_factory=lambda: [350.0, 750.0, 2900.0])
    gpoc: List[float] = field(default_factory=lambda: [0.0150, 0.0075])
    ah4at: List[float] = field(default_factory=lambda: [1.5e14, 1.5e14, 1.5e14])
    tabsat: List[float] = field(default_factory=lambda: [330.0, 340.0, 350.0])
    hat: List[float] = field(default_factory=lambda: [2000.0, 3000.0, 4000.0])
    gpat: List[float] = field(default_factory=lambda: [1.2, 0.4])

    # ---- derived grid parameters, src/parameters_data.F:78-88
    @property
    def nxto(self):
        return self.ndxr * self.nxaooc

    @property
    def nyto(self):
        return self.ndxr * self.nyaooc

    @property
    def nxpo(self):
        return self.nxto + 1

    @property
    def nypo(self):
        return self.nyto + 1

    @property
    def nxpa(self):
        return self.nxta + 1

    @property
    def nypa(self):
        return self.nyta + 1

    @property
    def nx1(self):
        return 1 + (self.nxta - self.nxaooc) // 2

    @property
    def ny1(self):
        return 1 + (self.nyta - self.nyaooc) // 2

    def has(self, flag):
        return flag in self.flags

    @property
    def dto(self):
        return self.nstr * self.dta

    @property
    def dxa(self):
        return self.ndxr * self.dxo

    def scaled(self, nxaooc, nyaooc, nxta=None, nyta=None, ndxr=None, name=None):
        """same physics on a smaller grid (used by parity tests)"""
        cyc = self.has("cyclic_ocean")
        nxta = nxta if nxta is not None else (nxaooc if cyc else max(nxaooc, 2 * nxaooc))
        nyta = nyta if nyta is not None else max(nyaooc, 2 * nyaooc)
        return replace(self, nxaooc=nxaooc, nyaooc=nyaooc, nxta=nxta, nyta=nyta,
                       ndxr=ndxr if ndxr is not None else self.ndxr,
                       name=name or (self.name + "_small"))


def named_config(name: str) -> Params:
    """the benchmark decks of BASELINE.json `configs`"""
    n = name.lower()
    if n in ("dg_oo", "double_gyre_ocean_only", "c1"):
        # examples/double_gyre_ocean_only: parameters_data.F.dg_oo:43,48, input.params.dg_oo,
        # make.config.dg_oo (ocean_only sb_hflux qoc_diag)
        return Params(name="dg_oo")
    if n in ("dg_coupled", "double_gyre_coupled", "c2"):
        # examples/double_gyre_coupled: make.config.coupled (sb_hflux only)
        return Params(name="dg_coupled", flags=["sb_hflux"])
    if n in ("so_coupled", "southern_ocean_coupled", "so5", "c3"):
        # src/parameters_data.F.SOcn.5km.wideatm:44,49,98; make.config.so_coupled
        return Params(name="so_coupled", nxta=288, nyta=108, nxaooc=288, nyaooc=36, ndxr=16,
                      fnot=-1.19467e-04, beta=1.31301e-11, flags=["cyclic_ocean", "nb_hflux"])
    if n in ("natl2km", "n2", "c4"):
        # src/parameters_data.F.NAtl.2km:45,50 + src/input.params.NAtl.2km + config.NAtl.oconly
        return Params(name="natl2km", nxta=768, nyta=192, nxaooc=120, nyaooc=120, ndxr=20,
                      flags=["ocean_only", "sb_hflux", "tau_udiff"], nstr=2, dxo=2.0e3,
                      bccooc=0.1, st2d=200.0, st4d=5.0e8, ah4oc=[5.0e8] * 3, ah4at=[1.0e14] * 3)
    if n in ("natl1km", "n1", "c5"):
        # src/parameters_data.F.NAtl.1km:45,50 + src/input.params.NAtl.1km (nstr=1: quirk 3)
        return Params(name="natl1km", nxta=768, nyta=192, nxaooc=120, nyaooc=120, ndxr=40,
                      flags=["ocean_only", "sb_hflux", "tau_udiff"], nstr=1, dxo=1.0e3,
                      bccooc=0.1, st2d=200.0, st4d=5.0e7, ah4oc=[5.0e7] * 3, ah4at=[1.0e14] * 3)
    raise KeyError("unknown config %r" % name)


def eigmod(nl, gpr, h, fnot, ocean):
    """src/eigmode.f:41-440.  Returns (amat, rdm2, ctl2m, ctm2l) as Fortran-ordered arrays
    with the reference's index meaning: ctl2m[k,m], ctm2l[m,k] (0-based here)."""
    import scipy.linalg as sla

    gpr = np.asarray(gpr, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    a = np.zeros((nl, nl))
    # eigmode.f:130-144
    a[0, 1] = -1.0 / (gpr[0] * h[0])
    a[0, 0] = -a[0, 1]
    for k in range(1, nl - 1):
        a[k, k - 1] = -1.0 / (gpr[k - 1] * h[k])
        a[k, k + 1] = -1.0 / (gpr[k] * h[k])
        a[k, k] = -a[k, k - 1] - a[k, k + 1]
    k = nl - 1
    a[k, k - 1] = -1.0 / (gpr[k - 1] * h[k])
    a[k, k] = -a[k, k - 1]
    w, vl, vr = sla.eig(a, left=True, right=True)
    wre = w.real
    evecl = vl.real.copy()
    evecr = vr.real.copy()
    if ocean:
        # eigmode.f:310-327 Flierl normalisation, +ve at the surface
        htotal = h.sum()
        for m in range(nl):
            dotp = np.sum(h * evecr[:, m] * evecr[:, m])
            flfac = np.copysign(np.sqrt(htotal / dotp), evecr[0, m])
            evecr[:, m] *= flfac
    elder = evecl.T @ evecr
    c2rabs = np.abs(wre)
    index = np.argsort(c2rabs, kind="stable")
    rdm2 = np.zeros(nl)
    for m in range(1, nl):
        rdm2[m] = fnot * fnot * c2rabs[index[m]]
    ctl2m = np.zeros((nl, nl))
    ctm2l = np.zeros((nl, nl))
    for m in range(nl):
        inm = index[m]
        for k in range(nl):
            ctl2m[k, m] = evecl[k, inm] / elder[inm, inm]
            ctm2l[m, k] = evecr[k, inm]
    return a, rdm2, ctl2m, ctm2l


def _put(dst, values):
    v = np.asarray(values, dtype=np.float64).ravel(order="F")
    for i, x in enumerate(v):
        dst[i] = float(x)


def build_config(p: Params, device: int = 0, rank: int = 0, nranks: int = 1, radiation=None) -> QgcmConfig:
    """fill the C struct.  ``radiation`` (optional dict) carries the outputs of radiat
    (src/radsubs.f:44-592); without it, layer temperatures are taken relative to a fixed
    mean (ocean-only synthetic runs never read the radiation coefficients)."""
    c = QgcmConfig()
    c.abi_version = ABI_VERSION
    c.struct_bytes = C.sizeof(QgcmConfig)
    c.flags = sum(FLAGS[f] for f in p.flags)
    c.device = device
    c.nxto, c.nyto, c.nlo = p.nxto, p.nyto, p.nlo
    c.nxta, c.nyta, c.nla = p.nxta, p.nyta, p.nla
    c.ndxr, c.nx1, c.ny1, c.nstr = p.ndxr, p.nx1, p.ny1, p.nstr
    c.nranks, c.rank = nranks, rank
    for k in ("fnot", "beta", "dxo", "dta", "delek", "cdat", "rhoat", "rhooc", "cpat", "cpoc", "bccoat",
              "bccooc", "xcexp", "ycexp", "xlamda", "hmoc", "st2d", "st4d", "hmat", "hmamin", "ahmd",
              "at2d", "at4d", "hmadmp"):
        setattr(c, k, float(getattr(p, k)))
    _put(c.hoc, p.hoc)
    _put(c.gpoc, p.gpoc)
    _put(c.ah2oc, p.ah2oc)
    _put(c.ah4oc, p.ah4oc)
    _put(c.hat, p.hat)
    _put(c.gpat, p.gpat)
    _put(c.ah4at, p.ah4at)
    a, rdm2, l2m, m2l = eigmod(p.nlo, p.gpoc, p.hoc, p.fnot, ocean=True)
    _put(c.amatoc, a)
    _put(c.rdm2oc, rdm2)
    _put(c.ctl2moc, l2m)
    _put(c.ctm2loc, m2l)
    a, rdm2, l2m, m2l = eigmod(p.nla, p.gpat, p.hat, p.fnot, ocean=False)
    _put(c.amatat, a)
    _put(c.rdm2at, rdm2)
    _put(c.ctl2mat, l2m)
    _put(c.ctm2lat, m2l)
    if radiation is None:
        from .radiat import radiat
        radiation = radiat(p)
    _put(c.toc, radiation["toc"])
    _put(c.tat, radiation["tat"])
    for k in ("tsbdy", "tnbdy", "fspco", "Bmup", "B1down", "Cmup", "C1down", "D0up", "Dmup", "Dmdown",
              "bface", "cface", "dface"):
        setattr(c, k, float(radiation[k]))
    for k in ("Aup", "Adown", "Bup", "Cup", "Dup", "rbetat", "aface"):
        _put(getattr(c, k), radiation[k])
    c._radiation = radiation  # keep sstbar/astbar handy for synth
    return c

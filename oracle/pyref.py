"""TEST INFRASTRUCTURE ONLY -- Python handle on the translated reference
(oracle/_ref/libqgcmref.so: the reference's own Fortran sources, translated statement by
statement by oracle/f2cpp.py and compiled by g++; see oracle/Makefile, target `ref`).

The library keeps the reference's program structure: module variables and SAVEd locals are
global to a loaded library, so every `Reference` loads its own private copy of the .so (a
temporary file, unlinked once mapped): several configurations can coexist in one process.
`setup` fills the module variables the Fortran main program
would fill from input.params (src/q-gcm.F:377-452, :929-973) and then calls the reference's
subroutines by name.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir("/root/reference/src") or os.path.exists(os.path.join(REFDIR, "libqgcmref_box.so"))


def provenance():
    """how the library under oracle/_ref was made (written by `make ref` next to it)"""
    try:
        with open(os.path.join(REFDIR, "PROVENANCE.txt")) as f:
            return f.read()
    except OSError:
        return "unknown"


def build():
    """translate + compile (needs /root/reference; on a box without it the prebuilt .so is used)"""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Reference:
    """one configuration of the translated reference; variant = 'box', 'chan' or 'coupled' (the
    cpp configuration the library was translated with, oracle/Makefile)"""

    def __init__(self, variant, params):
        path = os.path.join(REFDIR, "libqgcmref_%s.so" % variant)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise RuntimeError("%s is missing and /root/reference is not here to build it from" % path)
        import shutil
        import tempfile
        fd, tmp = tempfile.mkstemp(suffix=".so", prefix="qgcmref_")
        os.close(fd)
        shutil.copyfile(path, tmp)
        try:
            self.lib = C.CDLL(tmp)
        finally:
            os.unlink(tmp)
        self.lib.ref_last_error.restype = C.c_char_p
        for k, v in params.items():
            if self.lib.ref_set_param(k.encode(), C.c_double(float(v))) != 0:
                raise RuntimeError("ref_set_param(%s): the module storage is already initialised" % k)
        if self.lib.ref_init() != 0:
            raise RuntimeError("ref_init failed")
        self._views = {}

    def var(self, name):
        """numpy view (flat, Fortran element order) of a module variable"""
        if name not in self._views:
            addr, cnt, ty = C.c_void_p(), C.c_long(), C.c_int()
            if self.lib.ref_var(name.encode(), C.byref(addr), C.byref(cnt), C.byref(ty)) != 0:
                raise KeyError("the translated reference has no module variable %r" % name)
            ct = {0: C.c_int, 1: C.c_double, 2: C.c_bool}[ty.value]
            buf = (ct * cnt.value).from_address(addr.value)
            self._views[name] = np.frombuffer(buf, dtype={0: np.int32, 1: np.float64, 2: np.bool_}[ty.value])
        return self._views[name]

    def has(self, name):
        try:
            self.var(name)
            return True
        except KeyError:
            return False

    def set(self, name, value):
        v = self.var(name)
        a = np.asarray(value, dtype=v.dtype).ravel(order="F")
        if a.size != v.size:
            raise ValueError("%s: %d elements given, the reference declares %d" % (name, a.size, v.size))
        v[:] = a

    def get(self, name, shape=None):
        a = self.var(name).copy()
        return a.reshape(shape, order="F") if shape is not None else a

    def call(self, name, *args):
        """args: numpy arrays (passed by address, as Fortran does)"""
        arr = (C.c_void_p * max(1, len(args)))(*[a.ctypes.data for a in args])
        rc = self.lib.ref_call(name.encode(), arr, C.c_int(len(args)))
        if rc != 0:
            raise RuntimeError("reference procedure %s: rc=%d %s" % (name, rc, (self.lib.ref_last_error() or b"").decode()))


def base_parameters(p):
    """the PARAMETERs of src/parameters_data.F a deck edits (grid sizes, f0, beta); the derived ones
    (nxpo, nx1, atnorm ...) are evaluated by the translated module from these, in source order"""
    return {"nxta": p.nxta, "nyta": p.nyta, "nla": p.nla, "nxaooc": p.nxaooc, "nyaooc": p.nyaooc, "ndxr": p.ndxr,
            "nlo": p.nlo, "fnot": p.fnot, "beta": p.beta}


def variant_of(p):
    cyc = p.has("cyclic_ocean")
    if p.has("ocean_only"):
        return "chan" if cyc else "box"
    return "chancpl" if cyc else "boxcpl"


def setup(ref, p, cfg):
    """what the Fortran main program computes from input.params before the time loop
    (src/q-gcm.F:377-452 grids and derived constants, :929-973 Helmholtz coefficients and FFTPACK
    tables -- the latter through the reference's own dsinti / drffti)"""
    rad = cfg._radiation
    nlo, nla = p.nlo, p.nla
    dxo, dxa = p.dxo, p.ndxr * p.dxo
    dyo, dya = dxo, dxa
    xla, yla = p.nxta * dxa, p.nyta * dya
    v = {}
    # src/q-gcm.F:377-441
    xpa = np.arange(p.nxpa) * dxa
    ypa = np.arange(p.nypa) * dya
    v.update(dxa=dxa, dya=dya, hdxam1=0.5 / dxa, dxam2=1.0 / (dxa * dxa), xla=xla, yla=yla, xpa=xpa, xta=xpa[:-1] + 0.5 * dxa,
             ypa=ypa, yparel=ypa - 0.5 * yla, yta=ypa[:-1] + 0.5 * dya, ytarel=(ypa[:-1] + 0.5 * dya) - 0.5 * yla)
    xpo = np.arange(p.nxpo) * dxo + (p.nx1 - 1) * dxa
    ypo = (p.ny1 - 1) * dya + np.arange(p.nypo) * dyo
    v.update(dxo=dxo, dyo=dyo, hdxom1=0.5 / dxo, dxom2=1.0 / (dxo * dxo), xlo=p.nxto * dxo, ylo=p.nyto * dyo, xpo=xpo,
             xto=xpo[:-1] + 0.5 * dxo, ypo=ypo, yporel=ypo - 0.5 * yla, yto=ypo[:-1] + 0.5 * dyo,
             ytorel=(ypo[:-1] + 0.5 * dyo) - 0.5 * yla)
    v.update(rdxaf0=1.0 / (dxa * p.fnot), rdxof0=1.0 / (dxo * p.fnot), rrcpat=1.0 / (p.rhoat * p.cpat),
             rrcpoc=1.0 / (p.rhooc * p.cpoc), raoro=p.rhoat / p.rhooc, dta=p.dta, dto=p.nstr * p.dta,
             tdto=2.0 * p.nstr * p.dta, tdta=2.0 * p.dta, nstr=p.nstr, hto=sum(p.hoc[:nlo]), hta=sum(p.hat[:nla]))
    # input.params (src/in_param.f)
    for k in ("delek", "cdat", "rhoat", "rhooc", "cpat", "cpoc", "bccoat", "bccooc", "xcexp", "ycexp", "xlamda", "hmoc", "st2d",
              "st4d", "hmat", "hmamin", "ahmd", "at2d", "at4d", "hmadmp", "fsbar", "fspamp", "zm", "gamma"):
        v[k] = getattr(p, k)
    v.update(gpoc=p.gpoc[:nlo - 1], hoc=p.hoc[:nlo], ah2oc=p.ah2oc[:nlo], ah4oc=p.ah4oc[:nlo], tabsoc=p.tabsoc[:nlo],
             gpat=p.gpat[:nla - 1], hat=p.hat[:nla], ah4at=p.ah4at[:nla], tabsat=p.tabsat[:nla], zopt=p.zopt[:nla])
    # eigmod outputs (src/eigmode.f; LAPACK-based, supplied by the harness to every implementation alike)
    def mat(a, n):
        return np.asarray(list(a)[:n * n], dtype=np.float64)
    v.update(amatoc=mat(cfg.amatoc, nlo), ctl2moc=mat(cfg.ctl2moc, nlo), ctm2loc=mat(cfg.ctm2loc, nlo), rdm2oc=list(cfg.rdm2oc)[:nlo],
             amatat=mat(cfg.amatat, nla), ctl2mat=mat(cfg.ctl2mat, nla), ctm2lat=mat(cfg.ctm2lat, nla), rdm2at=list(cfg.rdm2at)[:nla])
    # radiat outputs (src/radsubs.f)
    v.update(toc=list(cfg.toc)[:nlo], tat=list(cfg.tat)[:nla], tsbdy=cfg.tsbdy, tnbdy=cfg.tnbdy, fspco=cfg.fspco,
             sstbar=rad["sstbar"], astbar=rad["astbar"])
    for k in ("Bmup", "B1down", "Cmup", "C1down", "D0up", "Dmup", "Dmdown", "bface", "cface", "dface"):
        v[k.lower()] = float(rad[k])
    for k in ("Aup", "Adown", "Bup", "Cup", "Dup", "rbetat", "aface"):
        v[k.lower()] = np.asarray(rad[k], dtype=np.float64)
    for k, val in v.items():
        if ref.has(k):
            ref.set(k, val)
    # src/q-gcm.F:929-973
    PI, TWOPI = 3.14159265358979324, 6.28318530717958648
    if ref.has("bd2oc"):
        aoc = 1.0 / (dyo * dyo)
        dxom2 = 1.0 / (dxo * dxo)
        nxto = p.nxto
        bd2 = np.zeros(nxto)
        if p.has("cyclic_ocean"):
            for i in range(2, nxto // 2 + 1):
                i1 = 2 * i - 1
                bd2[i1 - 2] = -2.0 * aoc + 2.0 * dxom2 * (np.cos((i - 1) * TWOPI / nxto) - 1.0)
                bd2[i1 - 1] = bd2[i1 - 2]
            bd2[0] = -2.0 * aoc
            bd2[nxto - 1] = -2.0 * aoc - 4.0 * dxom2
            ref.call("drffti", np.array([nxto], dtype=np.int32), ref.var("oftwrk"))
        else:
            for i in range(2, nxto + 1):
                bd2[i - 2] = -2.0 * aoc + 2.0 * dxom2 * (np.cos((i - 1) * PI / nxto) - 1.0)
            bd2[nxto - 1] = 0.0
            ref.call("dsinti", np.array([nxto - 1], dtype=np.int32), ref.var("oftwrk"))
        ref.set("aoc", aoc)
        ref.set("bd2oc", bd2)
    if ref.has("bd2at"):
        aat = 1.0 / (dya * dya)
        dxam2 = 1.0 / (dxa * dxa)
        nxta = p.nxta
        bd2 = np.zeros(nxta)
        for i in range(2, nxta // 2 + 1):
            i1 = 2 * i - 1
            bd2[i1 - 2] = -2.0 * aat + 2.0 * dxam2 * (np.cos((i - 1) * TWOPI / nxta) - 1.0)
            bd2[i1 - 1] = bd2[i1 - 2]
        bd2[0] = -2.0 * aat
        bd2[nxta - 1] = -2.0 * aat - 4.0 * dxam2
        ref.call("drffti", np.array([nxta], dtype=np.int32), ref.var("aftwrk"))
        ref.set("aat", aat)
        ref.set("bd2at", bd2)


class RefModel:
    """the translated reference behind the same method names as the oracle / CUDA bindings, so that
    tests drive all three alike (set_field / get_field use the reference's Fortran names)"""

    def __init__(self, p, cfg):
        self.p, self.cfg = p, cfg
        self.ref = Reference(variant_of(p), base_parameters(p))
        setup(self.ref, p, cfg)
        self._i = lambda x: np.array([x], dtype=np.int32)

    def set_field(self, name, arr):
        self.ref.set(name, arr)

    def get_field(self, name, shape=None):
        return self.ref.get(name, shape)

    def scalars(self, names):
        return {n: self.ref.get(n).tolist() for n in names if self.ref.has(n)}

    def _v(self, n):
        return self.ref.var(n)

    def constr(self): self.ref.call("constr")
    def homsol(self): self.ref.call("homsol")
    def xforc(self): self.ref.call("xforc")
    def oml(self): self.ref.call("oml")
    def qgostep(self): self.ref.call("qgostep")
    def ocinvq(self): self.ref.call("ocinvq")
    def ocqbdy(self): self.ref.call("ocqbdy", self._v("qo"), self._v("po"))
    def aml(self): self.ref.call("aml")
    def qgastep(self): self.ref.call("qgastep")
    def atinvq(self): self.ref.call("atinvq")
    def atqzbd(self): self.ref.call("atqzbd", self._v("qa"), self._v("pa"))

    def qcomp_ocean(self):
        """src/q-gcm.F:719-732"""
        p, v, i = self.p, self._v, self._i
        for q, pp in (("qo", "po"), ("qom", "pom")):
            self.ref.call("qcomp", v(q), v(pp), v("amatoc"), v("yporel"), v("dxom2"), i(p.nxpo), i(p.nypo), i(p.nlo), v("ddynoc"), i(p.nlo))
        for q, pp in (("qo", "po"), ("qom", "pom")):
            self.ref.call("ocqbdy", v(q), v(pp))
        if p.has("cyclic_ocean"):
            for q, pp in (("qo", "po"), ("qom", "pom")):
                self.ref.call("merqcy", v(q), v(pp), v("amatoc"), v("yporel"), v("dxom2"), i(p.nxpo), i(p.nypo), i(p.nlo), v("ddynoc"), i(p.nlo))

    def qcomp_atmos(self):
        """src/q-gcm.F:734-746"""
        p, v, i = self.p, self._v, self._i
        for q, pp in (("qa", "pa"), ("qam", "pam")):
            self.ref.call("qcomp", v(q), v(pp), v("amatat"), v("yparel"), v("dxam2"), i(p.nxpa), i(p.nypa), i(p.nla), v("ddynat"), i(1))
        for q, pp in (("qa", "pa"), ("qam", "pam")):
            self.ref.call("atqzbd", v(q), v(pp))
        for q, pp in (("qa", "pa"), ("qam", "pam")):
            self.ref.call("merqcy", v(q), v(pp), v("amatat"), v("yparel"), v("dxam2"), i(p.nxpa), i(p.nypa), i(p.nla), v("ddynat"), i(1))

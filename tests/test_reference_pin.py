"""The CPU oracle against the REFERENCE'S OWN CODE: oracle/_ref holds the reference's Fortran
sources translated statement by statement to C++ by oracle/f2cpp.py (the tool knows nothing of
what the code computes) and compiled by g++ -- the reference itself run here, since neither this
image nor the GPU box has a Fortran compiler (profiles/r02_fortran_probe.txt).  Every procedure
of the hot path is called in the reference's main-loop order (src/q-gcm.F:711-976 start-up,
:1222-1269 time loop) on both implementations from identical inputs and every field and scalar
is compared after every call.

What the translated reference does NOT cover (said here so nobody reads more into "pinned"):
eigmod (six LAPACK routines the reference does not vendor; the mode matrices are supplied by the
harness to all implementations alike, pinned by identities in tests/test_oracle_identities.py),
the main program's inline code (grid set-up, time-level average: restated in oracle/pyref.py
and the oracle from src/q-gcm.F:377-452, :929-973, :1328-1407), and LAPACK's DGETRF/DGETRS/DGERFS
(restated from the published algorithms in oracle/f2c_lapack.cpp).  Arithmetic differs from a
gfortran build only in what g++ and gfortran may legally do differently with the same
expression trees (both without fast-math; FMA contraction off here).

Tolerances: 1e-12 relative L2 (measured 1e-16 .. 1e-15: the oracle's FFT is not FFTPACK, so the
inversion differs in the last bits, everything else is bit-identical or at 1e-17)."""
import os

import numpy as np
import pytest

import pyref

pytestmark = pytest.mark.skipif(not pyref.available(), reason="neither /root/reference nor a prebuilt oracle/_ref")

TOL = 1e-12

OCEAN_FIELDS = ("po", "pom", "qo", "qom", "sst", "sstm", "entoc", "wekto", "wekpo", "tauxo", "tauyo", "fnetoc")
ATMOS_FIELDS = ("pa", "pam", "qa", "qam", "ast", "astm", "hmixa", "hmixam", "entat", "wekta", "wekpa", "tauxa", "tauya",
                "fnetat", "uekat", "vekat")
OCEAN_SCAL_BOX = ("dpioc", "dpiocp", "xon", "aipohs", "cdiffo", "cdhoc")
OCEAN_SCAL_CHAN = ("dpioc", "dpiocp", "xon", "ocncs", "ocncn", "ocncsp", "ocncnp", "enisoc", "eninoc", "ajisoc", "ajinoc", "ap3soc",
                   "ap3noc", "ap5soc", "ap5noc", "txisoc", "txinoc", "bdrins", "bdrinn", "hc1soc", "hc2soc", "hc1noc", "hc2noc",
                   "aipcho", "hbsioc", "aipbho")
ATMOS_SCAL = ("dpiat", "dpiatp", "xan", "atmcs", "atmcn", "atmcsp", "atmcnp", "enisat", "eninat", "ajisat", "ajinat", "ap5sat",
              "ap5nat", "txisat", "txinat", "hc1sat", "hc2sat", "hc1nat", "hc2nat", "aipcha", "hbsiat", "aipbha")


def decks(qg):
    box = qg.named_config("dg_oo").scaled(3, 2, ndxr=16, name="pin_box")                      # 48 x 32 T cells
    box1 = qg.named_config("natl1km").scaled(2, 3, ndxr=20, name="pin_box_natl")               # 40 x 60, nstr = 1 physics
    box1.flags = ["ocean_only", "sb_hflux"]          # tau_udiff only matters inside the coupled xforc
    chan = qg.named_config("so_coupled").scaled(4, 2, nxta=4, nyta=6, ndxr=16, name="pin_chan")
    chan.flags = ["ocean_only", "cyclic_ocean", "nb_hflux"]
    cpl = qg.named_config("dg_coupled").scaled(3, 2, ndxr=8, name="pin_boxcpl")
    ccpl = qg.named_config("so_coupled").scaled(6, 2, nxta=6, nyta=6, ndxr=8, name="pin_chancpl")
    return {"box": box, "box_natl": box1, "chan": chan, "boxcpl": cpl, "chancpl": ccpl}


def rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a))


def compare(cpu, ref, p, label, worst):
    names = ()
    scal = ()
    if not p.has("atmos_only"):
        names += OCEAN_FIELDS + (("pch1oc", "pch2oc", "pbhoc") if p.has("cyclic_ocean") else ("ochom",))
        scal += OCEAN_SCAL_CHAN if p.has("cyclic_ocean") else OCEAN_SCAL_BOX
    if not p.has("ocean_only"):
        names += ATMOS_FIELDS + ("pch1at", "pch2at", "pbhat")
        scal += ATMOS_SCAL
    bad = []
    for n in names:
        e = rel(cpu.get_field(n), ref.get_field(n))
        worst[n] = max(worst.get(n, 0.0), e)
        if not e <= TOL:
            bad.append((n, e))
    sc = cpu.get_scalars().as_dict()
    rs = ref.scalars(scal)
    for n, rv in rs.items():
        a = np.atleast_1d(np.asarray(sc[n], dtype=float))[: len(rv)]
        b = np.asarray(rv, dtype=float)
        if n in ("cdiffo", "cdhoc"):        # (nlo, nlo-1) / (nlo-1, nlo-1) matrices: same element order, packed
            a = np.atleast_1d(np.asarray(sc[n], dtype=float))[: b.size]
        # sums of signed terms: the scale is the largest magnitude among the entries (a zero entry that is
        # a cancelled sum is compared against its siblings)
        scale = max(np.abs(b).max(), 1e-300)
        if n == "xon":          # area integral of entoc after its mean was removed: pure rounding, scale = integral of |entoc|
            scale = max(scale, float(np.abs(cpu.get_field("entoc")).sum() * p.dxo ** 2))
        if n in ("dpioc", "dpiocp"):      # area integrals of layer differences of a field without a mean
            scale = max(scale, float(np.abs(cpu.get_field("po")).sum() * p.dxo ** 2))
        if n in ("dpiat", "dpiatp"):
            scale = max(scale, float(np.abs(cpu.get_field("pa")).sum() * (p.ndxr * p.dxo) ** 2))
        if n == "xan":
            scale = max(scale, float(np.abs(cpu.get_field("entat")).sum() * (p.ndxr * p.dxo) ** 2))
        e = float(np.abs(a - b).max() / scale)
        worst[n] = max(worst.get(n, 0.0), e)
        if not e <= 1e-9:
            bad.append((n, e, a.tolist(), b.tolist()))
    assert not bad, "%s: oracle and translated reference differ: %s" % (label, bad)


@pytest.mark.parametrize("deck", ["box", "box_natl", "chan", "boxcpl", "chancpl"])
def test_oracle_matches_the_translated_reference(qg, pyorc, deck):
    p = decks(qg)[deck]
    cfg = qg.build_config(p)
    cpu = pyorc.Oracle(cfg)
    ref = pyref.RefModel(p, cfg)
    amp = min(1.0, (p.nxto * p.dxo) / 4.8e6 * 4.0)
    st = qg.synth.ocean_state(p, cfg, "random", qg.synth.SEED, amp)
    if not p.has("ocean_only"):
        st.update(qg.synth.atmos_state(p, cfg, "random", qg.synth.SEED + 1))
    for k, v in st.items():
        cpu.set_field(k, v)
        ref.set_field(k, v)
    worst = {}
    # start-up, src/q-gcm.F:711-976
    seq = ["constr", "qcomp_ocean"] + ([] if p.has("ocean_only") else ["qcomp_atmos"]) + ["xforc", "homsol"]
    for name in seq:
        getattr(cpu, name)()
        getattr(ref, name)()
        compare(cpu, ref, p, "%s after %s" % (deck, name), worst)
    # the time loop, src/q-gcm.F:1222-1269, for nt = 1 .. 2*nstr+1 (three ocean steps)
    nstr = p.nstr
    for nt in range(1, 2 * nstr + 2):
        ocean = (nstr == 1) or (nt % nstr == 1)
        steps = []
        if ocean:
            steps += ([] if p.has("ocean_only") else ["xforc"]) + ["oml", "qgostep", "ocinvq", "ocqbdy"]
        if not p.has("ocean_only"):
            steps += ["aml", "qgastep", "atinvq", "atqzbd"]
        for name in steps:
            getattr(cpu, name)()
            getattr(ref, name)()
            compare(cpu, ref, p, "%s nt=%d after %s" % (deck, nt, name), worst)
    print("worst relative differences, %s:" % deck, {k: "%.1e" % v for k, v in sorted(worst.items()) if v > 0})


LAYERS = {2: ([350.0, 3650.0], [0.02]),
          4: ([300.0, 700.0, 1000.0, 2000.0], [0.02, 0.01, 0.005]),
          5: ([250.0, 450.0, 800.0, 1000.0, 1500.0], [0.02, 0.012, 0.008, 0.004])}


@pytest.mark.parametrize("nlo", [2, 4, 5])
@pytest.mark.parametrize("cyc", [0, 1])
def test_oracle_matches_the_translated_reference_for_other_layer_counts(qg, pyorc, nlo, cyc):
    """every deck ships three layers; the GPU parity tests also run two, four and five
    (tests/test_gpu_parity.py::test_other_layer_counts), so the oracle is pinned there too: start-up and
    three ocean steps of a box and a channel deck, every field and scalar after every call"""
    from dataclasses import replace
    base = decks(qg)["chan" if cyc else "box"]
    hoc, gp = LAYERS[nlo]
    tabs = (list(base.tabsoc) + [base.tabsoc[-1]] * nlo)[:nlo]
    p = replace(base, nlo=nlo, hoc=hoc, gpoc=gp, ah2oc=[0.0] * nlo, ah4oc=[2.0e9] * nlo, tabsoc=tabs, name="pin_nl%d_%d" % (nlo, cyc))
    p.flags = list(base.flags)
    cfg = qg.build_config(p)
    cpu = pyorc.Oracle(cfg)
    ref = pyref.RefModel(p, cfg)
    st = qg.synth.ocean_state(p, cfg, "random", qg.synth.SEED, min(1.0, (p.nxto * p.dxo) / 4.8e6 * 4.0))
    for k, v in st.items():
        cpu.set_field(k, v)
        ref.set_field(k, v)
    worst = {}
    for name in ["constr", "qcomp_ocean", "xforc", "homsol"] + 3 * ["oml", "qgostep", "ocinvq", "ocqbdy"]:
        getattr(cpu, name)()
        getattr(ref, name)()
        compare(cpu, ref, p, "%d layers, %s, after %s" % (nlo, "channel" if cyc else "box", name), worst)


def test_translator_handles_the_fortran_it_claims(tmp_path):
    """f2cpp on a hand-made unit: declared lower bounds, sequence association, DATA, labelled DO, GOTO,
    integer division, real->integer truncation, sign(), mod(), x**n, DO trip count fixed at entry"""
    import ctypes as C
    import subprocess
    import sys
    src = tmp_path / "t.f"
    src.write_text("""      module tmod
      implicit none
      integer n
      parameter ( n = 4 )
      double precision a(0:n,2), s
      end module tmod
      subroutine fill (v, m)
      integer m
      double precision v(m)
      integer i
      do 10 i=1,m
         v(i) = dble(i)**2 - 7/2
   10 continue
      end
      subroutine drive
      use tmod
      implicit none
      integer i, k, lim, w(3)
      double precision t
      data w /1, 2*5/
      call fill (a(1,2), 3)
      a(0,1) = sign(2.5d0, -1.0d0) + mod(7,4) + w(2)
      lim = 3
      k = 0
      do i=1,lim
         lim = 10
         k = k + 1
      enddo
      t = 7.9d0
      i = t
      if (k .eq. 3) goto 20
      k = -99
   20 s = a(1,2) + a(2,2) + a(3,2) + a(0,1) + k + i
      end
""")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "t.cpp"
    subprocess.check_call([sys.executable, os.path.join(root, "oracle", "f2cpp.py"), "--src", str(tmp_path), "--files", "t.f",
                           "--want", "drive", "-o", str(out)])
    so = tmp_path / "t.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-I", os.path.join(root, "oracle"), "-o", str(so), str(out)])
    lib = C.CDLL(str(so))
    assert lib.ref_init() == 0
    arr = (C.c_void_p * 1)()
    assert lib.ref_call(b"drive", arr, 0) == 0
    addr, cnt, ty = C.c_void_p(), C.c_long(), C.c_int()
    assert lib.ref_var(b"s", C.byref(addr), C.byref(cnt), C.byref(ty)) == 0
    s = C.c_double.from_address(addr.value).value
    # fill: v(i) = i**2 - 3 -> -2, 1, 6; a(0,1) = -2.5 + 3 + 5 = 5.5; k = 3 (trip count fixed at 3); i = 7
    assert s == (-2.0 + 1.0 + 6.0) + 5.5 + 3 + 7
    assert lib.ref_var(b"a", C.byref(addr), C.byref(cnt), C.byref(ty)) == 0 and cnt.value == 10


@pytest.mark.parametrize("deck", ["boxcpl", "chancpl"])
def test_radiat_matches_the_translated_reference(qg, deck):
    """row a20: the harness's Python restatement of radiat (q-gcm_b200/radiat.py) against the reference's own
    radiat (src/radsubs.f:44-592, translated; its LAPACK calls go to oracle/f2c_lapack.cpp)"""
    p = decks(qg)[deck]
    cfg = qg.build_config(p)
    rad = cfg._radiation
    ref = pyref.RefModel(p, cfg)
    # wipe what setup() copied from the Python restatement, then let the reference compute it
    for n in ("toc", "tat", "sstbar", "astbar", "rbetat", "aface", "aup", "adown", "bup", "cup", "dup"):
        ref.ref.var(n)[:] = np.nan
    for n in ("tsbdy", "tnbdy", "fspco", "tmbaro", "tmbara", "bmup", "b1down", "cmup", "c1down", "d0up", "dmup", "dmdown", "bface", "cface", "dface"):
        ref.ref.var(n)[:] = np.nan
    ref.ref.call("radiat")
    worst = 0.0
    for n in ("toc", "tat", "sstbar", "astbar", "rbetat", "aface", "Aup", "Adown", "Bup", "Cup", "Dup", "tsbdy", "tnbdy", "fspco", "Bmup",
              "B1down", "Cmup", "C1down", "D0up", "Dmup", "Dmdown", "bface", "cface", "dface"):
        want = ref.ref.get(n.lower())
        got = np.atleast_1d(np.asarray(rad[n], dtype=float)).ravel(order="F")
        assert np.isfinite(want).all(), n
        e = float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-300))
        worst = max(worst, e)
        assert e <= 1e-10, (n, e, got.tolist()[:4], want.tolist()[:4])
    print("radiat worst relative difference:", worst)


OC_SUMS = ("txocav", "tyocav", "wpocav", "wtocav", "fmocav", "sstav", "uufo", "tufo", "utufo", "vvfo", "tvfo", "vtvfo", "pocav", "qocav", "po_avg")
AT_SUMS = ("txatav", "tyatav", "wtatav", "fmatav", "astav", "uufa", "tufa", "utufa", "vvfa", "tvfa", "vtvfa", "patav", "qatav")


@pytest.mark.parametrize("deck", ["box", "chan", "boxcpl", "chancpl"])
def test_diagnostics_match_the_translated_reference(qg, pyorc, deck):
    """SURVEY 8f rows that moved to the device: valids (src/valsubs.F:43), the running sums of
    src/timavge.F:109-662 and monnc_comp with couroc / courat (src/monitor_diag.F:89-893, :1215-1928),
    oracle restatement against the reference's own translated code"""
    p = decks(qg)[deck]
    cfg = qg.build_config(p)
    cpu = pyorc.Oracle(cfg)
    ref = pyref.RefModel(p, cfg)
    amp = min(1.0, (p.nxto * p.dxo) / 4.8e6 * 4.0)
    st = qg.synth.ocean_state(p, cfg, "random", qg.synth.SEED, amp)
    if not p.has("ocean_only"):
        st.update(qg.synth.atmos_state(p, cfg, "random", qg.synth.SEED + 1))
    for k, v in st.items():
        cpu.set_field(k, v)
        ref.set_field(k, v)
    coupled = not p.has("ocean_only")
    for name in ["constr", "qcomp_ocean"] + (["qcomp_atmos"] if coupled else []) + ["xforc", "homsol"]:
        getattr(cpu, name)()
        getattr(ref, name)()

    def ocean_step():
        for name in (["xforc"] if coupled else []) + ["oml", "qgostep", "ocinvq", "ocqbdy"]:
            getattr(cpu, name)()
            getattr(ref, name)()

    def atmos_step():
        for name in ["aml", "qgastep", "atinvq", "atqzbd"]:
            getattr(cpu, name)()
            getattr(ref, name)()

    ocean_step()
    if coupled:
        atmos_step()
    # ---- running sums
    cpu.tavini()
    ref.ref.call("tavini")
    ref.ref.set("nsum_ocavg", 0)
    for _ in range(2):
        ocean_step()
        cpu.tavocn(); ref.ref.call("tavocn")
        cpu.avg_ocn_k247(); ref.ref.call("avg_ocn_k247")
        if coupled:
            atmos_step()
            cpu.tavatm(); ref.ref.call("tavatm")
    for n in OC_SUMS + (AT_SUMS if coupled else ()):
        e = rel(cpu.get_field(n), ref.get_field(n))
        assert e <= TOL, (deck, n, e)
    assert cpu.tav_counts()[1] == int(ref.ref.get("nsumoc")[0]) == 2
    # ---- monnc_comp: every member of the C structs that the reference's module `monitor` also holds
    ref.ref.call("monnc_comp")
    from test_gpu_monitor import SCALE, scales, atmos_scales      # the rounding scales of the signed integrals
    checked = 0
    rep_o = cpu.monnc_ocean().as_dict()
    sc_o = scales(cpu, p, cfg, rep_o)
    todo = [(rep_o, lambda n, xb: sc_o[SCALE[n]] if SCALE.get(n) in sc_o else
             (max(np.abs(np.atleast_1d(rep_o[SCALE[n]])).max(), 1e-300) if SCALE.get(n) else max(np.abs(xb).max(), 1e-300)))]
    if coupled:
        rep_a = cpu.monnc_atmos().as_dict()
        sc_a = atmos_scales(cpu, p, cfg, rep_a)
        todo.append((rep_a, lambda n, xb: max(sc_a.get(n, 0.0), np.abs(xb).max(), 1e-300)))
    for rep, scale_of in todo:
        for name, val in rep.items():
            if name.startswith("reserved") or not ref.ref.has(name):
                continue
            want = ref.ref.get(name).astype(float)
            got = np.atleast_1d(np.asarray(val, dtype=float))[: want.size]
            if name in ("ocjpos", "atstpos"):
                vname, uscale = ("ocjval", sc_o["u_scale"]) if name == "ocjpos" else ("atstval", sc_a["u_scale"])
                for k in range(want.size):       # a zonal mean that vanishes identically leaves rounding to pick the row
                    if np.atleast_1d(rep[vname])[k] > 1e-6 * uscale:
                        assert got[k] == want[k], (deck, name, k)
                checked += 1
                continue
            e = float(np.abs(got - want).max() / scale_of(name, want))
            assert e <= 1e-10, (deck, name, e, got.tolist(), want.tolist())
            checked += 1
    assert checked >= (30 if not coupled else 55), checked
    # ---- valids: same verdict on the healthy state and on a state whose top layer is too thin
    ok_ref = np.array([True], dtype=np.bool_)
    ref.ref.call("valids", ok_ref)
    assert bool(ok_ref[0]) == bool(cpu.valids().solnok) == True      # noqa: E712
    po = cpu.get_field("po", (p.nxpo, p.nypo, p.nlo)).copy()
    po[: p.nxpo // 3, :, 1] += 400.0 * cfg.gpoc[0]
    cpu.set_field("po", po)
    ref.set_field("po", po)
    ok_ref[0] = True
    ref.ref.call("valids", ok_ref)
    assert bool(ok_ref[0]) == bool(cpu.valids().solnok) == False      # noqa: E712

"""-m gpu: the CUDA path against the REFERENCE'S OWN CODE, procedure by procedure.  oracle/_ref holds the
reference's Fortran sources translated to C++ by oracle/f2cpp.py and compiled by g++ (no Fortran compiler
exists here or on the GPU box); the built libraries travel to the GPU box with the repository snapshot.
Both sides start from the same seeded state and are called in the reference's main-loop order
(src/q-gcm.F:711-976, :1222-1269); every field is compared after every call.
Tolerance: FP64 relative L2 <= 1e-11 per field (BASELINE.json north_star)."""
import os
import sys

import numpy as np
import pytest

import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_vectors as mrv  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pyref.available(), reason="no oracle/_ref on this box")]

TOL = 1e-11
OCEAN = ("po", "pom", "qo", "qom", "sst", "sstm", "entoc", "wekto", "wekpo")
ATMOS = ("pa", "pam", "qa", "qam", "ast", "astm", "hmixa", "hmixam", "entat", "wekta", "wekpa", "tauxa", "tauya", "tauxo", "tauyo",
         "fnetoc", "fnetat", "uekat", "vekat")


def rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a))


@pytest.mark.parametrize("deck", ["box", "box_natl", "box_fast", "chan", "boxcpl", "chancpl"])
def test_cuda_matches_the_translated_reference_call_by_call(qg, deck):
    p = mrv.decks(qg)[deck]
    cfg = qg.build_config(p)
    gpu = qg.Model(cfg)
    ref = pyref.RefModel(p, cfg)
    amp = min(1.0, (p.nxto * p.dxo) / 4.8e6 * 4.0)
    st = qg.synth.ocean_state(p, cfg, "random", qg.synth.SEED, amp)
    if not p.has("ocean_only"):
        st.update(qg.synth.atmos_state(p, cfg, "random", qg.synth.SEED + 1))
    for k, v in st.items():
        gpu.set_field(k, v)
        ref.set_field(k, v)
    names = OCEAN + (("pch1oc", "pch2oc", "pbhoc") if p.has("cyclic_ocean") else ("ochom",))
    if not p.has("ocean_only"):
        names += ATMOS + ("pch1at", "pch2at", "pbhat")
    worst = {}

    def check(label):
        bad = []
        for n in names:
            e = rel(gpu.get_field(n), ref.get_field(n))
            worst[n] = max(worst.get(n, 0.0), e)
            if not e <= TOL:
                bad.append((n, e))
        assert not bad, "%s %s: CUDA and the translated reference differ: %s" % (deck, label, bad)

    for name in ["constr", "qcomp_ocean"] + ([] if p.has("ocean_only") else ["qcomp_atmos"]) + ["xforc", "homsol"]:
        getattr(gpu, name)()
        getattr(ref, name)()
        check("after " + name)
    nstr = p.nstr
    for nt in range(1, 2 * nstr + 2):
        steps = []
        if nstr == 1 or nt % nstr == 1:
            steps += ([] if p.has("ocean_only") else ["xforc"]) + ["oml", "qgostep", "ocinvq", "ocqbdy"]
        if not p.has("ocean_only"):
            steps += ["aml", "qgastep", "atinvq", "atqzbd"]
        for name in steps:
            getattr(gpu, name)()
            getattr(ref, name)()
            check("nt=%d after %s" % (nt, name))
    print("worst CUDA-vs-reference relative L2, %s:" % deck, {k: "%.1e" % v for k, v in sorted(worst.items()) if v > 0})

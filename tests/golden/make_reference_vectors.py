"""Golden vectors made by the REFERENCE ITSELF (python tests/golden/make_reference_vectors.py; needs
/root/reference): the reference's Fortran sources, translated by oracle/f2cpp.py and compiled by
g++ (oracle/_ref, see oracle/README.md), run from seeded synthetic states through the start-up
sequence of src/q-gcm.F:711-976 and three ocean steps of the time loop (:1222-1269; the inline
time-level average of the main program is not part of the translated procedures and is left out
on every side).  Per field: element count, sum, sum of squares and 64 evenly spaced values.
tests/test_reference_vectors.py checks the CPU oracle (here) and the CUDA path (on the GPU box,
where /root/reference does not exist) against the committed file."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

OCEAN = ("po", "pom", "qo", "qom", "sst", "sstm", "entoc", "wekto", "wekpo")
ATMOS = ("pa", "pam", "qa", "qam", "ast", "astm", "hmixa", "hmixam", "entat", "wekta", "wekpa", "tauxa", "tauya", "tauxo",
         "tauyo", "fnetoc", "fnetat", "uekat", "vekat")
NSAMP = 64


def decks(qg):
    box = qg.named_config("dg_oo").scaled(3, 2, ndxr=16, name="pin_box")
    box1 = qg.named_config("natl1km").scaled(2, 3, ndxr=20, name="pin_box_natl")
    box1.flags = ["ocean_only", "sb_hflux"]
    # long enough for the CUDA library's three-pass DST plan and the fused inversion (half length 480)
    fast = qg.named_config("natl1km").scaled(24, 12, ndxr=40, name="pin_box_fast")
    fast.flags = ["ocean_only", "sb_hflux"]
    chan = qg.named_config("so_coupled").scaled(4, 2, nxta=4, nyta=6, ndxr=16, name="pin_chan")
    chan.flags = ["ocean_only", "cyclic_ocean", "nb_hflux"]
    cpl = qg.named_config("dg_coupled").scaled(3, 2, ndxr=8, name="pin_boxcpl")
    ccpl = qg.named_config("so_coupled").scaled(6, 2, nxta=6, nyta=6, ndxr=8, name="pin_chancpl")
    return {"box": box, "box_natl": box1, "box_fast": fast, "chan": chan, "boxcpl": cpl, "chancpl": ccpl}


def fingerprint(a):
    a = np.asarray(a, dtype=np.float64).ravel()
    idx = np.linspace(0, a.size - 1, NSAMP).astype(int)
    return {"n": int(a.size), "sum": float(a.sum()), "sumsq": float((a * a).sum()), "samples": [float(a[i]) for i in idx]}


def drive(qg, m, p, cfg):
    """start-up + three ocean steps in main-loop order through the procedure calls every binding
    has (oracle, CUDA model, translated reference)"""
    amp = min(1.0, (p.nxto * p.dxo) / 4.8e6 * 4.0)
    st = qg.synth.ocean_state(p, cfg, "random", qg.synth.SEED, amp)
    if not p.has("ocean_only"):
        st.update(qg.synth.atmos_state(p, cfg, "random", qg.synth.SEED + 1))
    for k, v in st.items():
        m.set_field(k, v)
    for name in ["constr", "qcomp_ocean"] + ([] if p.has("ocean_only") else ["qcomp_atmos"]) + ["xforc", "homsol"]:
        getattr(m, name)()
    nstr = p.nstr
    for nt in range(1, 2 * nstr + 2):
        if nstr == 1 or nt % nstr == 1:
            for name in ([] if p.has("ocean_only") else ["xforc"]) + ["oml", "qgostep", "ocinvq", "ocqbdy"]:
                getattr(m, name)()
        if not p.has("ocean_only"):
            for name in ["aml", "qgastep", "atinvq", "atqzbd"]:
                getattr(m, name)()
    names = OCEAN + (ATMOS if not p.has("ocean_only") else ())
    return {n: fingerprint(m.get_field(n)) for n in names}


def main():
    import _pkg
    import pyref
    qg = _pkg.load()
    pyref.build()
    out = {"_made_by": "tests/golden/make_reference_vectors.py", "_provenance": pyref.provenance().splitlines()[0]}
    for name, p in decks(qg).items():
        cfg = qg.build_config(p)
        out[name] = drive(qg, pyref.RefModel(p, cfg), p, cfg)
        print(name, "done")
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()

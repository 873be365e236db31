"""Regression fingerprints of the CPU oracle (python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot be compiled in the build image, so these are
NOT reference outputs: they freeze what the oracle -- after it was pinned by the identities and
the independent numpy derivations under tests/ -- produces on seeded reduced decks, so that a
later edit of oracle/ or of the synthetic states cannot silently move every parity target.
Per field: sum, sum of squares and five sampled values after the listed calls."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

OCEAN = ("po", "qo", "sst", "entoc", "wekpo")
ATMOS = ("pa", "qa", "ast", "hmixa", "entat", "tauxo", "fnetoc")


def cases(qg):
    from util import small_configs
    from test_gpu_parity import coupled_configs
    c = dict(small_configs(qg))
    c.pop("box_fast")
    c.update({k: v for k, v in coupled_configs(qg).items() if k != "cpl_dg_udiff"})
    return c


def fingerprint(a):
    a = np.asarray(a, dtype=np.float64).ravel()
    idx = np.linspace(0, a.size - 1, 5).astype(int)
    return {"n": int(a.size), "sum": float(a.sum()), "sumsq": float((a * a).sum()), "samples": [float(a[i]) for i in idx]}


def run_case(qg, pyorc, p, make=None):
    cfg = qg.build_config(p)
    m = (make or pyorc.Oracle)(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.run(1, 2 * p.nstr + 1)
    names = OCEAN + (ATMOS if not p.has("ocean_only") else ())
    out = {n: fingerprint(m.get_field(n)) for n in names}
    mon = m.monnc_ocean().as_dict()
    out["monitor"] = {k: mon[k] for k in ("utauoc", "btdgoc", "tmlmoc", "cnmloc")}
    out["monitor"]["kealoc"] = list(mon["kealoc"][: p.nlo])
    return out


def main():
    import _pkg
    import pyorc
    qg = _pkg.load()
    pyorc.build()
    out = {name: run_case(qg, pyorc, p) for name, p in cases(qg).items()}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_fingerprints.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()

"""Measured CUDA-vs-oracle differences behind the tolerances of tests/test_gpu_parity.py and
tests/test_gpu_slabs.py (python scripts/tolerance_survey.py > gpurun_out/tolerances.json on the GPU box).
The tests assert 10 x what this prints, rounded up to a power of ten."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402
import pyorc  # noqa: E402
from util import small_configs, make_pair, rel_l2, integral_scale, OCEAN_CHECK  # noqa: E402
from test_gpu_parity import coupled_configs, ATMOS_CHECK  # noqa: E402

qg = _pkg.load()
out = {}


def scal(gpu, cpu, names, floor=None):
    sg, sc = gpu.get_scalars().as_dict(), cpu.get_scalars().as_dict()
    worst = 0.0
    for n in names:
        a, b = np.atleast_1d(sg[n]).astype(float), np.atleast_1d(sc[n]).astype(float)
        scale = max(np.abs(b).max(), 1e-300)
        if floor is not None:
            scale = max(scale, floor)
        worst = max(worst, float(np.abs(a - b).max() / scale))
    return worst


for case, p in small_configs(qg).items():
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    for step in ("oml", "qgostep", "ocinvq", "ocqbdy"):
        getattr(gpu, step)()
        getattr(cpu, step)()
    r = {"centoc_cfraoc": scal(gpu, cpu, ("centoc", "cfraoc"))}
    if p.has("cyclic_ocean"):
        sc = cpu.get_scalars().as_dict()
        r["ocncs_ocncn"] = scal(gpu, cpu, ("ocncs", "ocncn"))
        fl2 = max(np.abs(np.atleast_1d(sc[g])).max() for g in ("ajisoc", "ajinoc", "ap5soc", "ap5noc"))
        r["ajis_ap5"] = scal(gpu, cpu, ("ajisoc", "ajinoc", "ap5soc", "ap5noc"), floor=fl2)
        fl3 = max(np.abs(np.atleast_1d(sc[g])).max() for g in ("enisoc", "eninoc"))
        r["enis"] = scal(gpu, cpu, ("enisoc", "eninoc"), floor=fl3)
    n = 100 * p.nstr
    gpu.run(1, n)
    cpu.run(1, n)
    r["drift100"] = max(rel_l2(gpu.get_field(f), cpu.get_field(f)) for f in ("po", "qo", "sst"))
    out[case] = r

for case, p in coupled_configs(qg).items():
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    for step in ("xforc", "oml", "qgostep", "ocinvq", "ocqbdy", "aml", "qgastep", "atinvq", "atqzbd"):
        getattr(gpu, step)()
        getattr(cpu, step)()
    r = {"cfraat_centat": scal(gpu, cpu, ("cfraat", "centat")), "atmcs_atmcn": scal(gpu, cpu, ("atmcs", "atmcn"))}
    gpu.run(1, 7)
    cpu.run(1, 7)
    r["seven_steps"] = max(rel_l2(gpu.get_field(f), cpu.get_field(f)) for f in OCEAN_CHECK + ATMOS_CHECK)
    gpu.run(8, 100)
    cpu.run(8, 100)
    r["drift100"] = max(rel_l2(gpu.get_field(f), cpu.get_field(f)) for f in ("po", "qo", "sst", "pa", "qa", "ast", "hmixa"))
    out[case] = r

for deck in ("dg_coupled", "so_coupled"):
    p = qg.named_config(deck)
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    n = p.nstr + 1
    gpu.run(1, n)
    cpu.run(1, n)
    out["full_" + deck] = {"fields": max(rel_l2(gpu.get_field(f), cpu.get_field(f)) for f in OCEAN_CHECK + ATMOS_CHECK + ("tauxo", "tauyo", "fnetoc"))}

print(json.dumps(out, indent=1))

"""Sweep the rows-per-march of the two marching kernels (QGCM_QG_MROWS / QGCM_OML_MROWS,
read by the library at launch time) on the NAtl 1 km grid and on a slab-sized grid of the
same width, and print the per-launch CUDA-event times of the library's own profile.

  python scripts/march_sweep.py [--nyaooc 120] [--qg 0,110,138,...] [--oml 0,166,...]
(0 = the library's own choice)"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def profile(m, steps):
    m._lib.qgcm_profile(m._h, 1)
    for _ in range(steps):
        m.ocean_step()
    buf = C.create_string_buffer(1 << 16)
    m._call("profile_report", buf, C.c_int64(len(buf)))
    m._lib.qgcm_profile(m._h, 0)
    out = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, tot = ln.split()
        out[name] = float(tot) / int(cnt)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nyaooc", type=int, default=120)
    ap.add_argument("--qg", default="0")
    ap.add_argument("--oml", default="0")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import _pkg
    qg = _pkg.load()
    p = qg.named_config("natl1km")
    if args.nyaooc != p.nyaooc:
        p = p.scaled(p.nxaooc, args.nyaooc, ndxr=p.ndxr, name="natl1km_ny%d" % (args.nyaooc * p.ndxr))
    cfg = qg.build_config(p, device=0)
    m = qg.Model(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    for _ in range(5):
        m.ocean_step()
    m.sync()
    for var, vals, kern in (("QGCM_QG_MROWS", args.qg, "k_qgstep"), ("QGCM_OML_MROWS", args.oml, "k_oml_step")):
        for v in [int(x) for x in vals.split(",")]:
            if v:
                os.environ[var] = str(v)
            else:
                os.environ.pop(var, None)
            profile(m, 3)
            t = profile(m, args.steps)
            print("%s ny=%d %s=%d  %s %.4f ms  (step kernels total %.4f ms)" % (p.name, p.nypo, var, v, kern, t[kern], sum(t.values())), flush=True)
        os.environ.pop(var, None)


if __name__ == "__main__":
    main()

"""helpers shared by the parity tests"""
import numpy as np

TOL = 1e-11   # BASELINE.json north_star: FP64 relative L2 after one step


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    if nb == 0.0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / nb)


def small_configs(qg):
    """reduced grids with the physics of each benchmark deck; every grid has > 2 chunks
    of the partitioned tridiagonal solve and a ragged last chunk"""
    box = qg.named_config("dg_oo").scaled(6, 5, ndxr=16, name="box_dg")          # 96 x 80
    box1 = qg.named_config("natl1km").scaled(3, 4, ndxr=40, name="box_natl1km")  # 120 x 160, nstr=1
    cyc = qg.named_config("so_coupled").scaled(6, 5, nxta=6, nyta=15, ndxr=16, name="chan_so")
    cyc.flags = ["ocean_only", "cyclic_ocean", "nb_hflux"]
    # 960 x 480: long enough for the four-pass TMA-fed DST plan (half length 480 = 240*2), and
    # 3*479 rows make every persistent block walk several rows
    fast = qg.named_config("natl1km").scaled(24, 12, ndxr=40, name="box_fast")
    return {"box_dg": box, "box_natl1km": box1, "chan_so": cyc, "box_fast": fast}


def make_pair(qg, pyorc, p, kind="random", seed=None):
    cfg = qg.build_config(p)
    gpu = qg.Model(cfg)
    cpu = pyorc.Oracle(cfg)
    kw = {} if seed is None else {"seed": seed}
    for m in (gpu, cpu):
        qg.synth.init_model(m, p, cfg, kind, **kw)
    return cfg, gpu, cpu


OCEAN_CHECK = ("po", "pom", "qo", "qom", "sst", "sstm", "entoc", "wekto", "wekpo")


def compare(gpu, cpu, names, tol=TOL, label=""):
    bad = []
    for n in names:
        e = rel_l2(gpu.get_field(n), cpu.get_field(n))
        if not e <= tol:
            bad.append((n, e))
    assert not bad, "%s fields beyond %.1e: %s" % (label, tol, bad)


def compare_scalars(gpu, cpu, names, tol=1e-9, floor=None):
    """floor: absolute scale for quantities that are sums with heavy cancellation (an area
    integral of a zero-mean field is rounding noise; compare it against the integral of
    the magnitude instead)"""
    sg, sc = gpu.get_scalars().as_dict(), cpu.get_scalars().as_dict()
    bad = []
    for n in names:
        a, b = np.atleast_1d(sg[n]).astype(float), np.atleast_1d(sc[n]).astype(float)
        scale = max(np.abs(b).max(), 1e-300)
        if floor is not None:
            scale = max(scale, floor)
        if not np.abs(a - b).max() <= tol * scale:
            bad.append((n, a.tolist(), b.tolist()))
    assert not bad, "scalars differ: %s" % bad


def integral_scale(m, p, name="po"):
    """dx*dy * sum |field|: the magnitude an area integral of that field is rounded against"""
    return float(np.abs(m.get_field(name)).sum() * p.dxo ** 2)

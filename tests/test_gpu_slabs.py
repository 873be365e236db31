"""-m gpu: the y-slab multi-GPU path on ONE device.  qgcm_group_create puts
every rank of a partition in this process (same kernels, same exchange pattern; device
copies stand in for NCCL), so the whole N > 1 algorithm -- halo exchange, the slab level of
the tridiagonal partition, the all-reduced constraint integrals -- is checked against the
CPU oracle and against the single-slab run on the single-GPU test box."""
import numpy as np
import pytest

from util import TOL, rel_l2, small_configs, compare, compare_scalars, integral_scale, OCEAN_CHECK

pytestmark = pytest.mark.gpu

# (deck, ranks): 81 p rows in 2, 3, 4 slabs (one chunk per slab, ragged), 481 rows in 2 and 8
# slabs (several chunks per slab, the four-... three-pass DST plan)
CASES = [("box_dg", 2), ("box_dg", 3), ("box_dg", 4), ("box_natl1km", 2), ("box_fast", 2), ("box_fast", 8)]


def make_group(qg, pyorc, p, nranks, kind="random"):
    cfg = qg.build_config(p)
    grp = qg.SlabGroup(cfg, nranks)
    cpu = pyorc.Oracle(cfg)
    for m in (grp, cpu):
        qg.synth.init_model(m, p, cfg, kind)
    return cfg, grp, cpu


def test_slab_bounds_cover_the_grid(qg):
    for nyp in (81, 161, 481, 2401, 4801):
        for n in (1, 2, 3, 4, 5, 8):
            edges = [qg.slab_bounds(nyp, n, r) for r in range(n)]
            assert edges[0][0] == 0 and sum(e[1] for e in edges) == nyp
            for a, b in zip(edges, edges[1:]):
                assert a[0] + a[1] == b[0]
            assert max(e[1] for e in edges) - min(e[1] for e in edges) <= 1


@pytest.mark.parametrize("case,nranks", CASES)
def test_slab_init_sequence(qg, pyorc, case, nranks):
    p = small_configs(qg)[case]
    cfg, grp, cpu = make_group(qg, pyorc, p, nranks)
    compare(grp, cpu, ("qo", "qom", "wekto", "wekpo", "ochom"), label="%s/%d" % (case, nranks))
    compare_scalars(grp, cpu, ("dpioc", "dpiocp", "aipohs", "cdiffo", "cdhoc"))


@pytest.mark.parametrize("case,nranks", CASES)
def test_slab_steps_match_oracle_and_single_gpu(qg, pyorc, case, nranks):
    p = small_configs(qg)[case]
    cfg, grp, cpu = make_group(qg, pyorc, p, nranks)
    one = qg.Model(cfg)
    qg.synth.init_model(one, p, cfg, "random")
    n = 3 * p.nstr + 1
    for m in (grp, cpu, one):
        m.run(1, n)           # ocean steps, time-level average at nt = 1
    compare(grp, cpu, OCEAN_CHECK, label="%s/%d vs oracle" % (case, nranks))
    compare(grp, one, OCEAN_CHECK, label="%s/%d vs one slab" % (case, nranks))
    fl = integral_scale(cpu, p)
    compare_scalars(grp, cpu, ("dpioc", "dpiocp", "xinhom_oc"), tol=1e-11, floor=fl)
    compare_scalars(grp, cpu, ("xon",), tol=1e-11, floor=integral_scale(cpu, p, "entoc"))
    compare_scalars(grp, cpu, ("centoc", "cfraoc"), tol=1e-12)      # sums over ranks in rank order: measured <= 1e-15


@pytest.mark.parametrize("case,nranks", [("box_dg", 3), ("box_fast", 8)])
def test_slab_hundred_steps_drift(qg, pyorc, case, nranks):
    p = small_configs(qg)[case]
    cfg, grp, cpu = make_group(qg, pyorc, p, nranks)
    n = 100 * p.nstr
    grp.run(1, n)
    cpu.run(1, n)
    for name in ("po", "qo", "sst"):
        a = grp.get_field(name)
        assert np.isfinite(a).all()
        assert rel_l2(a, cpu.get_field(name)) <= 1e-10, (case, nranks, name)      # as test_hundred_steps_drift


def test_slab_rejects_unsupported_decks(qg):
    p = small_configs(qg)["chan_so"]
    cfg = qg.build_config(p)
    with pytest.raises(RuntimeError):
        qg.Model(qg.slab_config(cfg, 2, 0))

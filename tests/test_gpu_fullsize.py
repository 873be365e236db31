"""-m gpu: parity at the sizes the benchmark quotes its numbers on (BASELINE.json configs 4-5):
NAtl 2 km (2401^2 x 3, src/parameters_data.F.NAtl.2km:45,50) and NAtl 1 km (4801^2 x 3,
src/parameters_data.F.NAtl.1km:45,50), single GPU and y-slab partitions, against the CPU
oracle on identical seeded inputs.  Tolerance: FP64 relative L2 <= 1e-11 per field
(BASELINE.json north_star).  One oracle run per deck is shared by the single-GPU model and
the partitions (a 1 km oracle step costs ~0.3 s on 16 cores, its start-up a few seconds)."""
import numpy as np
import pytest

from util import TOL, rel_l2, compare, compare_scalars, integral_scale, OCEAN_CHECK

pytestmark = pytest.mark.gpu


def _run_deck(qg, pyorc, deck, slab_counts, nsteps):
    p = qg.named_config(deck)
    cfg = qg.build_config(p)
    cpu = pyorc.Oracle(cfg)
    qg.synth.init_model(cpu, p, cfg, "random")
    n = (nsteps - 1) * p.nstr + 1       # ocean steps at nt = 1, 1+nstr, ...; time-level average at nt = 1
    cpu.run(1, n)
    ref = {k: cpu.get_field(k) for k in OCEAN_CHECK}
    sref = cpu.get_scalars().as_dict()
    fl = integral_scale(cpu, p)
    fle = integral_scale(cpu, p, "entoc")
    cpu.close()

    class Frozen:      # the oracle's answer, kept while the oracle's 6 GB are released
        def get_field(self, k, shape=None):
            return ref[k]

        def get_scalars(self):
            class S:
                def as_dict(_):
                    return sref
            return S()

    frozen = Frozen()
    models = [("one GPU", lambda: qg.Model(cfg))] + [("%d slabs" % s, (lambda s=s: qg.SlabGroup(cfg, s))) for s in slab_counts]
    for label, make in models:
        m = make()
        qg.synth.init_model(m, p, cfg, "random")
        m.run(1, n)
        compare(m, frozen, OCEAN_CHECK, label="%s, %s" % (deck, label))
        compare_scalars(m, frozen, ("dpioc", "dpiocp", "xinhom_oc"), tol=1e-11, floor=fl)
        compare_scalars(m, frozen, ("xon",), tol=1e-11, floor=fle)
        for nm in ("po", "qo", "sst"):
            assert np.isfinite(m.get_field(nm)).all()
        m.close()


def test_natl1km_full_size_one_gpu_and_8_slabs(qg, pyorc):
    """the bench workload itself: 4801 x 4801 x 3, k_dst3<10,*> inside a full step, 150
    chunks per wavenumber in the interface system, 8 slabs of 600 rows"""
    _run_deck(qg, pyorc, "natl1km", (8,), nsteps=2)


def test_natl2km_full_size_one_gpu_and_slabs(qg, pyorc):
    """BASELINE config 4: 2401 x 2401 x 3 at 1/2/4/8 slabs"""
    _run_deck(qg, pyorc, "natl2km", (2, 4, 8), nsteps=2)

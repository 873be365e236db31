"""-m gpu: device-side qocdiag_out (src/qocdiag.F:303-683) against the CPU oracle through the C
ABI: all five terms at the sub-sampled points, box and channel, and over y-slabs."""
import numpy as np
import pytest

from util import TOL, rel_l2, small_configs, make_pair

pytestmark = pytest.mark.gpu
TERMS = ("dqdt", "qotjac", "qt2dif", "qt4dif", "qotent")


def _same_inputs(gpu, cpu):
    """the diagnostic differentiates pom six times and differences qo in time on the walls, so
    the 1e-12 the two stepped states differ by would be amplified past the parity bar: the
    kernel is checked on the oracle's own fields"""
    for name in ("po", "pom", "qo", "qom", "wekpo", "entoc"):
        gpu.set_field(name, cpu.get_field(name))


def _check(got, want, shape, label):
    got, want = got.reshape(shape, order="F"), want.reshape(shape, order="F")
    assert np.isfinite(got).all(), label
    scale = np.linalg.norm(want[..., 0])          # the terms cancel in part: one scale, that of dqdt
    for t, name in enumerate(TERMS):
        e = np.linalg.norm(got[..., t] - want[..., t]) / scale
        assert e <= 1e-13, (label, name, e)       # unfused arithmetic on identical inputs


@pytest.mark.parametrize("case", ["box_dg", "box_natl1km", "chan_so", "box_fast"])
@pytest.mark.parametrize("nsk", [1, 4])
def test_qocdiag_terms(qg, pyorc, case, nsk):
    p = small_configs(qg)[case]
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    for m in (gpu, cpu):
        m.ocean_step()
        m.oml()                # the reference's call site: between oml and qgostep (src/q-gcm.F:1237)
    _same_inputs(gpu, cpu)
    iw, jw = -(-p.nxpo // nsk), -(-p.nypo // nsk)
    _check(gpu.qocdiag(nsk), cpu.qocdiag(nsk), (iw, jw, p.nlo, 5), "%s nsk=%d" % (case, nsk))
    # the diagnostic borrows the solver's work array: the step that follows must be unaffected
    for m in (gpu, cpu):
        m.qgostep(); m.ocinvq(); m.ocqbdy()
    for name in ("po", "qo"):
        assert rel_l2(gpu.get_field(name), cpu.get_field(name)) <= TOL, (case, name)


@pytest.mark.parametrize("nranks", [2, 4])
def test_qocdiag_over_slabs(qg, pyorc, nranks):
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    grp = qg.SlabGroup(cfg, nranks)
    cpu = pyorc.Oracle(cfg)
    for m in (grp, cpu):
        qg.synth.init_model(m, p, cfg, "random")
        m.ocean_step()
    _same_inputs(grp, cpu)
    for nsk in (1, 3):
        iw, jw = -(-p.nxpo // nsk), -(-p.nypo // nsk)
        _check(grp.qocdiag(nsk), cpu.qocdiag(nsk), (iw, jw, p.nlo, 5), "slabs=%d nsk=%d" % (nranks, nsk))
    for m in (grp, cpu):
        m.ocean_step()
    for name in ("po", "qo", "sst"):
        assert rel_l2(grp.get_field(name), cpu.get_field(name)) <= TOL, (nranks, name)

"""-m gpu: device-side running sums (tavocn, tavatm, avg_ocn_k247; src/timavge.F) and the packed
sub-sampled read of ocnc_out / atnc_out (src/nc_subs.F:869-890) against the CPU oracle, through
the C ABI.  Every sum is one add per contribution, so the bar is the
step's own 1e-11 on sums accumulated over stepped states and 1e-14 on a single contribution."""
import numpy as np
import pytest

from util import TOL, rel_l2, small_configs, make_pair, compare
from test_gpu_parity import coupled_configs

pytestmark = pytest.mark.gpu

OC_SUMS = ("txocav", "tyocav", "wpocav", "wtocav", "fmocav", "sstav", "uufo", "tufo", "utufo", "vvfo", "tvfo", "vtvfo",
           "pocav", "qocav")
AT_SUMS = ("txatav", "tyatav", "wtatav", "fmatav", "astav", "uufa", "tufa", "utufa", "vvfa", "tvfa", "vtvfa",
           "patav", "qatav")


@pytest.mark.parametrize("case", ["box_dg", "box_natl1km", "chan_so", "box_fast"])
def test_tavocn_and_k247_sums(qg, pyorc, case):
    p = small_configs(qg)[case]
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    for m in (gpu, cpu):
        m.tavini()
        m.tavocn()
        m.avg_ocn_k247()
    compare(gpu, cpu, OC_SUMS + ("po_avg",), tol=1e-14, label=case + " first contribution")
    for m in (gpu, cpu):
        for _ in range(2):
            m.ocean_step()
            m.avg_ocn_k247()
        m.tavocn()
    compare(gpu, cpu, OC_SUMS + ("po_avg",), tol=TOL, label=case)
    assert gpu.tav_counts() == cpu.tav_counts() == (0, 2, 3)
    for m in (gpu, cpu):
        m.tavini()
    assert not gpu.get_field("utufo").any() and not gpu.get_field("po_avg").any()
    assert gpu.tav_counts() == (0, 0, 0)


@pytest.mark.parametrize("case", ["cpl_dg", "cpl_so"])
def test_tavatm_and_tavocn_coupled(qg, pyorc, case):
    p = coupled_configs(qg)[case]
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    for m in (gpu, cpu):
        m.tavatm()              # sums allocate on first use
        m.tavocn()
    # 1e-13: the sums are copies of fields the start-up xforc produced, and its separable stress kernel
    # associates the 16-term interpolation sums differently from the oracle (measured 1.1e-14 on wtocav)
    compare(gpu, cpu, AT_SUMS + OC_SUMS, tol=1e-13, label=case + " first contribution")
    for m in (gpu, cpu):
        m.run(1, p.nstr + 1)
        m.tavatm()
        m.tavocn()
    compare(gpu, cpu, AT_SUMS + OC_SUMS, tol=TOL, label=case)
    assert gpu.tav_counts() == cpu.tav_counts() == (2, 2, 0)


def test_k247_flag_accumulates_inside_run(qg, pyorc):
    """-Docnc_avg_k247: the main loop adds po after every ocean step (src/q-gcm.F:1250-1252)"""
    p = small_configs(qg)["box_dg"]
    p.flags = list(p.flags) + ["ocnc_avg_k247"]
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    gpu.run(1, 2 * p.nstr)
    for nt in range(1, 2 * p.nstr + 1):     # the ocean-only main loop, spelled out (src/q-gcm.F:1222-1366)
        if nt % p.nstr == 1:
            cpu.ocean_step()
            cpu.avg_ocn_k247()              # before the time-level average of the same nt
        if (nt - 1) % (25 * p.nstr) == 0:
            cpu.tlavg_ocean()
    assert gpu.tav_counts()[2] == cpu.tav_counts()[2] == 2
    compare(gpu, cpu, ("po_avg",), tol=TOL)


@pytest.mark.parametrize("nranks", [2, 3])
def test_sums_over_slabs(qg, pyorc, nranks):
    p = small_configs(qg)["box_dg"]
    cfg = qg.build_config(p)
    grp = qg.SlabGroup(cfg, nranks)
    cpu = pyorc.Oracle(cfg)
    for m in (grp, cpu):
        qg.synth.init_model(m, p, cfg, "random")
        m.ocean_step()
        m.tavocn()
        m.avg_ocn_k247()
        m.ocean_step()
        m.tavocn()
        m.avg_ocn_k247()
    for name in OC_SUMS + ("po_avg",):
        a, b = grp.get_field(name), cpu.get_field(name)
        assert np.isfinite(a).all(), name       # every row is owned by some rank
        assert rel_l2(a, b) <= TOL, (nranks, name)
    for name, nsk in (("po", 4), ("sst", 3), ("qocav", 5)):
        shape = (p.nxto, p.nyto) if name == "sst" else (p.nxpo, p.nypo, p.nlo)
        got = grp.get_field_sub(name, nsk)
        assert np.array_equal(got, pyorc.subsample(grp.get_field(name, shape), nsk)), (nranks, name)


def test_run_accumulates_po_avg_on_every_slab(qg, pyorc):
    """qgcm_run with -Docnc_avg_k247 on a loopback partition: every rank adds its rows to po_avg
    (src/q-gcm.F:1250-1252), not only the rank the call was made on"""
    p = small_configs(qg)["box_dg"]
    p.flags = list(p.flags) + ["ocnc_avg_k247"]
    cfg = qg.build_config(p)
    grp = qg.SlabGroup(cfg, 3)
    cpu = pyorc.Oracle(cfg)
    for m in (grp, cpu):
        qg.synth.init_model(m, p, cfg, "random")
    grp.run(1, 2 * p.nstr)
    for nt in range(1, 2 * p.nstr + 1):
        if nt % p.nstr == 1:
            cpu.ocean_step()
            cpu.avg_ocn_k247()
        if (nt - 1) % (25 * p.nstr) == 0:
            cpu.tlavg_ocean()
    a, b = grp.get_field("po_avg"), cpu.get_field("po_avg")
    assert np.isfinite(a).all()
    assert rel_l2(a, b) <= TOL
    assert float(np.abs(a.reshape((p.nxpo, p.nypo, p.nlo), order="F")[:, -5:, 0]).max()) > 0.0   # the last rank's rows too


@pytest.mark.parametrize("nsk", [1, 2, 3, 7, 16])
def test_subsampled_read(qg, pyorc, nsk):
    """bit-exact: the packed vector is a copy of every nsk-th point"""
    p = coupled_configs(qg)["cpl_dg"]
    cfg, gpu, cpu = make_pair(qg, pyorc, p)
    gpu.run(1, p.nstr)
    shapes = {"po": (p.nxpo, p.nypo, p.nlo), "qo": (p.nxpo, p.nypo, p.nlo), "sst": (p.nxto, p.nyto),
              "wekto": (p.nxto, p.nyto), "tauxo": (p.nxpo, p.nypo), "pa": (p.nxta + 1, p.nyta + 1, p.nla),
              "ast": (p.nxta, p.nyta), "hmixa": (p.nxta, p.nyta)}
    for name, shape in shapes.items():
        got = gpu.get_field_sub(name, nsk)
        want = pyorc.subsample(gpu.get_field(name, shape), nsk)
        assert got.shape == want.shape and np.array_equal(got, want), (name, nsk)
    with pytest.raises(RuntimeError):
        gpu.get_field_sub("sstbar", 2)      # not a gridded field

"""-m gpu: examples/host_loop.c -- a compiled (C99) host that drives the C ABI the way the
Fortran main program would -- against the CPU oracle running the same loop.
The binary sees nothing but include/qgcm_b200.h, libqgcm_b200.so and raw files."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from util import TOL, rel_l2, small_configs
from test_gpu_parity import coupled_configs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler on this box")
    libdir = os.path.join(ROOT, "q-gcm_b200", "csrc")
    exe = str(tmp_path / "host_loop")
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-D_POSIX_C_SOURCE=200112L", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "host_loop.c"), "-L" + libdir, "-lqgcm_b200", "-Wl,-rpath," + libdir,
                           "-o", exe])
    return exe


@pytest.mark.parametrize("case", ["box_dg", "chan_so", "cpl_dg"])
def test_compiled_host_loop_matches_oracle(qg, pyorc, tmp_path, case):
    p = coupled_configs(qg)[case] if case.startswith("cpl") else small_configs(qg)[case]
    cfg = qg.build_config(p)
    exe = _build(tmp_path)
    d = tmp_path / "run"
    d.mkdir()
    with open(d / "config.bin", "wb") as f:
        f.write(bytes(cfg))
    amp = min(1.0, (p.nxto * p.dxo) / 4.8e6 * 4.0)
    st = dict(qg.synth.ocean_state(p, cfg, "random", qg.synth.SEED, amp))
    if not p.has("ocean_only"):
        st.update(qg.synth.atmos_state(p, cfg, "random", qg.synth.SEED + 1))
    with open(d / "fields.txt", "w") as f:
        for name, arr in st.items():
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).ravel(order="F"))
            a.tofile(str(d / (name + ".f64")))
            f.write("%s %d\n" % (name, a.size))
    nt_last = 3 * p.nstr + 1
    r = subprocess.run([exe, str(d), str(nt_last)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr

    cpu = pyorc.Oracle(cfg)                        # the same loop on the oracle
    qg.synth.init_model(cpu, p, cfg, "random")
    cpu.tavini()
    for nt in range(1, nt_last + 1):
        if nt % p.nstr == 1 or p.nstr == 1:
            if not p.has("ocean_only"):
                cpu.xforc()
            cpu.ocean_step()
            cpu.avg_ocn_k247()
        if not p.has("ocean_only"):
            cpu.atmos_step()
        if (nt - 1) % (25 * p.nstr) == 0:
            cpu.tlavg_ocean()
        if not p.has("ocean_only") and (nt - 1) % 100 == 0:
            cpu.tlavg_atmos()
        if nt % p.nstr == 0:
            cpu.tavocn()
    names = ["po", "qo", "sst", "po_avg", "pocav"] + ([] if p.has("ocean_only") else ["pa", "ast"])
    for name in names:
        got = np.fromfile(str(d / ("out_%s.f64" % name)))
        assert rel_l2(got, cpu.get_field(name)) <= TOL, (case, name)
    words = r.stdout.split()
    assert int(words[1]) == cpu.tav_counts()[1] and int(words[3]) == cpu.tav_counts()[2]
    mon = cpu.monnc_ocean().as_dict()
    assert np.isclose(float(words[5]), mon["kealoc"][0], rtol=1e-9)
    assert np.isclose(float(words[9]), mon["cnmloc"], rtol=1e-9)

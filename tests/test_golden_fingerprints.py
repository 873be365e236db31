"""The oracle against its committed fingerprints (tests/golden/oracle_fingerprints.json, made by
tests/golden/make_golden.py).  These are regression values of the oracle itself, not reference
outputs: they keep every parity target from moving when oracle/ or the synthetic states are
edited.  Sums are compared against the sum of squares they are rounded against (OpenMP
reductions inside the oracle depend on the thread count at the 1e-16 level)."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

with open(os.path.join(HERE, "golden", "oracle_fingerprints.json")) as _f:
    GOLD = json.load(_f)


def _check(case, now):
    for name, want in GOLD[case].items():
        got = now[name]
        if name == "monitor":
            for k, v in want.items():
                assert np.allclose(got[k], v, rtol=1e-9, atol=0.0), (case, k)
            continue
        scale = np.sqrt(want["sumsq"] * want["n"]) + 1e-300
        assert got["n"] == want["n"]
        assert abs(got["sum"] - want["sum"]) <= 1e-10 * scale, (case, name)
        assert abs(got["sumsq"] - want["sumsq"]) <= 1e-10 * want["sumsq"] + 1e-300, (case, name)
        amp = np.sqrt(want["sumsq"] / want["n"]) + 1e-300
        assert np.allclose(got["samples"], want["samples"], rtol=0.0, atol=1e-9 * amp), (case, name)


@pytest.mark.parametrize("case", sorted(GOLD))
def test_oracle_reproduces_its_fingerprints(qg, pyorc, case):
    _check(case, make_golden.run_case(qg, pyorc, make_golden.cases(qg)[case]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(GOLD))
def test_cuda_path_matches_the_committed_fingerprints(qg, pyorc, case):
    """the CUDA path against the committed fixtures alone: no oracle runs in this test"""
    _check(case, make_golden.run_case(qg, pyorc, make_golden.cases(qg)[case], make=qg.Model))


def test_oracle_work_arrays_carry_nothing_between_calls():
    """the oracle's persistent work arrays (the reference's automatic arrays) poisoned with NaN
    before every use: the fingerprints must not move"""
    import subprocess
    code = (
        "import sys, json\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import make_golden, _pkg, pyorc\n"
        "qg = _pkg.load()\n"
        "print(json.dumps({c: make_golden.run_case(qg, pyorc, make_golden.cases(qg)[c]) for c in ('box_dg', 'chan_so', 'cpl_dg')}))\n"
    ) % (os.path.join(HERE, "golden"), os.path.dirname(HERE))
    env = dict(os.environ, ORC_POISON="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    now = json.loads(r.stdout.strip().splitlines()[-1])
    for case, vals in now.items():
        _check(case, vals)

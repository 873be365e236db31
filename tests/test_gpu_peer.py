"""-m gpu: the peer-memory transport of the y-slab path (CUDA IPC mailboxes, epoch flags, the
exchanges issued by the kernels of the step) with ONE PROCESS PER RANK, as on the 8-GPU box.
The ranks use the visible GPUs round-robin; on the single-GPU test box they are separate
processes time-sliced on cuda:0, which exercises the same kernels and the same flag protocol
(scripts/peer_worker.py compares every rank's owned rows with the CPU oracle, 1e-11)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_ranks(case, nranks, steps):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "scripts", "peer_worker.py"), "--case", case, "--steps", str(steps)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    # the ranks share one stdout pipe: two reports can land on one line, so count the markers
    return r, (r.returncode == 0 and r.stdout.count("PEER_OK rank") == nranks)


@pytest.mark.parametrize("case,nranks,steps", [("box_dg", 2, 3), ("box_dg", 3, 2), ("box_fast", 2, 3), ("box_fast", 4, 2)])
def test_peer_transport_matches_oracle(case, nranks, steps):
    # the launcher's rendezvous (a TCP port picked here, then re-bound by torchrun) can lose a
    # race with another process on a busy box: one retry, and the first attempt's output is kept
    r, good = run_ranks(case, nranks, steps)
    if not good:
        log = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(log):
            with open(os.path.join(log, "peer_fail_%s_%d.log" % (case, nranks)), "w") as f:
                f.write("rc=%d\n%s\n%s" % (r.returncode, r.stdout, r.stderr))
        if "differs from the oracle" in r.stdout + r.stderr:
            pytest.fail("parity: %s" % (r.stdout + r.stderr)[-2000:])
        r, good = run_ranks(case, nranks, steps)
    assert good, "rc=%d\n%s\n%s" % (r.returncode, r.stdout[-3000:], r.stderr[-3000:])

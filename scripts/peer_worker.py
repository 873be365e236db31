"""One rank of a y-slab partition that talks to the others through the peer-memory transport
(qgcm_peer_handle / qgcm_comm_init_peer: CUDA IPC mailboxes, include/qgcm_b200.h).

Launched once per rank, e.g.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29517 scripts/peer_worker.py --case box_fast --steps 4
The ranks use the GPUs round-robin, so on a single-GPU box every rank is a separate process
on cuda:0 (the driver time-slices them; slow, but it exercises exactly the code that runs
over NVLink).  torch.distributed (gloo) only carries the 64-byte handles.  Every rank runs the
CPU oracle on the whole domain and compares the rows it owns; exit code 0 and a line
"PEER_OK" per rank mean parity within 1e-11."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="box_dg")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--also-nccl", action="store_true", help="repeat over NCCL and compare (needs one GPU per rank)")
    ap.add_argument("--stress", type=int, default=0, help="repeat the peer run this many times and count failures")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import _pkg
    import pyorc
    from util import small_configs, rel_l2, TOL, OCEAN_CHECK

    qg = _pkg.load()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    if rank == 0:
        pyorc.build()
    dist.barrier()

    p = small_configs(qg)[args.case]
    cfg = qg.build_config(p)
    cfg.device = dev

    def make(kind):
        m = qg.Model(qg.slab_config(cfg, world, rank))
        if kind == "peer":
            mine = m.peer_handle()
            allh = [None] * world
            dist.all_gather_object(allh, mine)
            m.comm_init_peer(allh)
        else:
            ident = [qg.Model.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ident, 0)
            m.comm_init_nccl(ident[0])
        dist.barrier()
        qg.synth.init_model(m, p, cfg, "random")
        return m

    cpu = pyorc.Oracle(cfg)
    qg.synth.init_model(cpu, p, cfg, "random")
    j0, nown = qg.slab_bounds(p.nypo, world, rank)

    def owned(name, a):
        n = a.size
        nyp, nxp = p.nypo, p.nxpo
        for ny, nx in ((nyp, nxp), (nyp - 1, nxp - 1)):
            if n % (nx * ny) == 0:
                a = a.reshape((nx, ny, n // (nx * ny)), order="F")
                return a[:, j0:min(j0 + nown, ny), :]
        raise RuntimeError("unexpected field size for %s" % name)

    def check(m, label):
        worst = 0.0
        for name in OCEAN_CHECK:
            e = rel_l2(owned(name, m.get_field(name)), owned(name, cpu.get_field(name)))
            worst = max(worst, e)
            if not e <= TOL:
                raise RuntimeError("rank %d %s: %s differs from the oracle: %.3e" % (rank, label, name, e))
        return worst

    if args.stress:
        n = args.steps * p.nstr + 1
        cpu.run(1, n)
        bad = []
        for it in range(args.stress):
            m = make("peer")
            dist.barrier()
            try:
                for nt in range(1, n + 1):
                    m.run(nt, nt)
                    m.sync()
                    nf = [k for k in ("sst", "entoc", "qo", "po", "pom") if not np.isfinite(owned(k, m.get_field(k))).all()]
                    if nf:
                        sc = m.get_scalars().as_dict()
                        raise RuntimeError("nt=%d non-finite %s xon=%s dpioc=%s xinhom=%s" % (nt, nf, sc["xon"][:1], sc["dpioc"][:2], sc["xinhom_oc"][:3]))
                check(m, "stress %d" % it)
            except RuntimeError as ex:
                bad.append((it, str(ex)[-200:]))
            dist.barrier()
            m.comm_close_peer()
            dist.barrier()
            m.close()
        print("STRESS rank %d: %d/%d failed %s" % (rank, len(bad), args.stress, bad[:4]), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return

    m = make("peer")
    w0 = check(m, "peer init")
    n = args.steps * p.nstr + 1
    dist.barrier()
    m.run(1, n)
    m.sync()
    cpu.run(1, n)
    w1 = check(m, "peer steps")
    msg = "PEER_OK rank %d/%d dev %d %s: init %.2e, %d steps %.2e" % (rank, world, dev, args.case, w0, args.steps, w1)
    if args.also_nccl:
        m2 = make("nccl")
        dist.barrier()
        m2.run(1, n)
        m2.sync()
        w2 = check(m2, "nccl steps")
        d = max(rel_l2(owned(k, m.get_field(k)), owned(k, m2.get_field(k))) for k in ("po", "qo", "sst"))
        msg += "; nccl %.2e, peer-vs-nccl %.2e" % (w2, d)
        dist.barrier()
        m2.close()
    print(msg, flush=True)
    dist.barrier()      # no rank unmaps a mailbox another rank may still be writing to
    m.comm_close_peer()
    dist.barrier()      # no rank frees a mailbox another rank still maps
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU tests of the host-side logic of the y-slab multi-GPU path, world_size 2 and 3 over gloo.

* qgcm_slab_bounds (pure host arithmetic in the library) tiles the grid;
* the slab level of the tridiagonal partition, restated in numpy with a gloo all_gather in
  the place of the library's ncclAllGather: every rank solves its rows of the global
  constant-coefficient system from (a) its local zero-neighbour solve, (b) the gathered
  first/last rows of all slabs, (c) the redundant 2*nranks-unknown system per wavenumber --
  the algorithm of helmholtz.cu (k_slab_spikes / k_slab_fg / k_slab_solve);
* the owned-row integral shares add up to xintp of the global field (intsubs.f:78-133).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.linalg import solve_banded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _thomas(b, a, rhs):
    """solve tridiag(a, b, a) u = rhs column-wise; rhs [n, ncol], b [ncol]"""
    n = rhs.shape[0]
    ab = np.zeros((3, n))
    out = np.empty_like(rhs)
    for c in range(rhs.shape[1]):
        ab[0, 1:] = a
        ab[1, :] = b[c]
        ab[2, :-1] = a
        out[:, c] = solve_banded((1, 1), ab, rhs[:, c])
    return out


def _worker(rank, world, port, nyp, ncol, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        import _pkg
        qg = _pkg.load()
        rng = np.random.default_rng(11)                   # same global problem on every rank
        a = 1.0
        b = -(2.0 + rng.uniform(1e-4, 2.0, ncol))          # from nearly singular (low wavenumber) to dominant
        rhs = rng.standard_normal((nyp, ncol))
        rhs[0] = rhs[-1] = 0.0
        want = np.zeros_like(rhs)
        want[1:-1] = _thomas(b, a, rhs[1:-1])
        # rows this rank solves: owned rows minus the walls
        j0, n = qg.slab_bounds(nyp, world, rank)
        lo, hi = max(j0, 1), min(j0 + n, nyp - 1)
        rows = [(max(s, 1), min(s + m, nyp - 1)) for s, m in (qg.slab_bounds(nyp, world, r) for r in range(world))]
        # (a) local solve with zero neighbours; first/last rows
        loc = _thomas(b, a, rhs[lo:hi])
        mine = torch.from_numpy(np.stack([loc[0], loc[-1]]))
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        fg = [g.numpy() for g in gathered]
        # (b) spikes of every slab: response to a unit neighbour below
        ae = []
        for (s, e) in rows:
            unit = np.zeros((e - s, ncol))
            unit[0] = -a
            v = _thomas(b, a, unit)
            ae.append((v[0], v[-1]))
        # (c) the inter-slab system, solved redundantly
        yprev = np.zeros(ncol)
        xnext = np.zeros(ncol)
        for c in range(ncol):
            m = np.zeros((2 * world, 2 * world))
            r_ = np.zeros(2 * world)
            for r in range(world):
                al, ep = ae[r][0][c], ae[r][1][c]
                m[2 * r, 2 * r] = m[2 * r + 1, 2 * r + 1] = 1.0
                if r > 0:
                    m[2 * r, 2 * r - 1] = -al
                    m[2 * r + 1, 2 * r - 1] = -ep
                if r < world - 1:
                    m[2 * r, 2 * r + 2] = -ep
                    m[2 * r + 1, 2 * r + 2] = -al
                r_[2 * r], r_[2 * r + 1] = fg[r][0][c], fg[r][1][c]
            z = np.linalg.solve(m, r_)
            yprev[c] = z[2 * rank - 1] if rank > 0 else 0.0
            xnext[c] = z[2 * rank + 2] if rank < world - 1 else 0.0
        # final rows: neighbours moved to the right-hand side
        mod = rhs[lo:hi].copy()
        mod[0] -= a * yprev
        mod[-1] -= a * xnext
        got = _thomas(b, a, mod)
        err = np.abs(got - want[lo:hi]).max() / np.abs(want).max()
        # owned-row shares of the p-grid integral
        w = np.ones(nyp)
        w[0] = w[-1] = 0.5
        share = torch.tensor([float((w[j0:j0 + n, None] * want[j0:j0 + n]).sum())], dtype=torch.float64)
        dist.all_reduce(share)
        tot = float((w[:, None] * want).sum())
        q.put((rank, err, abs(share.item() - tot) / max(abs(tot), 1e-300) if tot != 0 else 0.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nyp", [(2, 81), (3, 100)])
def test_slab_partitioned_tridiagonal_over_gloo(world, nyp):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nyp, 24, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, ierr in res:
        assert err <= 1e-10, (rank, err)      # low wavenumbers are ill conditioned: cond ~ 1e4 here
        assert ierr <= 1e-10, (rank, ierr)


def test_slab_bounds_host_only():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import _pkg
    qg = _pkg.load()
    assert qg.slab_bounds(4801, 8, 0) == (0, 601)
    assert qg.slab_bounds(4801, 8, 7) == (4201, 600)
    assert sum(qg.slab_bounds(4801, 8, r)[1] for r in range(8)) == 4801


def _monitor_worker(rank, world, port, nyp, nx, q):
    """the share algebra of qgcm_monnc_ocean on a y-slab partition (monitor.cu: mon_share /
    mon_finish), restated in numpy over gloo: per sum slot the interior-row sum a rank owns plus
    the global south / north row values if it owns them, combined by ONE all-reduce(sum);
    extrema and jet candidates travel in a one-hot block per rank of the same vector"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        import _pkg
        qg = _pkg.load()
        rng = np.random.default_rng(5)
        nyt = nyp - 1
        pfield = rng.standard_normal((nx, nyp))          # a p-grid integrand (weights 0.5 W/E, 0.5 S/N)
        tfield = rng.standard_normal((nx, nyt))          # a T-row integrand (u points: weights 0.5 W/E, 1 S/N)
        ujet = rng.standard_normal(nyt)
        ujet[[3, nyt - 4]] = 9.0                         # a tie: the first row must win (src/monitor_diag.F:696-704)
        wx = np.ones(nx); wx[0] = wx[-1] = 0.5
        j0, n = qg.slab_bounds(nyp, world, rank)
        vec = np.zeros(6 + 4 * world)
        for slot, (f, ny) in enumerate(((pfield, nyp), (tfield, nyt))):
            rows = wx @ f                                # what the row kernels produce
            j1 = min(j0 + n, ny)
            for j in range(j0, j1):
                if j == 0:
                    vec[3 * slot + 1] = rows[j]
                elif j == ny - 1:
                    vec[3 * slot + 2] = rows[j]
                else:
                    vec[3 * slot] += rows[j]
        own = slice(j0, min(j0 + n, nyt))
        g = 6 + 4 * rank
        vec[g], vec[g + 1] = tfield[:, own].min(), tfield[:, own].max()
        best, brow = 0.0, 0
        for j in range(own.start, own.stop):
            if abs(ujet[j]) > best:
                best, brow = abs(ujet[j]), j + 1
        vec[g + 2], vec[g + 3] = best, brow
        t = torch.from_numpy(vec)
        dist.all_reduce(t)
        vec = t.numpy()
        gi = lambda slot, facsn: vec[3 * slot] + facsn * (vec[3 * slot + 1] + vec[3 * slot + 2])
        wy = np.ones(nyp); wy[0] = wy[-1] = 0.5
        e_p = abs(gi(0, 0.5) - wx @ pfield @ wy)
        e_t = abs(gi(1, 1.0) - (wx @ tfield).sum())
        lo = min(vec[6 + 4 * r] for r in range(world))
        hi = max(vec[7 + 4 * r] for r in range(world))
        jval, jpos = 0.0, 0
        for r in range(world):                           # ranks hold increasing rows: first strict maximum wins
            if vec[8 + 4 * r] > jval:
                jval, jpos = vec[8 + 4 * r], int(vec[9 + 4 * r])
        ok = (lo == tfield.min() and hi == tfield.max() and jval == 9.0 and jpos == 4)
        q.put((rank, max(e_p, e_t), ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nyp", [(2, 41), (3, 100)])
def test_monitor_share_algebra_over_gloo(world, nyp):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_monitor_worker, args=(r, world, port, nyp, 33, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, ok in res:
        assert err <= 1e-11 and ok, (rank, err, ok)

#!/usr/bin/env python
"""Benchmark of the Q-GCM ocean step on B200 (BASELINE.json metric: ocean timesteps/s and
grid-point-updates/s on the NAtl 1 km ocean-only deck, with the HBM-roofline fraction).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload natl1km] [--impl reference]

A "step" is one ocean timestep (oml + qgostep + ocinvq + ocqbdy, plus the time-level
average on its 1-in-25 cadence) on synthetic fields of the deck's grid shape.

* value : steps/s with the whole state resident in HBM (timed with CUDA events on the
          library's stream, max over ranks).
* e2e   : the same step driven through the C ABI with HOST buffers: every step uploads
          the externally supplied forcing (tauxo, tauyo, fnetoc) from pinned memory with
          qgcm_set_field, runs qgcm_ocean_step and reads the scalar state back with
          qgcm_get_scalars.
* roofline : per-kernel CUDA-event times from the library's own instrumentation
          (qgcm_profile) over a second pass of the same K steps; the dominant kernel's
          algorithmic bytes / its mean launch time against MEASURED_PEAKS.json.
* cpu_baseline / --impl reference : the CPU restatement of the reference algorithm
          (oracle/liborc.so, C++/OpenMP, all host cores) on a bounded sample of the same
          workload.  The real Fortran reference cannot be built in this image (no Fortran
          compiler), so kind = "port".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ocean_timesteps_per_s"
UNIT = "steps/s"


def passes_per_launch(kernel, cyclic):
    """algorithmic field passes (8*nxpo*nypo bytes each) one launch of `kernel` must move;
    SURVEY.md section 8(d) / DESIGN.md kernel table.  The synthetic decks have a flat bottom
    (ddynoc = 0, as topset 'flat' leaves it), which the library detects: the right-hand side
    kernel then moves 6 passes instead of SURVEY's 7 (its entry below charges the 6 it moves;
    the step-level figure keeps SURVEY's 61 and reports the 60-pass one beside it)."""
    table = {
        "k_oml_step": 9.0, "k_oml_entoc": 2.0,
        "k_qgstep": 17.0,
        # box decks with the fast DST plan run ocinvq fused: the forward transform reads q and writes the
        # spectral layers (6), the inverse reads them and ochom and writes p (8 box); k_l2m / k_m2l are
        # launched only on the unfused path (channel, atmosphere), where both transforms move 6
        "k_l2m": 6.0, "k_xform": 6.0, "k_xform_inv": 6.0 if cyclic else 8.0, "k_tri_local": 6.0, "k_tri_fg": 3.0,
        "k_m2l": 6.0 if cyclic else 8.0,
        "k_avg2": None,
    }
    return table.get(kernel)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self, first=0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[first:]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(p, world, parallelism, transport=None, transport_note=None, extra=None):
    """the `config` object of the JSON line: the same key set in both arms (the driver compares them)"""
    coupled = not p.has("ocean_only")
    fieldpass = 8.0 * p.nxpo * p.nypo
    cfgd = {
        "workload": "%s %s %dx%dx%d, %s, dto=%gs" % (p.name, "coupled" if coupled else "ocean-only", p.nxpo, p.nypo, p.nlo,
                                                      "channel" if p.has("cyclic_ocean") else "box", p.dto),
        "parallelism": parallelism,
        "transport": transport, "transport_note": transport_note,
        "l2": ("state (%.1f GB) is far larger than L2; no flush needed" if 4 * p.nlo * fieldpass > 4 * 126e6 else
               "state (%.2f GB) is comparable to the 126 MB L2 and is NOT flushed between steps: a parity-size "
               "deck, not a bench line") % (4 * p.nlo * fieldpass / 1e9),
        "state_finite": None,
    }
    if extra:
        cfgd.update(extra)
    return cfgd


def omp_all_cores():
    """the oracle's OpenMP team on every host core this process may use.  torchrun exports
    OMP_NUM_THREADS=1 to its workers and torch's libgomp has read it by the time the oracle is
    loaded, so the team size is also set through the runtime call."""
    cores = len(os.sched_getaffinity(0))
    if "TORCHELASTIC_RUN_ID" in os.environ or "OMP_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
    except OSError:
        pass
    return int(os.environ["OMP_NUM_THREADS"])


def verify_against_oracle(qg, m, p, cfg, world, rank, dist, torch, nsteps=2):
    """--verify leg, outside every timed region: the model `m` (one GPU, or this rank's y-slab over
    the transport the bench is about to time) and the CPU oracle on rank 0 take the same `nsteps`
    ocean steps from the same synthetic state; every rank compares the rows it owns.  Returns
    {field: relative L2 over the whole domain} on every rank.  The oracle's fields travel through
    /dev/shm (ranks of one node), not through the GPU."""
    import numpy as np
    names = ("po", "qo", "sst", "entoc")
    tag = "/dev/shm/qgcm_verify_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", str(os.getppid())))
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyorc
        pyorc.build()
        omp_all_cores()
        o = pyorc.Oracle(cfg)
        qg.synth.init_model(o, p, cfg, "random")
        for _ in range(nsteps):
            o.ocean_step()
        os.makedirs(tag, exist_ok=True)
        for n in names:
            np.save(os.path.join(tag, n + ".npy"), o.get_field(n))
        o.close()
    if world > 1:
        dist.barrier()      # the ranks enter the first exchange together (rank 0 was busy on the host)
    for _ in range(nsteps):
        m.ocean_step()
    m.sync()
    if world > 1:
        dist.barrier()
    j0, nown = qg.slab_bounds(p.nypo, world, rank) if world > 1 else (0, p.nypo)
    acc = np.zeros(2 * len(names))
    for i, n in enumerate(names):
        ref = np.load(os.path.join(tag, n + ".npy"), mmap_mode="r")
        got = m.get_field(n)
        nx = p.nxpo if n in ("po", "qo", "entoc") else p.nxto
        nyg = p.nypo if n in ("po", "qo", "entoc") else p.nyto
        nl = ref.size // (nx * nyg)
        a = got.reshape((nx, nyg, nl), order="F")[:, j0:min(j0 + nown, nyg), :]
        b = np.asarray(ref).reshape((nx, nyg, nl), order="F")[:, j0:min(j0 + nown, nyg), :]
        acc[2 * i] = float(((a - b) ** 2).sum())
        acc[2 * i + 1] = float((b ** 2).sum())
    if world > 1:
        t = torch.tensor(acc, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        acc = t.cpu().numpy()
        dist.barrier()
    if rank == 0:
        import shutil
        shutil.rmtree(tag, ignore_errors=True)
    return {n: float(np.sqrt(acc[2 * i] / acc[2 * i + 1])) if acc[2 * i + 1] > 0 else float(np.sqrt(acc[2 * i]))
            for i, n in enumerate(names)}


def build_case(qg, workload, device):
    p = qg.named_config(workload)
    cfg = qg.build_config(p, device=device)
    return p, cfg


def slab_model(qg, cfg, world, rank, dist, torch, transport):
    """one rank of the y-slab partition with its communicator(s): NCCL always (rank 0 makes the
    id, torch.distributed carries it), the peer-memory mailboxes when asked for"""
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(qg.Model.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(ident, 0)
    m = qg.Model(qg.slab_config(cfg, world, rank))
    m.comm_init_nccl(bytes(ident.cpu().numpy().tobytes()))
    if transport == "peer":
        mine = torch.frombuffer(bytearray(m.peer_handle()), dtype=torch.uint8).cuda()
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        m.comm_init_peer([bytes(h.cpu().numpy().tobytes()) for h in allh])
    dist.barrier()      # the mailbox waits give up after ~10 s: enter the first exchange together
    return m


def peer_self_check(qg, local, world, rank, dist, torch):
    """the same three steps of a small box deck over NCCL and over the peer mailboxes must agree
    on every rank (the sums differ only in their order); returns None or the reason to fall back"""
    import numpy as np
    p = qg.named_config("natl1km").scaled(24, 12, ndxr=40, name="peer_check")     # 960 x 480
    if p.nypo < 16 * world:
        return "check deck too small"
    cfg = qg.build_config(p, device=local)
    reason = None
    out = {}
    try:
        for kind in ("nccl", "peer"):
            m = slab_model(qg, cfg, world, rank, dist, torch, kind)
            qg.synth.init_model(m, p, cfg, "random")
            dist.barrier()
            for _ in range(3):
                m.ocean_step()
            m.sync()
            j0, n = qg.slab_bounds(p.nypo, world, rank)
            out[kind] = {k: m.get_field(k).reshape((p.nxpo, p.nypo, p.nlo), order="F")[:, j0:j0 + n, :].copy() for k in ("po", "qo")}
            dist.barrier()      # no rank unmaps a mailbox another rank may still be writing to
            if kind == "peer":
                m.comm_close_peer()
                dist.barrier()  # no rank frees a mailbox another rank still maps
            m.close()
        for k in ("po", "qo"):
            a, b = out["peer"][k], out["nccl"][k]
            e = float(np.linalg.norm(a - b) / np.linalg.norm(b))
            if not e <= 1e-12:
                reason = "peer and NCCL transports differ on %s: %.2e" % (k, e)
    except Exception as ex:       # noqa: BLE001 - any failure means "use NCCL"
        reason = "peer transport unavailable: %s" % str(ex)[:200]
    flag = torch.tensor([0 if reason is None else 1], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()) and reason is None:
        reason = "peer transport failed its self-check on another rank"
    return reason


def cpu_sample(qg, p, cfg, budget_s=15.0, max_steps=64):
    """time the CPU port on the same workload: as many ocean steps as fit the budget"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyorc
    pyorc.build()
    cores = omp_all_cores()
    o = pyorc.Oracle(cfg)
    qg.synth.init_model(o, p, cfg, "random")
    coupled = not p.has("ocean_only")
    state = {"nt": 1}

    def step():
        if coupled:      # xforc + ocean step + nstr atmosphere steps, as the CUDA arm's step
            o.run(state["nt"], state["nt"] + p.nstr - 1)
            state["nt"] += p.nstr
        else:
            o.ocean_step()

    step()   # warm-up (page faults of the automatic arrays)
    t0 = time.time()
    n = 0
    while n < max_steps:
        step()
        n += 1
        if time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    o.close()
    return n / dt, cores, "%d ocean steps of %s (%dx%dx%d) after 1 warm-up" % (n, p.name, p.nxpo, p.nypo, p.nlo)


def run_reference(args, qg):
    """--impl reference: the CPU restatement of the reference's own path (oracle/liborc.so, C++/OpenMP,
    every host core) on the same workload; W warm-up steps and K timed steps as asked (a 1 km CPU step
    costs ~0.3 s), bounded at a few minutes"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p, cfg = build_case(qg, args.workload, 0)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyorc
    pyorc.build()
    cores = omp_all_cores()
    o = pyorc.Oracle(cfg)
    qg.synth.init_model(o, p, cfg, "random")
    coupled = not p.has("ocean_only")
    state = {"nt": 1}

    def step():
        if coupled:
            o.run(state["nt"], state["nt"] + p.nstr - 1)
            state["nt"] += p.nstr
        else:
            o.ocean_step()

    warm = max(args.warmup, 1)
    t_w = time.time()
    done_w = 0
    for _ in range(warm):
        step()
        done_w += 1
        if time.time() - t_w > 60.0:
            break
    t0 = time.time()
    n = 0
    while n < args.steps:
        step()
        n += 1
        if time.time() - t0 > 150.0:
            break
    dt = time.time() - t0
    v = n / dt
    step_passes = 59.0 if p.has("cyclic_ocean") else 61.0
    # the reference's own Fortran, translated to C++ (oracle/_ref, oracle/f2cpp.py) on the same state: single
    # threaded -- the translator drops the OpenMP directives -- so it is reported beside the OpenMP restatement,
    # which stays the arm's value (the faster, i.e. the more demanding, baseline)
    translated = None
    if not args.no_translated and not coupled:
        try:
            import pyref
            if pyref.available():
                r = pyref.RefModel(p, cfg)
                qg.synth.init_model(r, p, cfg, "random")

                def rstep():
                    r.oml(); r.qgostep(); r.ocinvq(); r.ocqbdy()

                rstep()
                t1 = time.time()
                k = 0
                while k < 3:
                    rstep()
                    k += 1
                    if time.time() - t1 > 40.0:
                        break
                translated = {"value": k / (time.time() - t1), "unit": UNIT, "cores": 1, "kind": "reference",
                              "sample": "%d ocean steps after 1 warm-up of the reference's Fortran sources translated "
                                        "statement by statement to C++ (oracle/_ref; OpenMP directives not honoured)" % k}
                del r
        except Exception as e:      # the arm's own number does not depend on this leg
            translated = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": done_w, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(p, args.gpus, "%d host threads (OpenMP), no GPU" % cores),
        "gpt_updates_per_s": v * p.nxpo * p.nypo * p.nlo / 1e9,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "effective_GBps": v * step_passes * 8.0 * p.nxpo * p.nypo / 1e9,
                         "sample": "%d CPU ocean steps after %d warm-up (C++/OpenMP restatement of the reference; the "
                                   "Fortran reference cannot be compiled here: no Fortran compiler in the image or on "
                                   "the GPU box, profiles/r02_fortran_probe.txt)" % (n, done_w)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_translated": translated,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="natl1km")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-translated", action="store_true",
                    help="--impl reference: skip the single-threaded run of the translated Fortran reference")
    ap.add_argument("--no-verify", action="store_true",
                    help="skip the parity leg (two ocean steps against the CPU oracle before the warm-up, outside every timed region)")
    ap.add_argument("--transport", default=os.environ.get("QGCM_SLAB_TRANSPORT", "peer"), choices=["peer", "nccl"],
                    help="y-slab exchanges at N > 1: peer-memory mailboxes (default) or NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    # the bench prints ONE JSON line on stdout: keep NCCL's version banner out of it
    # (NCCL writes its debug output, version banner included, to stdout unless told otherwise)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ.pop("NCCL_DEBUG", None)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import _pkg
    qg = _pkg.load()
    if args.impl == "reference":
        run_reference(args, qg)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p, cfg = build_case(qg, args.workload, local)
    transport, transport_note = None, None
    if world > 1:
        # strong scaling: ONE domain cut into y-slabs, one per GPU; the ranks exchange halo
        # rows, the 2-row-per-mode slab coupling of the Helmholtz solve and a few scalars
        # (q-gcm_b200/csrc/slab.cu) -- through peer-memory mailboxes written by the kernels of
        # the step over NVLink, or through NCCL (--transport nccl, and the fallback when the
        # peer transport fails its self-check against NCCL on a small deck)
        transport = args.transport
        if transport == "peer":
            transport_note = peer_self_check(qg, local, world, rank, dist, torch)
            if transport_note is not None:
                transport = "nccl"
        m = slab_model(qg, cfg, world, rank, dist, torch, transport)
    else:
        m = qg.Model(cfg)
    qg.synth.init_model(m, p, cfg, "random")
    m.sync()
    if world > 1:
        dist.barrier()
    parity = None
    if not args.no_verify and p.has("ocean_only"):
        # parity at the benched size over the benched transport, before anything is timed
        parity = verify_against_oracle(qg, m, p, cfg, world, rank, dist, torch)
    stream = torch.cuda.ExternalStream(m.stream(), device=local)
    nstr = p.nstr
    cad = 25   # time-level average every 25 ocean steps (src/q-gcm.F:1328)

    coupled = not p.has("ocean_only")

    def ocean_steps(n, first):
        if coupled:
            # one ocean step = xforc + ocean step + nstr atmosphere steps (+ the averaging on its
            # cadence): the loop body of src/q-gcm.F:1220-1408 for nstr values of nt
            m.run((first - 1) * nstr + 1, (first - 1 + n) * nstr)
            return
        for s in range(first, first + n):
            m.ocean_step()
            if s % cad == 0:
                m.tlavg_ocean()

    def barrier():
        if world > 1:
            dist.barrier()
        m.sync()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.5 s before its first sample
    ocean_steps(args.warmup, 1)
    barrier()
    def step_while(cond):
        """keep stepping (untimed) while rank 0's condition holds; the step is collective on
        y-slabs, so every rank follows rank 0's decision"""
        while True:
            go = torch.tensor([1 if (rank == 0 and cond()) else 0], dtype=torch.int32, device="cuda")
            if world > 1:
                dist.broadcast(go, 0)
            if int(go.item()) == 0:
                break
            ocean_steps(5, 1)
            m.sync()

    # keep the GPU under the same load until the sampler is live, so that the samples taken
    # while the timed region runs are not its start-up transient
    t_w = time.time()
    step_while(lambda: len(sampler.rows) < 2 and time.time() - t_w < 5.0)
    barrier()
    n_before = len(sampler.rows)
    l0 = m.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ocean_steps(args.steps, args.warmup + 1)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = m.launch_count() - l0
    clocks = None
    # a timed region shorter than the sampling period: keep stepping (untimed) until at least
    # three samples under the identical load exist
    t_w = time.time()
    step_while(lambda: len(sampler.rows) - n_before < 3 and time.time() - t_w < 3.0)
    if rank == 0:
        clocks = sampler.stop(first=n_before)
        clocks["note"] = "sampled every 100 ms from the start of the timed region, same step loop"
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # N > 1: one domain in y-slabs, so the job's throughput is steps of that one domain
    value = args.steps / (ms * 1e-3)

    # ---- per-kernel profile pass (CUDA events inside the library) ----
    m._lib.qgcm_profile(m._h, 1)
    ocean_steps(args.steps, args.warmup + args.steps + 1)
    buf = C.create_string_buffer(1 << 16)
    m._call("profile_report", buf, C.c_int64(len(buf)))
    m._lib.qgcm_profile(m._h, 0)
    prof = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, tot = ln.split()
        prof[name] = (int(cnt), float(tot))
    # the right-hand-side kernel runs as one interior launch plus two small edge launches that overlap it
    # on a side stream; the profile pass serialises them, so charging their time to k_qgstep is the
    # conservative reading (its 17 passes are the three launches' work together)
    merged = {}
    if "k_qgstep_edge" in prof and "k_qgstep" in prof:
        ec, et = prof.pop("k_qgstep_edge")
        qc, qt = prof["k_qgstep"]
        prof["k_qgstep"] = (qc, qt + et)
        merged["k_qgstep"] = {"includes": "k_qgstep_edge", "edge_launches": ec, "edge_ms_total": et}
    tot_ms = sum(v[1] for v in prof.values())
    fieldpass = 8.0 * p.nxpo * p.nypo
    peak, peak_src = hbm_peak()
    kern = {}
    for name, (cnt, tms) in prof.items():
        pp = passes_per_launch(name, p.has("cyclic_ocean"))
        ent = {"launches": cnt, "ms_per_launch": tms / cnt, "share": tms / tot_ms}
        if pp:
            ent["GBps"] = pp * fieldpass / world / (tms / cnt * 1e-3) / 1e9     # a rank moves 1/world of the rows
            ent["frac"] = ent["GBps"] / peak
        ent.update(merged.get(name, {}))
        kern[name] = ent
    # the dominant kernel is chosen per CUDA FUNCTION, as the ncu launch list names it: the forward and the
    # inverse sine transform are two instantiations of k_dst3 (profile names k_xform / k_xform_inv), so their
    # launches are averaged -- algorithmic bytes of all its launches over the time of all its launches
    func_of = {"k_xform": "k_dst3", "k_xform_inv": "k_dst3"}
    funcs = {}
    for name, (cnt, tms) in prof.items():
        f = funcs.setdefault(func_of.get(name, name), {"launches": 0, "ms": 0.0, "passes": 0.0, "names": []})
        f["launches"] += cnt
        f["ms"] += tms
        f["passes"] += cnt * (passes_per_launch(name, p.has("cyclic_ocean")) or 0.0)
        f["names"].append(name)
    dom = max(funcs, key=lambda k: funcs[k]["ms"])
    fd = funcs[dom]
    dpp = fd["passes"] / fd["launches"]
    dms = fd["ms"] / fd["launches"]
    ach = dpp * fieldpass / world / (dms * 1e-3) / 1e9
    traffic = traffic_src = None
    try:      # DRAM bytes per launch from the committed ncu --set full capture (single GPU, natl1km)
        src = os.path.join("profiles", "r02_dram_traffic.json")
        with open(os.path.join(ROOT, src)) as f:
            if world == 1 and args.workload == "natl1km":
                traffic = json.load(f)["bytes_per_launch"].get(dom)
                traffic_src = src + " (ncu --set full capture of this command, not measured in this run)"
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "profile_names": sorted(fd["names"]), "achieved": ach, "peak": peak,
            "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "traffic_source": traffic_src, "launches_per_step": fd["launches"] / max(1, args.steps),
            "algorithmic_bytes_per_launch": dpp * fieldpass / world, "ms_per_launch": dms,
            "share_of_step": fd["ms"] / tot_ms}
    # ocean-only byte model: SURVEY.md 8(d)'s algorithmic figure is 61 field passes per step (59 in a
    # channel), topography field included; over the flat bottom of the synthetic decks the library
    # skips that field, so the step actually has to move one pass less.  A coupled step moves the
    # atmosphere and xforc too, which neither figure counts, so its fraction is a lower bound.
    ocean_passes = 61.0 if not p.has("cyclic_ocean") else 59.0
    step_bytes = ocean_passes * fieldpass
    byte_model = "SURVEY.md 8(d): %d field passes of 8*nxpo*nypo bytes per ocean step" % ocean_passes
    # what this implementation has to move: the fused inversion of the box decks needs 23 passes where
    # SURVEY's unfused minimum counts 33 (no separate layer<->mode projections), and the flat bottom of
    # the synthetic decks saves the topography pass
    fused = "k_l2m" not in prof and not p.has("cyclic_ocean")
    moved_bytes = step_bytes - (11.0 if fused else 1.0) * fieldpass
    if coupled:
        # SURVEY.md 8(d), coupled step accounting: per ocean step one xforc on the ocean-resolution
        # atmosphere grid (about 10 passes of 8 bytes over (nxta*ndxr+1)*(nyta*ndxr+1) points) and nstr
        # atmosphere steps (the same 59-pass channel model on the nxpa*nypa grid)
        fine = 8.0 * (p.nxta * p.ndxr + 1) * (p.nyta * p.ndxr + 1)
        atm = 8.0 * (p.nxta + 1) * (p.nyta + 1)
        step_bytes += 10.0 * fine + p.nstr * 59.0 * atm
        moved_bytes += 10.0 * fine + p.nstr * 59.0 * atm
        byte_model += " + xforc 10 passes over the %dx%d fine grid + %d atmosphere steps of 59 passes" % (
            p.nxta * p.ndxr + 1, p.nyta * p.ndxr + 1, p.nstr)
    step_frac = step_bytes * value / world / 1e9 / peak     # per-GPU share of the step's bytes against one GPU's peak

    # ---- end to end through the C ABI with host buffers ----
    e2e = e2e_cadence = restart = None
    if not args.no_e2e and not coupled:
        names = ("tauxo", "tauyo", "fnetoc")
        # the externally supplied forcing as the Fortran side holds it: global host arrays
        st = qg.synth.ocean_state(p, cfg, "random", qg.synth.SEED, min(1.0, (p.nxto * p.dxo) / 4.8e6 * 4.0))
        host = {n: torch.from_numpy(np.ascontiguousarray(np.asarray(st[n], dtype=np.float64).ravel(order="F"))).pin_memory()
                for n in names}
        ptr = {n: C.cast(host[n].data_ptr(), C.POINTER(C.c_double)) for n in names}
        nel = {n: host[n].numel() for n in names}
        scal = qg.QgcmScalars()
        ke = max(3, min(args.steps, 20))

        def upload():
            for n in names:
                m._call("set_field_async", n.encode(), ptr[n], C.c_int64(nel[n]))

        def e2e_step():
            # this step runs on the forcing committed last time while the next step's forcing
            # crosses PCIe on the copy stream; the scalar read-back synchronises the step
            m.ocean_step()
            upload()
            m._call("get_scalars", C.byref(scal))
            m._call("commit_fields")

        upload()
        m._call("commit_fields")
        e2e_step()
        barrier()
        e0.record(stream)
        t_host = time.perf_counter()
        for _ in range(ke):
            e2e_step()
        e1.record(stream)
        barrier()
        t_host = time.perf_counter() - t_host      # host clock around the calls a user makes (incl. the last sync)
        t = torch.tensor([e0.elapsed_time(e1), t_host * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ev, t_wall = float(t[0].item()), float(t[1].item())
        e2e = {"value": ke / (t_wall * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(sum(nel.values()) * 8), "d2h_bytes_per_step": int(C.sizeof(scal)),
               "steps": ke, "clock": "host perf_counter around the API calls, max over ranks",
               "value_cuda_events": ke / (t_ev * 1e-3),
               "def": "per step: qgcm_set_field_async(tauxo,tauyo,fnetoc) from pinned host memory (overlapped "
                      "with the step on a copy stream) + qgcm_ocean_step + qgcm_get_scalars + "
                      "qgcm_commit_fields; PCIe-bound: the forcing is 3 full fields per step, which the "
                      "ocean-only reference uploads never (time-invariant inputs, SURVEY quirk 6): see e2e_cadence "
                      "for the main program's real traffic"}

        # ---- e2e_cadence: one model day driven as the Fortran main program would drive the library
        # (src/q-gcm.F:1271-1489): every ocean step qgcm_ocean_step (+ the time-level average on its 1-in-25
        # cadence); valids every valday = 0.25 d on the device (a 200-byte report instead of 4 fields);
        # monnc_ocean + tavocn every dgnday / dtavoc = 1 d; one sub-sampled ocnc_out read (po, qo, sst, wekto,
        # tauxo, tauyo every nsko-th point, src/nc_subs.F:845-890); time on the host clock.
        nday = max(4, int(round(86400.0 / p.dto)))
        nsko = 8
        m.tavini()
        m.sync()
        barrier()
        d2h = 0
        t_host = time.perf_counter()
        for sidx in range(1, nday + 1):
            m.ocean_step()
            if sidx % cad == 0:
                m.tlavg_ocean()
            if sidx % max(1, nday // 4) == 0:
                rep = m.valids()
                d2h += C.sizeof(rep)
                if world == 1 and not rep.solnok:
                    raise RuntimeError("valids: the benchmark state left its valid range")
        mon = m.monnc_ocean()
        m.tavocn()
        d2h += C.sizeof(mon)
        for nm in ("po", "qo", "sst", "wekto", "tauxo", "tauyo"):
            d2h += m.get_field_sub(nm, nsko).nbytes
        m.sync()
        t_host = time.perf_counter() - t_host
        t = torch.tensor([t_host], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_cadence = {"value": nday / float(t.item()), "unit": UNIT, "steps": nday, "model_days": nday * p.dto / 86400.0,
                       "d2h_bytes_total": int(d2h), "h2d_bytes_total": 0, "clock": "host perf_counter, max over ranks",
                       "def": "one model day as src/q-gcm.F:1271-1489 drives it: %d x qgcm_ocean_step, time-level average "
                              "every 25 steps, 4 x qgcm_valids, qgcm_monnc_ocean + qgcm_tavocn, one ocnc_out read of 6 fields "
                              "sub-sampled by %d; state resident, forcing time-invariant" % (nday, nsko)}

        # ---- restart download (resave_nc, src/nc_subs.F:1331-1360: po, pom, sst, sstm of the ocean) into
        # page-locked caller arrays with one synchronisation (qgcm_host_register + qgcm_get_fields)
        if world == 1:
            arrs = {n: np.empty(m.field_size(n)) for n in ("po", "pom", "sst", "sstm")}
            for a in arrs.values():
                a.fill(0.0)
                qg.Model.host_register(a)
            m.get_fields(arrs)          # warm
            t_r = time.perf_counter()
            m.get_fields(arrs)
            t_r = time.perf_counter() - t_r
            pageable = {n: np.empty(m.field_size(n)) for n in arrs}
            for a in pageable.values():
                a.fill(0.0)
            t_p = time.perf_counter()
            for n, a in pageable.items():
                m._call("get_field", n.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(a.size))
            t_p = time.perf_counter() - t_p
            nbytes = sum(a.nbytes for a in arrs.values())
            restart = {"bytes": int(nbytes), "GBps_registered_batched": nbytes / t_r / 1e9, "GBps_pageable_per_field": nbytes / t_p / 1e9,
                       "def": "po, pom, sst, sstm -> host: qgcm_get_fields into arrays page-locked once with "
                              "qgcm_host_register, against qgcm_get_field into pageable arrays"}
            for a in arrs.values():
                qg.Model.host_unregister(a)

    po = m.get_field("po")
    if world > 1:       # a slab fills only the rows it owns
        j0, n = qg.slab_bounds(p.nypo, world, rank)
        po = po.reshape((p.nxpo, p.nypo, p.nlo), order="F")[:, j0:j0 + n, :]
    finite = bool(np.isfinite(po).all())
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_sample(qg, p, cfg)
        # effective_GBps: the same byte model as step_roofline_frac
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "effective_GBps": v * step_bytes / 1e9}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong",      # one fixed domain at every N (N > 1: y-slabs of it)
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(
                p, world,
                ("%d y-slabs of one domain, %s, slab-coupled solve (2 rows per mode exchanged)" %
                 (world, "exchanges by the step's own kernels through NVLink peer-memory mailboxes"
                  if transport == "peer" else "NCCL halos and reductions")) if world > 1 else "single GPU",
                transport, transport_note, {"state_finite": finite}),
            # --verify: relative L2 against the CPU oracle after two ocean steps of this very workload
            # on this very transport (every rank compares its own rows); the bar is 1e-11
            "parity_rel_l2": parity,
            "parity_ok": (max(parity.values()) <= 1e-11) if parity else None,
            "gpt_updates_per_s": value * p.nxpo * p.nypo * p.nlo / 1e9,
            # the step against the HBM roofline: SURVEY's algorithmic byte model (the figure the >= 60 % target
            # and round 1 are quoted on), and beside it the bytes this implementation actually has to move
            "step_roofline_frac": step_frac,
            "step_algorithmic_bytes": step_bytes,
            "step_byte_model": byte_model,
            "step_roofline_frac_moved_bytes": step_frac * moved_bytes / step_bytes,
            "step_moved_bytes": moved_bytes,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": kern,
            "e2e": e2e,
            "e2e_cadence": e2e_cadence,
            "restart_download": restart,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        m.sync()
        dist.barrier()          # no rank unmaps a mailbox another rank may still be writing to
        if transport == "peer":
            m.comm_close_peer()
            dist.barrier()      # no rank frees a mailbox another rank still maps
        m.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

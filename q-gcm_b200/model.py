"""ctypes binding of the C ABI in include/qgcm_b200.h.

``Model`` is the product binding (libqgcm_b200.so, CUDA only).  Its methods carry the
names of the reference subroutines they replace (src/q-gcm.F:1222-1269) so the parity
tests read like the reference's main loop.  ``CModel`` is the generic binding (library + symbol
prefix); the tests reuse it to drive their CPU checker through the same struct layout.
"""
import ctypes as C
import os

import numpy as np

from .abi import QgcmConfig, QgcmScalars, QgcmValidsReport, QgcmMonitorOcean, QgcmMonitorAtmos, declared_functions

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path():
    # QGCM_B200_LIB: an alternative build of the same library (kernel A/B experiments)
    return os.environ.get("QGCM_B200_LIB") or os.path.join(_HERE, "csrc", "libqgcm_b200.so")


def load_library():
    """load libqgcm_b200.so; raises (never falls back) if it has not been built"""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "libqgcm_b200.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback" % path)
        _LIB = C.CDLL(path, mode=C.RTLD_GLOBAL)
    return _LIB


class CModel:
    """one model instance behind a C ABI with the given symbol prefix"""

    def __init__(self, lib, prefix, cfg: QgcmConfig):
        self._lib = lib
        self._pfx = prefix
        self.cfg = cfg
        self._h = C.c_void_p()
        create = self._fn("create")
        create.argtypes = [C.POINTER(QgcmConfig), C.POINTER(C.c_void_p)]
        rc = create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise RuntimeError("%screate failed: %s" % (prefix, self._err()))
        self.nxpo, self.nypo = cfg.nxto + 1, cfg.nyto + 1
        self.nxpa, self.nypa = cfg.nxta + 1, cfg.nyta + 1

    # -- plumbing
    def _fn(self, name):
        return getattr(self._lib, self._pfx + name)

    def _err(self):
        f = self._fn("last_error")
        f.restype = C.c_char_p
        return (f() or b"").decode()

    def _call(self, name, *args):
        f = self._fn(name)
        f.restype = C.c_int
        rc = f(self._h, *args)
        if rc != 0:
            raise RuntimeError("%s%s failed: %s" % (self._pfx, name, self._err()))

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state transfer
    def field_size(self, name):
        n = C.c_int64()
        self._call("field_size", name.encode(), C.byref(n))
        return n.value

    def set_field(self, name, arr):
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).ravel(order="F"))
        self._call("set_field", name.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(a.size))

    def get_field(self, name, shape=None):
        n = self.field_size(name)
        out = np.empty(n, dtype=np.float64)
        self._call("get_field", name.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(n))
        if shape is not None:
            out = out.reshape(shape, order="F")
        return out

    def get_fields(self, arrays: dict):
        """batched download into caller-owned flat float64 arrays (one synchronisation): {name: array}"""
        self._fields_call("get_fields", arrays)

    def set_fields(self, arrays: dict):
        self._fields_call("set_fields", arrays)

    def _fields_call(self, which, arrays):
        n = len(arrays)
        names = (C.c_char_p * n)(*[k.encode() for k in arrays])
        ptrs = (C.POINTER(C.c_double) * n)(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrays.values()])
        cnts = (C.c_int64 * n)(*[a.size for a in arrays.values()])
        self._call(which, C.c_int32(n), names, ptrs, cnts)

    def set_scalars(self, s: QgcmScalars):
        self._call("set_scalars", C.byref(s))

    def get_scalars(self) -> QgcmScalars:
        s = QgcmScalars()
        self._call("get_scalars", C.byref(s))
        return s

    def helmholtz(self, which, wrk, b):
        w = np.ascontiguousarray(np.asarray(wrk, dtype=np.float64).ravel(order="F"))
        bb = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
        self._call("helmholtz", C.c_int(which), w.ctypes.data_as(C.POINTER(C.c_double)),
                   bb.ctypes.data_as(C.POINTER(C.c_double)))
        return w.reshape(np.asarray(wrk).shape, order="F")

    # -- the reference's procedures
    def constr(self): self._call("constr")
    def homsol(self): self._call("homsol")
    def qcomp_ocean(self): self._call("qcomp_ocean")
    def qcomp_atmos(self): self._call("qcomp_atmos")
    def xforc(self): self._call("xforc")
    def oml(self): self._call("oml")
    def qgostep(self): self._call("qgostep")
    def ocinvq(self): self._call("ocinvq")
    def ocqbdy(self): self._call("ocqbdy")
    def aml(self): self._call("aml")
    def qgastep(self): self._call("qgastep")
    def atinvq(self): self._call("atinvq")
    def atqzbd(self): self._call("atqzbd")
    def tlavg_ocean(self): self._call("tlavg_ocean")
    def tlavg_atmos(self): self._call("tlavg_atmos")
    def ocean_step(self): self._call("ocean_step")
    def atmos_step(self): self._call("atmos_step")

    def run(self, nt_first, nt_last):
        self._call("run", C.c_int64(nt_first), C.c_int64(nt_last))

    def valids(self) -> QgcmValidsReport:
        """valids, src/valsubs.F:43: extreme-value and layer-thickness scan"""
        r = QgcmValidsReport()
        self._call("valids", C.byref(r))
        return r

    # -- running sums of src/timavge.F (read them with get_field under the reference's names)
    def tavini(self): self._call("tavini")
    def tavatm(self): self._call("tavatm")
    def tavocn(self): self._call("tavocn")
    def avg_ocn_k247(self): self._call("avg_ocn_k247")

    def tav_counts(self):
        """(nsumat, nsumoc, nsum_ocavg), src/timavge.F:46, :76 and src/timinfo_data.F"""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._call("tav_counts", C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def monnc_ocean(self) -> QgcmMonitorOcean:
        """ocean section of monnc_comp, src/monitor_diag.F:480-840"""
        r = QgcmMonitorOcean()
        self._call("monnc_ocean", C.byref(r))
        return r

    def monnc_atmos(self) -> QgcmMonitorAtmos:
        """atmosphere section of monnc_comp, src/monitor_diag.F:186-478"""
        r = QgcmMonitorAtmos()
        self._call("monnc_atmos", C.byref(r))
        return r

    def qocdiag(self, nsko, out=None):
        """qocdiag_out, src/qocdiag.F:303: (ipwk, jpwk, nlo, 5) = dqdt, qotjac, qt2dif, qt4dif, qotent"""
        nx, ny = self.nxpo, self.nypo
        iw = min(nx % nsko, 1) + (nx - nx % nsko) // nsko
        jw = min(ny % nsko, 1) + (ny - ny % nsko) // nsko
        n = 5 * iw * jw * self.cfg.nlo
        if out is None:
            out = np.full(n, np.nan, dtype=np.float64)
        self._call("qocdiag", C.c_int32(nsko), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(n))
        return out

    # -- convenience: whole-state load/store used by tests and bench
    OCEAN_FIELDS = ("po", "pom", "qo", "qom", "sst", "sstm", "wekto", "wekpo", "entoc", "tauxo", "tauyo",
                    "fnetoc", "ddynoc")
    ATMOS_FIELDS = ("pa", "pam", "qa", "qam", "ast", "astm", "hmixa", "hmixam", "wekta", "wekpa", "entat",
                    "tauxa", "tauya", "fnetat", "ddynat", "dtopat", "xc1ast", "uekat", "vekat")

    def load_state(self, state: dict):
        for k, v in state.items():
            self.set_field(k, v)


class Model(CModel):
    """the CUDA model (libqgcm_b200.so)"""

    def __init__(self, cfg: QgcmConfig):
        super().__init__(load_library(), "qgcm_", cfg)

    def sync(self):
        self._call("sync")

    def launch_count(self):
        f = self._lib.qgcm_launch_count
        f.restype = C.c_int64
        return int(f(self._h))

    def get_field_sub(self, name, nsk, out=None):
        """the sub-sampled vector of ocnc_out / atnc_out (src/nc_subs.F:869-890), packed on the device"""
        n = C.c_int64()
        self._call("field_sub_size", name.encode(), C.c_int32(nsk), C.byref(n))
        if out is None:
            out = np.full(n.value, np.nan, dtype=np.float64)
        self._call("get_field_sub", name.encode(), C.c_int32(nsk), out.ctypes.data_as(C.POINTER(C.c_double)), n)
        return out

    @staticmethod
    def host_register(arr):
        """page-lock a caller-owned numpy array for DMA (qgcm_host_register)"""
        lib = load_library()
        if lib.qgcm_host_register(C.c_void_p(arr.ctypes.data), C.c_int64(arr.nbytes)) != 0:
            f = lib.qgcm_last_error
            f.restype = C.c_char_p
            raise RuntimeError("qgcm_host_register failed: %s" % (f() or b"").decode())

    @staticmethod
    def host_unregister(arr):
        load_library().qgcm_host_unregister(C.c_void_p(arr.ctypes.data))

    def stream(self):
        f = self._lib.qgcm_stream
        f.restype = C.c_void_p
        return f(self._h)

    @staticmethod
    def exported_symbols_ok():
        lib = load_library()
        return [n for n in declared_functions() if not hasattr(lib, n)]

    # -- y-slab multi-GPU, one process per GPU
    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        lib = load_library()
        if lib.qgcm_nccl_unique_id(buf) != 0:
            f = lib.qgcm_last_error
            f.restype = C.c_char_p
            raise RuntimeError("qgcm_nccl_unique_id failed: %s" % (f() or b"").decode())
        return buf.raw

    def comm_init_nccl(self, id128: bytes):
        self._call("comm_init_nccl", C.c_char_p(id128))

    # -- peer-memory transport: CUDA IPC mailboxes, exchanged by the host program
    def peer_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._call("peer_handle", buf)
        return buf.raw

    def comm_init_peer(self, handles):
        """handles: every rank's peer_handle(), in rank order"""
        blob = b"".join(handles)
        self._call("comm_init_peer", C.c_char_p(blob), C.c_int32(len(handles)))

    def comm_close_peer(self):
        """unmap the other ranks' mailboxes; barrier between this and close() (CUDA IPC rule)"""
        self._call("comm_close_peer")

    def comm_peer_timeout(self, seconds: float):
        """give-up time of a mailbox wait (default 120 s)"""
        self._call("comm_peer_timeout", C.c_double(seconds))

    def comm_transport(self, kind: str):
        self._call("comm_transport", C.c_int32({"nccl": 0, "peer": 1}[kind]))


def slab_config(cfg: QgcmConfig, nranks: int, rank: int) -> QgcmConfig:
    """copy of cfg that selects y-slab `rank` of `nranks` (include/qgcm_b200.h, nranks/rank)"""
    c = type(cfg).from_buffer_copy(cfg)
    c.nranks, c.rank = nranks, rank
    for k, v in cfg.__dict__.items():      # host-side extras attached by build_config
        setattr(c, k, v)
    return c


def slab_bounds(nyp_global, nranks, rank):
    """owned p rows [j0, j0 + n) of a rank (pure host arithmetic in the library)"""
    j0, n = C.c_int32(), C.c_int32()
    lib = load_library()
    if lib.qgcm_slab_bounds(C.c_int32(nyp_global), C.c_int32(nranks), C.c_int32(rank), C.byref(j0), C.byref(n)) != 0:
        raise RuntimeError("qgcm_slab_bounds failed")
    return j0.value, n.value


class SlabGroup:
    """every rank of a y-slab partition in this process, on one device (qgcm_group_create).
    Mirrors the Model interface for the procedures that act on a partition, so the parity
    tests drive 2..8 slabs on a single GPU exactly as they drive one model."""

    def __init__(self, cfg: QgcmConfig, nranks: int):
        self.cfg = cfg
        self.ranks = [Model(slab_config(cfg, nranks, r)) for r in range(nranks)]
        arr = (C.c_void_p * nranks)(*[m._h for m in self.ranks])
        lib = load_library()
        if lib.qgcm_group_create(arr, C.c_int32(nranks)) != 0:
            raise RuntimeError("qgcm_group_create failed: %s" % self.ranks[0]._err())

    def set_field(self, name, arr):
        for m in self.ranks:
            m.set_field(name, arr)

    def get_field(self, name, shape=None):
        n = self.ranks[0].field_size(name)
        out = np.full(n, np.nan, dtype=np.float64)
        for m in self.ranks:      # each rank fills the rows it owns
            m._call("get_field", name.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(n))
        return out.reshape(shape, order="F") if shape is not None else out

    def get_scalars(self):
        return self.ranks[0].get_scalars()

    # a partition call on any member steps every rank
    def constr(self): self.ranks[0].constr()
    def homsol(self): self.ranks[0].homsol()
    def qcomp_ocean(self): self.ranks[0].qcomp_ocean()
    def ocean_step(self): self.ranks[0].ocean_step()
    def tlavg_ocean(self): self.ranks[0].tlavg_ocean()
    def run(self, a, b): self.ranks[0].run(a, b)

    def xforc(self):
        for m in self.ranks:
            m.xforc()

    # the diagnostic sums act on each rank's own rows
    def tavini(self):
        for m in self.ranks:
            m.tavini()

    def tavocn(self):
        for m in self.ranks:
            m.tavocn()

    def avg_ocn_k247(self):
        for m in self.ranks:
            m.avg_ocn_k247()

    def tav_counts(self):
        return self.ranks[0].tav_counts()

    def monnc_ocean(self):
        return self.ranks[0].monnc_ocean()      # a partition call: the ranks' row sums are combined

    def qocdiag(self, nsko):
        out = None
        for m in self.ranks:      # each rank fills the sub-sampled rows it owns
            out = m.qocdiag(nsko, out)
        return out

    def get_field_sub(self, name, nsk):
        out = None
        for m in self.ranks:      # each rank fills the sub-sampled rows it owns
            out = m.get_field_sub(name, nsk, out)
        return out

    def valids(self):
        """combine the per-slab reports as a multi-process driver would (min / max / sum)"""
        reps = [m.valids() for m in self.ranks]
        out = QgcmValidsReport()
        for name, typ in out._fields_:
            vals = [getattr(r, name) for r in reps]
            if name == "hfbad":
                for k in range(len(out.hfbad)):
                    out.hfbad[k] = sum(v[k] for v in vals)
            elif name.endswith("min") or name in ("hfmint", "hfmini", "hfminb"):
                setattr(out, name, min(vals))
            elif name.endswith("max") or name in ("hfmaxt", "hfmaxi", "hfmaxb"):
                setattr(out, name, max(vals))
        out.solnok = int(all(r.solnok for r in reps) and max(out.hfbad) <= 20.0)
        return out

    def sync(self):
        for m in self.ranks:
            m.sync()

    def close(self):
        for m in reversed(self.ranks):
            m.close()

    def launch_count(self):
        return sum(m.launch_count() for m in self.ranks)

"""ctypes mirror of include/qgcm_b200.h.

The two structs are parsed from the header itself so the Python side cannot drift
from the C ABI; qgcm_create additionally checks ``struct_bytes``.
"""
import ctypes as C
import os
import re

HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "qgcm_b200.h")
NLMAX = 9
ABI_VERSION = 1

FLAGS = {
    "ocean_only": 1 << 0,
    "atmos_only": 1 << 1,
    "cyclic_ocean": 1 << 2,
    "sb_hflux": 1 << 3,
    "nb_hflux": 1 << 4,
    "tau_udiff": 1 << 5,
    "ocnc_avg_k247": 1 << 6,
}


def _parse_struct(text, name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S)
    if not m:
        raise RuntimeError("struct %s not found in %s" % (name, HEADER))
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        ctype, rest = stmt.split(" ", 1)
        base = {"int32_t": C.c_int32, "double": C.c_double, "int64_t": C.c_int64}[ctype]
        for decl in rest.split(","):
            decl = decl.strip()
            am = re.match(r"(\w+)\[(.*)\]$", decl)
            if am:
                n = eval(am.group(2).replace("QGCM_NLMAX", str(NLMAX)))
                fields.append((am.group(1), base * n))
            else:
                fields.append((decl, base))
    return fields


with open(HEADER) as _f:
    _TEXT = _f.read()


class QgcmConfig(C.Structure):
    _fields_ = _parse_struct(_TEXT, "qgcm_config")


class QgcmScalars(C.Structure):
    _fields_ = _parse_struct(_TEXT, "qgcm_scalars")

    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


class QgcmValidsReport(C.Structure):
    _fields_ = _parse_struct(_TEXT, "qgcm_valids_report")

    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


def declared_functions():
    """names of every extern "C" function the header declares"""
    return sorted(set(re.findall(r"\b(qgcm_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", _TEXT, flags=re.S))))


class QgcmMonitorOcean(C.Structure):
    _fields_ = _parse_struct(_TEXT, "qgcm_monitor_ocean")

    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


class QgcmMonitorAtmos(C.Structure):
    _fields_ = _parse_struct(_TEXT, "qgcm_monitor_atmos")

    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/qgcm_b200.h declares, agrees with the ctypes mirror on struct sizes, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest


def test_header_declares_reference_procedures(qg):
    names = qg.abi.declared_functions()
    for proc in ("xforc", "oml", "qgostep", "ocinvq", "ocqbdy", "aml", "qgastep", "atinvq", "atqzbd",
                 "tlavg_ocean", "tlavg_atmos", "ocean_step", "run", "homsol", "constr", "helmholtz"):
        assert "qgcm_" + proc in names


def test_header_cites_reference_lines():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "qgcm_b200.h")).read()
    assert len(re.findall(r"src/[\w\-\.]+\.[Ff]:\d+", text)) >= 20


def test_library_exports_every_declared_symbol(qg):
    lib = qg.load_library()
    missing = [n for n in qg.abi.declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.qgcm_abi_version() == qg.abi.ABI_VERSION


def test_struct_layout_is_plain_c(qg):
    # only int32/double members: no padding surprises between C and ctypes
    assert C.sizeof(qg.QgcmConfig) % 8 == 0
    assert C.sizeof(qg.QgcmScalars) % 8 == 0
    n_int = sum(1 for _, t in qg.QgcmConfig._fields_ if t is C.c_int32 or getattr(t, "_type_", None) is C.c_int32)
    assert n_int >= 16


def test_no_cpu_fallback(qg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = qg.build_config(qg.named_config("dg_oo").scaled(6, 5))
    with pytest.raises(RuntimeError) as e:
        qg.Model(cfg)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_abi_mismatch_is_rejected(qg):
    cfg = qg.build_config(qg.named_config("dg_oo").scaled(6, 5))
    cfg.struct_bytes = 12
    with pytest.raises(RuntimeError):
        qg.Model(cfg)


def test_product_never_imports_oracle():
    """the package and bench's product path must not reference oracle/"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "q-gcm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"pyorc|liborc|\borc_|oracle/", text), f


def test_header_and_library_work_from_plain_c(qg, tmp_path):
    """examples/abi_smoke.c: the header is valid C99, the library links from C, the error
    convention works, and the struct sizes the C compiler sees are the ones ctypes computes"""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "q-gcm_b200", "csrc")
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "examples", "abi_smoke.c"), "-L" + libdir, "-lqgcm_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    sizes = dict(re.findall(r"(\w+) (\d+)", r.stdout.split("sizeof:")[1].splitlines()[0]))
    assert int(sizes["config"]) == C.sizeof(qg.QgcmConfig)
    assert int(sizes["scalars"]) == C.sizeof(qg.QgcmScalars)
    assert int(sizes["valids"]) == C.sizeof(qg.QgcmValidsReport)
    assert int(sizes["monitor_ocean"]) == C.sizeof(qg.QgcmMonitorOcean)
    assert int(sizes["monitor_atmos"]) == C.sizeof(qg.QgcmMonitorAtmos)
    assert "ABI mismatch" in r.stdout


def test_fortran_mirror_of_the_config_matches_the_header(qg):
    """no Fortran compiler exists here, so the bind(C) mirror of qgcm_config in
    integration/qgcm_types.f90 (the file INTEGRATION.md tells a maintainer to compile) is checked
    textually: same members, same order, same extents as the header"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "integration", "qgcm_types.f90")).read()
    block = re.search(r"type, bind\(C\) :: qgcm_config(.*?)end type", text, re.S).group(1)
    members = []
    for line in block.splitlines():
        line = line.split("!")[0].strip()
        if "::" not in line:
            continue
        ftype, names = line.split("::")
        base = C.c_int32 if "c_int32_t" in ftype else C.c_double
        for decl in re.findall(r"(\w+)(?:\(([^)]*)\))?", names):
            name, ext = decl
            if not name:
                continue
            n = 1
            if ext:
                n = eval(ext.replace("QGCM_NLMAX", str(qg.abi.NLMAX)))
            members.append((name.lower(), base, n))
    header = []
    for name, typ in qg.QgcmConfig._fields_:
        is_array = hasattr(typ, "_length_")
        header.append((name.lower(), typ._type_ if is_array else typ, typ._length_ if is_array else 1))
    assert members == header


def qg_abi_functions():
    import _pkg
    return _pkg.load().abi.declared_functions()


def test_generated_fortran_module_is_current():
    """integration/qgcm_types.f90 (bind(C) types and one interface per entry point) is generated
    from the header; a stale copy fails here"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_fortran_types.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(root, "integration", "qgcm_types.f90")).read()
    assert r.stdout.count("end function") == len(qg_abi_functions()) and max(len(l) for l in r.stdout.splitlines()) <= 132

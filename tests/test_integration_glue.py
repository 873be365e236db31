"""Static checks of the Fortran glue (integration/qgcm_cuda_mod.F90), which no compiler in this image can
build: every qgcm_* entry point it calls is declared in include/qgcm_b200.h with that many arguments, every
field name it passes is a field the library registers, every struct member it touches exists in the generated
bind(C) types, and (where /root/reference is present) every reference variable it names is declared in the
reference's *_data.F modules."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = open(os.path.join(ROOT, "integration", "qgcm_cuda_mod.F90")).read()
CODE = "\n".join(l.split("!")[0] if not l.lstrip().startswith("#") else "" for l in GLUE.splitlines())
HEADER = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "qgcm_b200.h")).read(), flags=re.S)
TYPES = open(os.path.join(ROOT, "integration", "qgcm_types.f90")).read()


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def calls(name_re):
    """(name, [args]) of every call name(...) in the glue"""
    for m in re.finditer(r"\b(%s)\s*\(" % name_re, CODE):
        depth, i = 1, m.end()
        while depth:
            depth += CODE[i] == "("
            depth -= CODE[i] == ")"
            i += 1
        yield m.group(1), split_args(CODE[m.end():i - 1])


def test_entry_points_and_argument_counts():
    protos = {n: len([a for a in args.split(",") if a.strip() and a.strip() != "void"])
              for n, args in re.findall(r"\b(qgcm_\w+)\s*\(([^)]*)\)\s*;", HEADER)}
    seen = set()
    for name, args in calls(r"qgcm_\w+"):
        if name == "qgcm_host_register_array":
            name = "qgcm_host_register"
        if name in ("qgcm_cuda", "qgcm_types", "qgcm_config", "qgcm_scalars"):
            continue
        assert name in protos, "the glue calls %s, which include/qgcm_b200.h does not declare" % name
        if "bind" in "".join(args):        # the interface block itself
            continue
        assert len(args) == protos[name], (name, args, protos[name])
        seen.add(name)
    assert {"qgcm_create", "qgcm_set_field", "qgcm_get_field", "qgcm_get_scalars", "qgcm_constr", "qgcm_homsol", "qgcm_xforc",
            "qgcm_qcomp_ocean", "qgcm_host_register", "qgcm_destroy"} <= seen


def test_field_names_are_registered_by_the_library():
    src = "".join(open(f).read() for f in glob.glob(os.path.join(ROOT, "q-gcm_b200", "csrc", "*.cu")))
    fields = set(re.findall(r'add_field\(m(?:d)?, "(\w+)"', src))
    for grp in re.findall(r"for \(const char \*n : \{([^}]*)\}\) add_field", src):
        fields |= set(re.findall(r'"(\w+)"', grp))
    used = {a[0].strip().strip("'") for n, a in calls(r"put|get") if a and a[0].strip().startswith("'")}
    assert len(used) >= 30
    assert used <= fields, "not library fields: %s" % sorted(used - fields)
    # the array passed is the module variable of the same name
    for n, a in calls(r"put|get"):
        if a and a[0].strip().startswith("'"):
            assert a[1].strip() == a[0].strip().strip("'"), a


def test_struct_members_exist_in_the_generated_types():
    def members(tname):
        body = re.search(r"type, bind\(C\) :: %s\b(.*?)end type" % tname, TYPES, re.S).group(1)
        return set(re.findall(r"::\s*(\w+)", body))
    cfgm, scm = members("qgcm_config"), members("qgcm_scalars")
    for m in set(re.findall(r"cfg%(\w+)", CODE)):
        assert m in cfgm, "qgcm_config has no member %s" % m
    for m in set(re.findall(r"\bs%(\w+)", CODE)):
        assert m in scm, "qgcm_scalars has no member %s" % m
    # every member of the config is filled
    missing = {m for m in cfgm if not re.search(r"cfg%%%s\b" % m, CODE)}
    assert not missing, "gpu_fill_config leaves qgcm_config members unset: %s" % sorted(missing)
    for flag in re.findall(r"QGCM_[A-Z0-9_]+", CODE):
        assert re.search(r"\b%s\b" % flag, TYPES), flag


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="/root/reference is not on this box")
def test_reference_variables_exist_in_the_reference_modules():
    decl = ""
    for f in glob.glob("/root/reference/src/*_data.F"):
        decl += open(f, errors="replace").read().lower()
    names = set()
    for n, a in calls(r"put|get"):
        if a and a[0].strip().startswith("'"):
            names.add(a[1].strip())
    for lhs, rhs in re.findall(r"cfg%(\w+)\s*=\s*(\w+)\b", CODE):
        if rhs not in ("int", "ior", "device", "nranks", "rank") and not rhs[0].isdigit() and not rhs.startswith("QGCM"):
            names.add(rhs)
    for lhs in re.findall(r"^\s*(\w+)\s*=\s*s%", CODE, re.M) + re.findall(r";\s*(\w+)\s*=\s*s%", CODE):
        names.add(lhs)
    assert len(names) > 80
    for n in sorted(names):
        assert re.search(r"\b%s\b" % n.lower(), decl), "the reference's *_data.F modules declare no %s" % n

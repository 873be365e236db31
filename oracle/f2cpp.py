#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- f2cpp: a source-to-source translator from the fixed-form Fortran
77/90 subset the reference's hot path is written in to C++.

Why: the reference is Fortran and neither this image nor the GPU box has a Fortran compiler
(profiles/r02_fortran_probe.txt).  The hand-written restatement in oracle/*.cpp could share a
misreading with the CUDA code because one reader wrote both.  This tool removes the reader: it
translates the reference's OWN source files, read where they lie under /root/reference/src,
statement by statement, with no knowledge of what they compute; the result is compiled by g++
into oracle/_ref/libqgcmref.so (oracle/Makefile, target `ref`).  Nothing of the reference is
copied into the repository: the generated C++ lives only under the git-ignored oracle/_ref/.
The restatement is then checked against the translated reference (tests/test_reference_pin.py)
and golden vectors produced by it are committed under tests/golden/.

What is translated: program units (MODULE with PARAMETERs, module variables and CONTAINed
procedures; external SUBROUTINE / FUNCTION), type declarations with explicit-shape, lower-bound
and assumed-size arrays, PARAMETER, SAVE, DATA, INTENT, assignment, DO / labelled DO / ENDDO /
CONTINUE, block and logical IF, GOTO, CALL (sequence association: an array element as actual
argument passes the address), RETURN, STOP and the numeric intrinsics the sources use.  cpp
directives are resolved first with the reference's own configuration macros.  OpenMP directives
are comments in fixed form and are dropped: the translation is the serial program, which is
one of the orderings an OpenMP REDUCTION may legally produce.  PRINT / WRITE / FORMAT produce
no code (diagnostic text only).

Semantics kept: column-major storage, 1-based (or declared) lower bounds, integer division,
by-reference argument passing, PARAMETER evaluation order, zero-initialised static storage
(gfortran puts SAVE/module arrays in .bss), DO trip counts fixed at loop entry, evaluation in
source order with Fortran's operator precedence (the generated expressions are fully
parenthesised; g++ is run with -ffp-contract=off and without -ffast-math).

Not a general Fortran compiler: anything outside the subset raises an error naming the line.
"""
import os
import re
import subprocess
import sys

# --------------------------------------------------------------------------------------
# source -> logical statements
# --------------------------------------------------------------------------------------


def run_cpp(path, defines):
    """resolve #ifdef / #ifndef / #if defined(...) with the reference's configuration macros"""
    cmd = ["cpp", "-traditional-cpp", "-P", "-undef", "-w"] + ["-D%s" % d for d in defines] + [path]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("cpp failed on %s: %s" % (path, out.stderr[:500]))
    return out.stdout.splitlines()


def strip_inline_comment(line):
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


class Stmt:
    __slots__ = ("label", "text", "where")

    def __init__(self, label, text, where):
        self.label, self.text, self.where = label, text, where

    def __repr__(self):
        return "%s: %s" % (self.where, self.text)


def logical_statements(path, defines):
    lines = run_cpp(path, defines)
    stmts = []
    cur = None
    for n, raw in enumerate(lines, 1):
        line = raw.rstrip("\n").expandtabs(8)[:80]
        if not line.strip():
            continue
        if line[0] in "cC*!#":
            continue
        line = strip_inline_comment(line).rstrip()
        if not line.strip():
            continue
        if len(line) > 5 and line[5] not in " 0" and not line[:5].strip():
            if cur is None:
                raise RuntimeError("%s:%d: continuation without a statement" % (path, n))
            cur.text += " " + line[6:].strip()
            continue
        label = line[:5].strip()
        body = line[6:].strip() if len(line) > 6 else ""
        if not body and not label:
            continue
        cur = Stmt(label, body, "%s:%d" % (os.path.basename(path), n))
        stmts.append(cur)
    # lower-case outside character literals; fold the embedded blanks of long numeric literals
    for s in stmts:
        s.text = normalise(s.text)
    return stmts


def normalise(text):
    out, q = [], None
    for ch in text:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
                out.append(ch)
            else:
                out.append(ch.lower())
    t = "".join(out)
    # fixed form ignores blanks: "0.8660254037 84438647 d0" is one literal
    prev = None
    while prev != t:
        prev = t
        t = re.sub(r"(\d*\.\d+|\d+\.\d*) (\d+)", r"\1\2", t)
        t = re.sub(r"(\d*\.\d+|\d+\.\d*) ([de][+-]?\d)", r"\1\2", t)
    return t


# --------------------------------------------------------------------------------------
# expressions
# --------------------------------------------------------------------------------------
TOKEN = re.compile(r"""
    (?P<num>(\d+\.\d*|\.\d+|\d+)([de][+-]?\d+)?)
  | (?P<dotop>\.(eq|ne|lt|le|gt|ge|and|or|not|eqv|neqv|true|false)\.)
  | (?P<name>[a-z_][a-z0-9_]*)
  | (?P<str>'([^']|'')*'|"([^"]|"")*")
  | (?P<op>\*\*|==|/=|<=|>=|//|[-+*/(),=<>:])
  | (?P<ws>\s+)
""", re.X)


def tokenize(text, where):
    toks, pos = [], 0
    while pos < len(text):
        m = TOKEN.match(text, pos)
        if not m:
            raise RuntimeError("%s: cannot tokenise %r at %r" % (where, text, text[pos:pos + 20]))
        pos = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        v = m.group(k)
        # "1.eq.2": the number regex must not swallow the dot of a dotted operator
        if k == "num" and v.endswith(".") and re.match(r"(eq|ne|lt|le|gt|ge|and|or|not|eqv|neqv)\.", text[pos:]):
            v = v[:-1]
            pos -= 1
        toks.append((k, v))
    return toks


class Node:
    pass


class Num(Node):
    def __init__(self, text):
        self.text = text
        self.is_int = re.fullmatch(r"\d+", text) is not None


class Name(Node):
    def __init__(self, name):
        self.name = name


class Ref(Node):       # name(args): array element or function call
    def __init__(self, name, args):
        self.name, self.args = name, args


class Un(Node):
    def __init__(self, op, a):
        self.op, self.a = op, a


class Bin(Node):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b


class Str(Node):
    def __init__(self, text):
        self.text = text


class Logical(Node):
    def __init__(self, v):
        self.v = v


class Range(Node):     # lo:hi in declarations; '*' for assumed size
    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi


REL = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=",
       "==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}


class Parser:
    def __init__(self, toks, where):
        self.t, self.i, self.where = toks, 0, where

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, v):
        if self.peek()[1] == v:
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise RuntimeError("%s: expected %r, found %r" % (self.where, v, self.peek()[1]))

    def done(self):
        return self.i >= len(self.t)

    # precedence climbing, lowest first
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        a = self.p_or()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            a = Bin("==" if op == ".eqv." else "!=", a, self.p_or())
        return a

    def p_or(self):
        a = self.p_and()
        while self.peek()[1] == ".or.":
            self.next()
            a = Bin("||", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek()[1] == ".and.":
            self.next()
            a = Bin("&&", a, self.p_not())
        return a

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.next()
            return Un("!", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_add()
        if self.peek()[1] in REL:
            op = self.next()[1]
            return Bin(REL[op], a, self.p_add())
        return a

    def p_add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = Un(op, self.p_mul())
        else:
            a = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = Bin(op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            a = Bin(op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek()[1] == "**":
            self.next()
            # right associative; a unary minus may follow the operator (x**-2)
            if self.peek()[1] in ("+", "-"):
                op = self.next()[1]
                b = Un(op, self.p_pow())
            else:
                b = self.p_pow()
            return Bin("**", a, b)
        return a

    def p_primary(self):
        k, v = self.next()
        if k == "num":
            return Num(v)
        if k == "str":
            return Str(v)
        if k == "dotop" and v in (".true.", ".false."):
            return Logical(v == ".true.")
        if v == "(":
            e = self.expr()
            self.expect(")")
            return Un("()", e)
        if k == "name":
            if self.peek()[1] == "(":
                self.next()
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.arg())
                        if self.accept(")"):
                            break
                        self.expect(",")
                return Ref(v, args)
            return Name(v)
        raise RuntimeError("%s: unexpected token %r" % (self.where, v))

    def arg(self):
        # subscripts of a declaration may be lo:hi or *
        if self.peek()[1] == "*":
            nxt = self.t[self.i + 1][1] if self.i + 1 < len(self.t) else None
            if nxt in (")", ","):
                self.next()
                return Range(None, None)
        e = self.expr()
        if self.accept(":"):
            if self.peek()[1] == "*":
                self.next()
                return Range(e, None)
            return Range(e, self.expr())
        return e


def parse_expr(text, where):
    p = Parser(tokenize(text, where), where)
    e = p.expr()
    if not p.done():
        raise RuntimeError("%s: trailing tokens in expression %r" % (where, text))
    return e


# --------------------------------------------------------------------------------------
# program structure
# --------------------------------------------------------------------------------------
CTYPE = {"integer": "int", "double": "double", "logical": "bool", "character": "std::string", "real": "float"}


class Sym:
    def __init__(self, name, ftype):
        self.name, self.ftype = name, ftype
        self.dims = None          # list of Range (lo may be None = 1; hi None = assumed size)
        self.param = None         # Node: PARAMETER value
        self.dummy = False
        self.save = False
        self.module = None
        self.data = None          # list of Node: DATA values
        self.where = None

    @property
    def ctype(self):
        return CTYPE[self.ftype]


class Unit:
    def __init__(self, kind, name, args, where):
        self.kind, self.name, self.args, self.where = kind, name, args, where
        self.syms = {}
        self.order = []           # declaration order (PARAMETERs are evaluated in it)
        self.body = []            # executable statements
        self.uses = []
        self.contains = []
        self.rtype = None         # functions
        self.parent = None
        self.save_all = False
        self.param_order = []


INTRINSICS = {"abs", "dabs", "iabs", "sign", "dsign", "isign", "max", "min", "dmax1", "dmin1", "max0", "min0", "amax1", "amin1",
              "mod", "sqrt", "dsqrt", "exp", "dexp", "log", "dlog", "alog", "log10", "sin", "dsin", "cos", "dcos", "tan", "dtan",
              "atan", "datan", "atan2", "datan2", "asin", "acos", "sinh", "cosh", "tanh", "dble", "dfloat", "float", "real",
              "int", "nint", "idnint", "idint", "ifix", "aint", "anint", "dim", "trim", "len", "index"}

DECL_RE = re.compile(r"^(integer|double\s*precision|real\s*\*\s*8|real|logical|character)\b(.*)$")


def split_top(text, sep=","):
    out, depth, cur, q = [], 0, [], None
    for ch in text:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == "(":
            depth += 1
            cur.append(ch)
        elif ch == ")":
            depth -= 1
            cur.append(ch)
        elif ch == sep and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur).strip())
    return out


class Program:
    def __init__(self):
        self.modules = {}
        self.procs = {}           # every subroutine / function by name
        self.units = []

    # ---- parsing -----------------------------------------------------------------
    def add_file(self, path, defines):
        stmts = logical_statements(path, defines)
        i = 0
        stack = []
        cur = None
        while i < len(stmts):
            s = stmts[i]
            t = s.text
            i += 1
            m = re.match(r"^module\s+(\w+)$", t)
            if m and not t.startswith("module procedure"):
                cur = Unit("module", m.group(1), [], s.where)
                self.modules[cur.name] = cur
                self.units.append(cur)
                stack = [cur]
                continue
            m = re.match(r"^(?:(integer|double\s*precision|logical|real)\s+)?(subroutine|function|program)\s+(\w+)\s*(?:\((.*)\))?$", t)
            if m:
                args = [a.strip() for a in (m.group(4) or "").split(",") if a.strip()]
                u = Unit(m.group(2), m.group(3), args, s.where)
                if m.group(1):
                    u.rtype = "double" if m.group(1).startswith("double") else m.group(1)
                if stack and stack[-1].kind == "module":
                    u.parent = stack[-1]
                    stack[-1].contains.append(u)
                if u.kind != "program":
                    if u.name in self.procs:
                        raise RuntimeError("%s: duplicate procedure %s" % (s.where, u.name))
                    self.procs[u.name] = u
                self.units.append(u)
                stack.append(u)
                cur = u
                continue
            if re.match(r"^end\s*(module|subroutine|function|program)?(\s+\w+)?$", t) or t == "end":
                if not stack:
                    raise RuntimeError("%s: END without a unit" % s.where)
                stack.pop()
                cur = stack[-1] if stack else None
                continue
            if cur is None:
                raise RuntimeError("%s: statement outside a program unit: %s" % (s.where, t))
            if t == "contains":
                continue
            self.statement(cur, s)

    def statement(self, u, s):
        t = s.text
        if t.startswith("use "):
            u.uses.append(re.match(r"use\s+(\w+)", t).group(1))
            return
        if t.startswith("implicit") or t in ("public", "private") or t.startswith("public ") or t.startswith("private "):
            return
        if t == "save":
            u.save_all = True
            return
        if t.startswith("save ") or t.startswith("save::"):
            for n in split_top(re.sub(r"^save\s*(::)?", "", t)):
                self.sym(u, n.strip(), None, s).save = True
            return
        if t.startswith("external") or t.startswith("intrinsic"):
            return
        m = re.match(r"^parameter\s*\((.*)\)$", t)
        if m:
            for item in split_top(m.group(1)):
                name, val = item.split("=", 1)
                sy = self.sym(u, name.strip(), None, s)
                sy.param = parse_expr(val.strip(), s.where)
                u.param_order.append(sy.name)
            return
        m = re.match(r"^data\s+(.*)$", t)
        if m and not re.match(r"^data\s*=", t) and not re.match(r"^data\s*\(", t):
            self.data_stmt(u, m.group(1), s)
            return
        m = DECL_RE.match(t)
        if m and not re.match(r"^(real|integer|logical)\s*=", t) and not re.match(r"^(real|integer|logical)\s*\(.*\)\s*=", t):
            self.declaration(u, m.group(1), m.group(2), s)
            return
        if t.startswith("format"):
            return
        u.body.append(s)

    def sym(self, u, name, ftype, s):
        sy = u.syms.get(name)
        if sy is None:
            sy = Sym(name, ftype or "?")
            sy.where = s.where
            u.syms[name] = sy
            u.order.append(name)
        elif ftype:
            sy.ftype = ftype
        return sy

    def declaration(self, u, tword, rest, s):
        ftype = "double" if tword.startswith("double") or "*" in tword else tword
        rest = rest.strip()
        attrs = []
        if rest.startswith("*"):          # character*72 ...
            m = re.match(r"^\*\s*(\(\s*\*\s*\)|\d+)\s*(.*)$", rest)
            rest = m.group(2)
        if rest.startswith("("):          # character (len=*) etc.
            depth = 0
            for k, ch in enumerate(rest):
                depth += ch == "("
                depth -= ch == ")"
                if depth == 0:
                    rest = rest[k + 1:].strip()
                    break
        if "::" in rest:
            a, rest = rest.split("::", 1)
            attrs = [x.strip() for x in split_top(a.lstrip(","))]
        dim_attr = None
        is_param = False
        for a in attrs:
            if a.startswith("dimension"):
                dim_attr = a[len("dimension"):].strip()
            if a == "parameter":
                is_param = True
        for item in split_top(rest):
            init = None
            if "=" in item and is_param:
                item, init = item.split("=", 1)
            item = item.strip()
            m = re.match(r"^(\w+)\s*(\(.*\))?(\s*\*\s*\d+)?$", item)
            if not m:
                raise RuntimeError("%s: cannot parse declaration item %r" % (s.where, item))
            sy = self.sym(u, m.group(1), ftype, s)
            dims = m.group(2) or dim_attr
            if dims:
                ref = parse_expr("x" + dims, s.where)
                sy.dims = [a if isinstance(a, Range) else Range(None, a) for a in ref.args]
            if "save" in attrs:
                sy.save = True
            if init is not None:
                sy.param = parse_expr(init.strip(), s.where)
                u.param_order.append(sy.name)

    def data_stmt(self, u, text, s):
        # data obj [, obj ...] / v1, v2, ... / [[,] obj ... / ... /]: objects are scalars, whole
        # arrays (filled in element order) or array elements with constant subscripts
        for m in re.finditer(r"([^/]+)/([^/]*)/\s*,?", text):
            vals = []
            for v in split_top(m.group(2)):
                r = re.match(r"^(\d+)\s*\*\s*(.*)$", v)
                if r:
                    vals += [parse_expr(r.group(2), s.where)] * int(r.group(1))
                else:
                    vals.append(parse_expr(v, s.where))
            for obj in split_top(m.group(1).strip().lstrip(",")):
                e = parse_expr(obj, s.where)
                name = e.name
                sy = self.sym(u, name, None, s)
                sy.save = True
                if sy.data is None:
                    sy.data = []
                if isinstance(e, Ref):
                    if not vals:
                        raise RuntimeError("%s: DATA has too few values" % s.where)
                    sy.data.append((e.args, vals.pop(0)))
                elif sy.dims is not None:
                    n = 1
                    for rng in sy.dims:
                        lo = int(rng.lo.text) if rng.lo is not None else 1
                        n *= int(rng.hi.text) - lo + 1
                    for k in range(n):
                        sy.data.append((k, vals.pop(0)))
                else:
                    sy.data.append((None, vals.pop(0)))
            if vals:
                raise RuntimeError("%s: DATA has too many values" % s.where)


# --------------------------------------------------------------------------------------
# code generation
# --------------------------------------------------------------------------------------
class Emitter:
    def __init__(self, prog, wanted, stubs):
        self.p = prog
        self.wanted = wanted      # procedures to emit (closure over calls is added)
        self.stubs = stubs        # external procedures supplied by hand-written C++ (LAPACK)
        self.out = []
        self.tmp = 0

    # ---- symbol lookup
    def lookup(self, u, name):
        if name in u.syms and u.syms[name].ftype != "?":
            return u.syms[name], u
        if name in u.syms and (u.syms[name].param is not None or u.syms[name].dims is not None):
            return u.syms[name], u
        scopes = []
        if u.parent:
            scopes.append(u.parent)
        for mname in u.uses + (u.parent.uses if u.parent else []):
            if mname in self.p.modules:
                scopes.append(self.p.modules[mname])
        seen = set()
        while scopes:
            mod = scopes.pop(0)
            if mod.name in seen:
                continue
            seen.add(mod.name)
            if name in mod.syms:
                return mod.syms[name], mod
            for mname in mod.uses:
                if mname in self.p.modules:
                    scopes.append(self.p.modules[mname])
        if name in u.syms:
            return u.syms[name], u
        return None, None

    def cname(self, sy, owner):
        if owner.kind == "module":
            return "%s_" % sy.name
        return "%s_" % sy.name

    # ---- expressions
    def etype(self, u, e):
        """'int', 'double', 'bool' or 'str'"""
        if isinstance(e, Num):
            return "int" if e.is_int else "double"
        if isinstance(e, Str):
            return "str"
        if isinstance(e, Logical):
            return "bool"
        if isinstance(e, Name):
            if u.kind == "function" and e.name == u.name:
                rt = u.rtype or u.syms[u.name].ftype
                return "int" if rt == "integer" else "bool" if rt == "logical" else "double"
            sy, _ = self.lookup(u, e.name)
            if sy is None:
                raise RuntimeError("%s: undeclared name %s" % (u.where, e.name))
            return {"integer": "int", "double": "double", "logical": "bool", "character": "str", "real": "double"}[sy.ftype]
        if isinstance(e, Ref):
            sy, _ = self.lookup(u, e.name)
            if sy is not None and sy.dims is not None:
                return {"integer": "int", "double": "double", "logical": "bool", "character": "str", "real": "double"}[sy.ftype]
            if e.name in self.p.procs and self.p.procs[e.name].kind == "function":
                f = self.p.procs[e.name]
                rt = f.rtype or (f.syms[f.name].ftype if f.name in f.syms else "double")
                return "int" if rt == "integer" else "bool" if rt == "logical" else "double"
            if e.name in ("int", "nint", "idnint", "idint", "ifix", "iabs", "isign", "max0", "min0", "len", "index"):
                return "int"
            if e.name in ("abs", "sign", "max", "min", "mod", "dim"):
                ts = [self.etype(u, a) for a in e.args]
                return "int" if all(t == "int" for t in ts) else "double"
            if e.name in INTRINSICS:
                return "double"
            raise RuntimeError("%s: unknown function or array %s" % (u.where, e.name))
        if isinstance(e, Un):
            if e.op == "!":
                return "bool"
            return self.etype(u, e.a)
        if isinstance(e, Bin):
            if e.op in ("==", "!=", "<", "<=", ">", ">=", "&&", "||"):
                return "bool"
            ta, tb = self.etype(u, e.a), self.etype(u, e.b)
            if e.op == "**":
                return ta if tb == "int" else "double"
            return "int" if ta == "int" and tb == "int" else "double"
        raise RuntimeError("etype: %r" % e)

    def index(self, u, sy, args, where):
        if len(args) != len(sy.dims):
            raise RuntimeError("%s: rank mismatch on %s" % (where, sy.name))
        off = None
        # offset = (i1-lo1) + n1*((i2-lo2) + n2*(...)), built from the last dimension inwards
        for d in range(len(args) - 1, -1, -1):
            rng = sy.dims[d]
            lo = self.ex(u, rng.lo) if rng.lo is not None else "1"
            term = "((%s)-(%s))" % (self.ex(u, args[d]), lo)
            if off is None:
                off = term
            else:
                if rng.hi is None:
                    raise RuntimeError("%s: assumed size in a non-final dimension of %s" % (where, sy.name))
                ext = "((%s)-(%s)+1)" % (self.ex(u, rng.hi), lo)
                off = "(%s+(long)%s*%s)" % (term, ext, off)
        return off

    def ex(self, u, e):
        if isinstance(e, Num):
            t = e.text
            if e.is_int:
                return t
            t = t.replace("d", "e")
            if "." not in t and "e" not in t:
                t += ".0"
            return t
        if isinstance(e, Str):
            inner = e.text[1:-1].replace("''", "'").replace('\\', '\\\\').replace('"', '\\"')
            return 'std::string("%s")' % inner
        if isinstance(e, Logical):
            return "true" if e.v else "false"
        if isinstance(e, Name):
            if u.kind == "function" and e.name == u.name:
                return "%s_res" % u.name
            sy, owner = self.lookup(u, e.name)
            if sy is None:
                raise RuntimeError("%s: undeclared name %s" % (u.where, e.name))
            if sy.dummy and sy.dims is None:
                return "(*%s_)" % sy.name if sy.ftype != "character" else "%s_" % sy.name
            return "%s_" % sy.name
        if isinstance(e, Ref):
            sy, owner = self.lookup(u, e.name)
            if sy is not None and sy.dims is not None:
                return "%s_[%s]" % (sy.name, self.index(u, sy, e.args, u.where))
            if e.name in self.p.procs and self.p.procs[e.name].kind == "function":
                return "%s_(%s)" % (e.name, ", ".join(self.actual(u, a, None) for a in e.args))
            if e.name in INTRINSICS:
                return self.intrinsic(u, e)
            raise RuntimeError("%s: unknown function or array %s" % (u.where, e.name))
        if isinstance(e, Un):
            if e.op == "()":
                return "(%s)" % self.ex(u, e.a)
            return "(%s%s)" % (e.op, self.ex(u, e.a))
        if isinstance(e, Bin):
            if e.op == "**":
                tb = self.etype(u, e.b)
                ta = self.etype(u, e.a)
                if tb == "int":
                    return "f2c_ipow<%s>(%s, %s)" % ("int" if ta == "int" else "double", self.ex(u, e.a), self.ex(u, e.b))
                return "std::pow((double)(%s), (double)(%s))" % (self.ex(u, e.a), self.ex(u, e.b))
            a, b = self.ex(u, e.a), self.ex(u, e.b)
            if e.op in ("==", "!=") and self.etype(u, e.a) == "str":
                return "(f2c_trim(%s) %s f2c_trim(%s))" % (a, e.op, b)
            return "(%s %s %s)" % (a, e.op, b)
        raise RuntimeError("ex: %r" % e)

    def intrinsic(self, u, e):
        n = e.name
        a = [self.ex(u, x) for x in e.args]
        ts = [self.etype(u, x) for x in e.args]
        allint = all(t == "int" for t in ts)
        if n in ("abs", "dabs", "iabs"):
            return "std::abs(%s)" % a[0]
        if n in ("sign", "dsign", "isign"):
            return ("f2c_isign(%s, %s)" if allint else "f2c_sign((double)(%s), (double)(%s))") % (a[0], a[1])
        if n in ("max", "dmax1", "max0", "amax1", "min", "dmin1", "min0", "amin1"):
            fn = "std::max" if n.startswith(("max", "dmax", "amax")) else "std::min"
            cast = "(int)" if allint else "(double)"
            r = "%s(%s)" % (cast, a[0])
            for x in a[1:]:
                r = "%s(%s, %s(%s))" % (fn, r, cast, x)
            return r
        if n == "mod":
            return ("((%s) %% (%s))" if allint else "std::fmod((double)(%s), (double)(%s))") % (a[0], a[1])
        if n == "dim":
            return "std::max(%s - %s, %s)" % (a[0], a[1], "0" if allint else "0.0")
        simple = {"sqrt": "sqrt", "dsqrt": "sqrt", "exp": "exp", "dexp": "exp", "log": "log", "dlog": "log", "alog": "log",
                  "log10": "log10", "sin": "sin", "dsin": "sin", "cos": "cos", "dcos": "cos", "tan": "tan", "dtan": "tan",
                  "atan": "atan", "datan": "atan", "asin": "asin", "acos": "acos", "sinh": "sinh", "cosh": "cosh", "tanh": "tanh"}
        if n in simple:
            return "std::%s((double)(%s))" % (simple[n], a[0])
        if n in ("atan2", "datan2"):
            return "std::atan2((double)(%s), (double)(%s))" % (a[0], a[1])
        if n in ("dble", "dfloat", "float", "real"):
            return "((double)(%s))" % a[0]
        if n in ("int", "idint", "ifix"):
            return "((int)(%s))" % a[0]
        if n in ("nint", "idnint"):
            return "((int)std::lround((double)(%s)))" % a[0]
        if n == "aint":
            return "std::trunc((double)(%s))" % a[0]
        if n == "anint":
            return "std::round((double)(%s))" % a[0]
        if n == "trim":
            return "f2c_trim(%s)" % a[0]
        raise RuntimeError("%s: intrinsic %s not supported" % (u.where, n))

    def actual(self, u, e, formal):
        """actual argument: everything goes by address (sequence association for arrays)"""
        if isinstance(e, Name):
            sy, owner = self.lookup(u, e.name)
            if sy is None:
                raise RuntimeError("%s: undeclared actual argument %s" % (u.where, e.name))
            if sy.dims is not None:
                return "%s_" % sy.name                       # whole array: its base address
            if sy.ftype == "character":
                return "%s_" % sy.name
            if sy.dummy:
                return "%s_" % sy.name                       # already a pointer
            if sy.param is not None and owner.kind != "module" and False:
                pass
            return "&%s_" % sy.name
        if isinstance(e, Ref):
            sy, owner = self.lookup(u, e.name)
            if sy is not None and sy.dims is not None:
                return "&%s_[%s]" % (sy.name, self.index(u, sy, e.args, u.where))
        if isinstance(e, Str):
            return self.ex(u, e)
        t = self.etype(u, e)
        if t == "str":
            return self.ex(u, e)
        return "f2c_tmp<%s>(%s)" % ({"int": "int", "double": "double", "bool": "bool"}[t], self.ex(u, e))

    # ---- units
    def closure(self):
        todo, seen = list(self.wanted), []
        while todo:
            n = todo.pop()
            if n in seen or n in self.stubs:
                continue
            if n not in self.p.procs:
                raise RuntimeError("procedure %s is neither in the translated sources nor a declared stub" % n)
            seen.append(n)
            u = self.p.procs[n]
            for s in u.body:
                for m in re.finditer(r"\bcall\s+(\w+)", s.text):
                    todo.append(m.group(1))
                for fn, f in self.p.procs.items():
                    if f.kind == "function" and re.search(r"\b%s\s*\(" % re.escape(fn), s.text):
                        todo.append(fn)
        return seen

    def emit(self):
        o = self.out
        o.append("// GENERATED by oracle/f2cpp.py from the reference's Fortran sources -- not part of the repository")
        o.append('#include "f2c_runtime.h"')
        procs = self.closure()
        mods = self.used_modules(procs)
        names = {}
        for mod in mods:
            for n in mod.order:
                if n in names and names[n] is not mod:
                    raise RuntimeError("module variable %s is defined in %s and %s" % (n, names[n].name, mod.name))
                names[n] = mod
        # module storage
        for mod in mods:
            o.append("// ---- module %s (%s)" % (mod.name, mod.where))
            for n in mod.order:
                sy = mod.syms[n]
                if sy.ftype == "?":
                    continue
                if sy.dims is not None:
                    o.append("static %s *%s_ = nullptr;" % (sy.ctype, n))
                else:
                    o.append("static %s %s_ = %s;" % (sy.ctype, n, "std::string()" if sy.ftype == "character" else "0"))
        # prototypes
        for n in procs:
            o.append(self.signature(self.p.procs[n]) + ";")
        for n, proto in self.stubs.items():
            o.append(proto)
        # module initialisation: PARAMETERs in order (unless overridden), then allocation
        o.append("static std::map<std::string, double> f2c_override;")
        o.append("static bool f2c_ready = false;")
        o.append("static void f2c_module_init() {")
        o.append("  if (f2c_ready) return;")
        for mod in mods:
            for n in mod.order:
                sy = mod.syms[n]
                if sy.param is not None:
                    o.append('  %s_ = f2c_override.count("%s") ? (%s)f2c_override["%s"] : (%s)(%s);' %
                             (n, n, sy.ctype, n, sy.ctype, self.ex(mod, sy.param)))
        for mod in mods:
            for n in mod.order:
                sy = mod.syms[n]
                if sy.dims is not None and sy.ftype != "?":
                    o.append("  %s_ = f2c_alloc<%s>(%s);" % (n, sy.ctype, self.total(mod, sy)))
                    for where, v in (sy.data or []):
                        o.append("  %s_[%s] = %s;" % (n, self.data_index(mod, sy, where), self.ex(mod, v)))
        o.append("  f2c_ready = true;")
        o.append("}")
        for n in procs:
            self.procedure(self.p.procs[n])
        self.api(mods, procs)
        return "\n".join(o) + "\n"

    def used_modules(self, procs):
        seen, order = set(), []

        def visit(mname):
            if mname in seen or mname not in self.p.modules:
                return
            seen.add(mname)
            for dep in self.p.modules[mname].uses:
                visit(dep)
            order.append(self.p.modules[mname])
        for n in procs:
            u = self.p.procs[n]
            if u.parent:
                for dep in u.parent.uses:
                    visit(dep)
                visit(u.parent.name)
            for dep in u.uses:
                visit(dep)
        return order

    def data_index(self, u, sy, where):
        if isinstance(where, int):
            return str(where)
        return self.index(u, sy, where, u.where)

    def total(self, u, sy):
        parts = []
        for rng in sy.dims:
            lo = self.ex(u, rng.lo) if rng.lo is not None else "1"
            parts.append("((long)(%s)-(%s)+1)" % (self.ex(u, rng.hi), lo))
        return "*".join(parts)

    def signature(self, u):
        args = []
        for a in u.args:
            sy = u.syms.get(a)
            if sy is None or sy.ftype == "?":
                raise RuntimeError("%s: dummy argument %s of %s has no type" % (u.where, a, u.name))
            if sy.ftype == "character":
                args.append("std::string %s_" % a)
            else:
                args.append("%s *%s_" % (sy.ctype, a))
        if u.kind == "function":
            rt = u.rtype or u.syms[u.name].ftype
            ret = CTYPE["double" if rt == "double" else rt]
        else:
            ret = "void"
        return "%s %s_(%s)" % (ret, u.name, ", ".join(args))

    def procedure(self, u):
        o = self.out
        for a in u.args:
            u.syms[a].dummy = True
        o.append("")
        o.append("// %s %s (%s)" % (u.kind, u.name, u.where))
        o.append(self.signature(u) + " {")
        o.append("  f2c_module_init();")
        # local parameters first (in order), then scalars, then arrays (their bounds may use the former)
        for n in u.order:
            sy = u.syms[n]
            if sy.dummy or sy.ftype == "?":
                continue
            if u.kind == "function" and n == u.name:
                continue
            if sy.dims is None:
                static = "static " if (sy.save or u.save_all or sy.data) else ""
                if sy.param is not None:
                    # not const: Fortran passes named constants by address like any other object
                    o.append("  %s %s_ = (%s)(%s);" % (sy.ctype, n, sy.ctype, self.ex(u, sy.param)))
                elif sy.ftype == "character":
                    o.append("  %sstd::string %s_;" % (static, n))
                elif sy.data:
                    o.append("  static %s %s_ = %s;" % (sy.ctype, n, self.ex(u, sy.data[0][1])))
                else:
                    # uninitialised in Fortran; zero here so that a use-before-set is at least repeatable
                    o.append("  %s%s %s_ = 0;" % (static, sy.ctype, n))
        for n in u.order:
            sy = u.syms[n]
            if sy.dummy or sy.ftype == "?" or sy.dims is None:
                continue
            if sy.save or u.save_all or sy.data:
                o.append("  static %s *%s_ = nullptr;" % (sy.ctype, n))
                o.append("  if (!%s_) {" % n)
                o.append("    %s_ = f2c_alloc<%s>(%s);" % (n, sy.ctype, self.total(u, sy)))
                for where, v in (sy.data or []):
                    o.append("    %s_[%s] = %s;" % (n, self.data_index(u, sy, where), self.ex(u, v)))
                o.append("  }")
            else:
                # automatic array: the reference gets stack garbage, we get zeros (repeatable)
                o.append("  std::vector<%s> %s_v((size_t)(%s)); %s *%s_ = %s_v.data();" %
                         ("char" if sy.ctype == "bool" else sy.ctype, n, self.total(u, sy), sy.ctype, n, n)
                         if sy.ctype != "bool" else
                         "  std::vector<char> %s_v((size_t)(%s)); bool *%s_ = reinterpret_cast<bool *>(%s_v.data());" %
                         (n, self.total(u, sy), n, n))
        if u.kind == "function":
            rt = u.rtype or u.syms[u.name].ftype
            o.append("  %s %s_res = 0;" % (CTYPE["double" if rt == "double" else rt], u.name))
        labels = set()
        for s in u.body:
            for m in re.finditer(r"\bgo\s*to\s+(\d+)", s.text):
                labels.add(m.group(1))
        self.blocks = []
        self.do_labels = []
        for s in u.body:
            self.exec_stmt(u, s, labels)
        if self.blocks:
            raise RuntimeError("%s: unterminated block in %s" % (u.where, u.name))
        if u.kind == "function":
            o.append("  return %s_res;" % u.name)
        o.append("}")

    def ind(self):
        return "  " * (1 + len(self.blocks))

    def exec_stmt(self, u, s, labels, nested=False):
        o = self.out
        t = s.text
        if s.label and not nested:
            if s.label in labels:
                o.append("L%s_%s: ;" % (s.label, u.name))
        try:
            self.exec_inner(u, s, t, labels)
        except RuntimeError as ex:
            raise RuntimeError("%s: %s\n    in statement: %s" % (s.where, ex, t))
        # a labelled statement may terminate labelled DO loops
        if s.label and not nested:
            while self.do_labels and self.do_labels[-1] == s.label:
                self.do_labels.pop()
                self.blocks.pop()
                o.append(self.ind() + "} }")

    def exec_inner(self, u, s, t, labels):
        o = self.out
        if t == "continue":
            return
        if t == "return":
            o.append(self.ind() + ("return %s_res;" % u.name if u.kind == "function" else "return;"))
            return
        if t.startswith("stop"):
            o.append(self.ind() + 'f2c_stop("%s");' % s.where)
            return
        if re.match(r"^(print\b|write\s*\()", t):
            return
        m = re.match(r"^go\s*to\s+(\d+)$", t)
        if m:
            o.append(self.ind() + "goto L%s_%s;" % (m.group(1), u.name))
            return
        m = re.match(r"^do\s+(?:(\d+)\s*,?\s*)?(\w+)\s*=(.*)$", t)
        if m:
            lab, var, rest = m.group(1), m.group(2), m.group(3)
            parts = split_top(rest)
            if len(parts) not in (2, 3):
                raise RuntimeError("bad DO bounds")
            v = self.ex(u, Name(var))
            a, b = self.ex(u, parse_expr(parts[0], s.where)), self.ex(u, parse_expr(parts[1], s.where))
            self.tmp += 1
            k = self.tmp
            if len(parts) == 3:
                c = self.ex(u, parse_expr(parts[2], s.where))
                o.append(self.ind() + "{ const int f2c_s%d = %s, f2c_e%d = %s; int f2c_n%d = (f2c_e%d - (%s) + f2c_s%d) / f2c_s%d;" % (k, c, k, b, k, k, a, k, k))
                o.append(self.ind() + "for (%s = %s; f2c_n%d > 0; --f2c_n%d, %s += f2c_s%d) {" % (v, a, k, k, v, k))
            else:
                o.append(self.ind() + "{ const int f2c_e%d = %s;" % (k, b))
                o.append(self.ind() + "for (%s = %s; %s <= f2c_e%d; ++%s) {" % (v, a, v, k, v))
            self.blocks.append("do")
            if lab:
                self.do_labels.append(lab)
            return
        if re.match(r"^end\s*do$", t):
            if not self.blocks or self.blocks[-1] != "do":
                raise RuntimeError("ENDDO without DO")
            self.blocks.pop()
            o.append(self.ind() + "} }")
            return
        m = re.match(r"^if\s*\((.*)$", t)
        if m:
            # find the matching parenthesis of the condition
            depth, k, q = 1, 0, None
            body = m.group(1)
            while k < len(body) and depth:
                ch = body[k]
                if q:
                    if ch == q:
                        q = None
                elif ch in "'\"":
                    q = ch
                elif ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                k += 1
            cond, rest = body[:k - 1], body[k:].strip()
            c = self.ex(u, parse_expr(cond, s.where))
            if rest == "then":
                o.append(self.ind() + "if (%s) {" % c)
                self.blocks.append("if")
                return
            o.append(self.ind() + "if (%s) {" % c)
            self.blocks.append("if1")
            self.exec_inner(u, s, rest, labels)
            self.blocks.pop()
            o.append(self.ind() + "}")
            return
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", t)
        if m:
            self.blocks.pop()
            o.append(self.ind() + "} else if (%s) {" % self.ex(u, parse_expr(m.group(1), s.where)))
            self.blocks.append("if")
            return
        if t == "else":
            self.blocks.pop()
            o.append(self.ind() + "} else {")
            self.blocks.append("if")
            return
        if re.match(r"^end\s*if$", t):
            if not self.blocks or self.blocks[-1] != "if":
                raise RuntimeError("ENDIF without IF")
            self.blocks.pop()
            o.append(self.ind() + "}")
            return
        m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", t)
        if m:
            name = m.group(1)
            args = []
            if m.group(2) and m.group(2).strip():
                ref = parse_expr("x(" + m.group(2) + ")", s.where)
                args = ref.args
            callee = self.p.procs.get(name)
            acts = [self.actual(u, a, None) for a in args]
            if callee is None and name not in self.stubs:
                raise RuntimeError("call to unknown procedure %s" % name)
            if callee is not None:
                if len(callee.args) != len(args):
                    raise RuntimeError("call of %s with %d arguments, it declares %d" % (name, len(args), len(callee.args)))
                for k, (a, f) in enumerate(zip(args, callee.args)):
                    fs = callee.syms[f]
                    if fs.ftype == "character":
                        continue
                    at = self.etype(u, a)
                    ft = {"integer": "int", "double": "double", "logical": "bool", "real": "double"}[fs.ftype]
                    if at != ft:
                        # storage association across types: pass the same address, as Fortran does
                        # (FFTPACK keeps its integer factor table in the tail of the double work array)
                        if not isinstance(a, (Name, Ref)):
                            raise RuntimeError("type mismatch on an expression argument of %s" % name)
                        acts[k] = "(%s *)(void *)(%s)" % (fs.ctype, acts[k])
            o.append(self.ind() + "%s_(%s);" % (name, ", ".join(acts)))
            return
        # assignment
        toks = tokenize(t, s.where)
        p = Parser(toks, s.where)
        lhs = p.p_primary()
        if not p.accept("="):
            raise RuntimeError("statement not understood")
        rhs = p.expr()
        if not p.done():
            raise RuntimeError("trailing tokens after assignment")
        if isinstance(lhs, Name) and u.kind == "function" and lhs.name == u.name:
            o.append(self.ind() + "%s_res = %s;" % (u.name, self.ex(u, rhs)))
            return
        lt = self.etype(u, lhs)
        r = self.ex(u, rhs)
        if lt == "int" and self.etype(u, rhs) == "double":
            r = "(int)(%s)" % r          # Fortran truncates on real -> integer assignment
        o.append(self.ind() + "%s = %s;" % (self.ex(u, lhs), r))

    # ---- C API over the module variables and the argument-less procedures
    def api(self, mods, procs):
        o = self.out
        o.append("")
        o.append('extern "C" {')
        o.append("int ref_set_param(const char *name, double v) { if (f2c_ready) return 2; f2c_override[name] = v; return 0; }")
        o.append("int ref_init(void) { try { f2c_module_init(); } catch (...) { return 1; } return 0; }")
        o.append("// address, element count and type (0 int, 1 double, 2 logical) of a module variable")
        o.append("int ref_var(const char *name, void **addr, long *count, int *type) {")
        o.append("  f2c_module_init();")
        o.append("  const std::string n(name);")
        for mod in mods:
            for n in mod.order:
                sy = mod.syms[n]
                if sy.ftype in ("?", "character"):
                    continue
                ty = {"integer": 0, "double": 1, "logical": 2, "real": 1}[sy.ftype]
                if sy.dims is not None:
                    o.append('  if (n == "%s") { *addr = (void *)%s_; *count = %s; *type = %d; return 0; }' % (n, n, self.total(mod, sy), ty))
                else:
                    o.append('  if (n == "%s") { *addr = (void *)&%s_; *count = 1; *type = %d; return 0; }' % (n, n, ty))
        o.append("  return 1;")
        o.append("}")
        o.append("// call a procedure; arguments (if any) are addresses, as Fortran passes them")
        o.append("int ref_call(const char *name, void **args, int nargs) {")
        o.append("  f2c_module_init();")
        o.append("  const std::string n(name);")
        o.append("  try {")
        for n in procs:
            u = self.p.procs[n]
            if u.kind != "subroutine":
                continue
            if any(u.syms[a].ftype == "character" for a in u.args):
                continue
            cast = ", ".join("(%s *)args[%d]" % (u.syms[a].ctype, k) for k, a in enumerate(u.args))
            o.append('    if (n == "%s") { if (nargs != %d) return 3; %s_(%s); return 0; }' % (n, len(u.args), n, cast))
        o.append("  } catch (const std::exception &e) { f2c_last_error = e.what(); return 2; }")
        o.append("  return 1;")
        o.append("}")
        o.append("const char *ref_last_error(void) { return f2c_last_error.c_str(); }")
        o.append("}")


def main(argv):
    import argparse
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--src", required=True, help="directory of the reference's Fortran sources")
    ap.add_argument("--files", nargs="+", required=True)
    ap.add_argument("--define", "-D", action="append", default=[])
    ap.add_argument("--want", nargs="+", required=True, help="procedures to translate (their callees are added)")
    ap.add_argument("--stub", action="append", default=[], help="name=C++ prototype of a procedure supplied by hand (LAPACK)")
    ap.add_argument("-o", required=True)
    a = ap.parse_args(argv)
    prog = Program()
    for f in a.files:
        prog.add_file(os.path.join(a.src, f), a.define)
    stubs = {}
    for sdef in a.stub:
        n, proto = sdef.split("=", 1)
        stubs[n] = proto
    em = Emitter(prog, a.want, stubs)
    text = em.emit()
    with open(a.o, "w") as fh:
        fh.write(text)
    print("f2cpp: %d procedures, %d lines -> %s" % (len(em.closure()), text.count("\n"), a.o))


if __name__ == "__main__":
    main(sys.argv[1:])

"""f90interp.py -- a small interpreter for the Fortran 90 subset the reference is written in.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  There is no Fortran compiler in the image, so the
reference's four source files (/root/reference/{kd_tree,query_esrg,triangle_walk,interpolation}.f90) cannot be
built.  This module executes their *unmodified source text* statement by statement instead: it knows Fortran
semantics (1-based arrays and sections, integer division, DO-variable values after a loop, pointers to derived
types, argument association by reference, default REAL kind) and nothing about interpolation.  It is used once,
in this container, by tests/golden/make_golden_f90.py to produce golden vectors from the reference itself; the
C++ oracle and the CUDA library are then checked against those vectors.  It is slow (a tree walker in Python)
and only meant for inputs of a few hundred points.

Supported: MODULE / CONTAINS / [RECURSIVE|PURE] SUBROUTINE / FUNCTION ... RESULT, derived TYPEs with default
initialisation, INTEGER / REAL / REAL(8) / LOGICAL / TYPE(x) declarations with DIMENSION / INTENT / POINTER /
ALLOCATABLE, assignment (scalar, whole array, section), pointer assignment, IF / ELSE IF / ELSE, one-line IF,
DO (counted, WHILE, endless), EXIT, CYCLE, RETURN, CALL, ALLOCATE / DEALLOCATE, CONTINUE, array constructors with
implied DO, the intrinsics the four files use.  PRINT / WRITE and OpenMP / f2py directives are ignored (a serial
execution of an OpenMP loop gives the same results: the loops are independent per target).
"""
from __future__ import annotations

import re
import sys

import numpy as np

sys.setrecursionlimit(20000)


class F90Error(Exception):
    pass


# ------------------------------------------------------------------------------------------------- source
def logical_lines(text):
    """Comments stripped, continuation lines joined, lower-cased (string literals are only used in PRINT / WRITE,
    which are ignored, so they are dropped)."""
    out, cur = [], ""
    for raw in text.splitlines():
        line, quote = "", None
        for ch in raw:
            if quote:
                if ch == quote:
                    quote = None
                continue
            if ch in "'\"":
                quote = ch
                line += "''"
                continue
            if ch == "!":
                break
            line += ch
        line = line.strip().lower()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        out.append((cur + line).strip())
        cur = ""
    return out


TOKEN = re.compile(r"""
    (?P<num>\d+\.(?![a-z]+\.)\d*(?:[ed][+-]?\d+)?|\.\d+(?:[ed][+-]?\d+)?|\d+[ed][+-]?\d+|\d+)
  | (?P<dot>\.(?:and|or|not|eqv|neqv|true|false|lt|le|gt|ge|eq|ne)\.)
  | (?P<name>[a-z_][a-z0-9_]*)
  | (?P<str>'')
  | (?P<op>\*\*|==|/=|<=|>=|=>|\(/|/\)|[-+*/<>=(),:%\[\]])
  | (?P<ws>\s+)
""", re.X)


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m:
            raise F90Error(f"cannot tokenize {s[pos:pos + 20]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        toks.append((kind, m.group()))
    toks.append(("end", ""))
    return toks


DOT_REL = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "/="}


class Parser:
    """Recursive-descent parser for expressions and designators -> nested tuples."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i]

    def take(self, val=None):
        tok = self.t[self.i]
        if val is not None and tok[1] != val:
            raise F90Error(f"expected {val!r}, got {tok[1]!r}")
        self.i += 1
        return tok

    def at(self, val):
        return self.t[self.i][1] == val

    # precedence: .eqv. < .or. < .and. < .not. < relational < +,- < *,/ < unary < **
    def expr(self):
        left = self.p_or()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.take()[1]
            left = ("bin", op, left, self.p_or())
        return left

    def p_or(self):
        left = self.p_and()
        while self.at(".or."):
            self.take()
            left = ("bin", ".or.", left, self.p_and())
        return left

    def p_and(self):
        left = self.p_not()
        while self.at(".and."):
            self.take()
            left = ("bin", ".and.", left, self.p_not())
        return left

    def p_not(self):
        if self.at(".not."):
            self.take()
            return ("not", self.p_not())
        return self.p_rel()

    def p_rel(self):
        left = self.p_add()
        op = self.peek()[1]
        op = DOT_REL.get(op, op)
        if op in ("<", "<=", ">", ">=", "==", "/="):
            self.take()
            return ("bin", op, left, self.p_add())
        return left

    def p_add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            left = self.p_mul()
            if op == "-":
                left = ("neg", left)
        else:
            left = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            left = ("bin", op, left, self.p_mul())
        return left

    def p_mul(self):
        left = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            left = ("bin", op, left, self.p_pow())
        return left

    def p_pow(self):
        base = self.p_primary()
        if self.at("**"):
            self.take()
            if self.peek()[1] in ("+", "-"):
                sign = self.take()[1]
                expo = self.p_pow()
                expo = ("neg", expo) if sign == "-" else expo
            else:
                expo = self.p_pow()  # right associative
            return ("bin", "**", base, expo)
        return base

    def p_primary(self):
        kind, val = self.peek()
        if kind == "num":
            self.take()
            if re.fullmatch(r"\d+", val):
                return ("int", int(val))
            return ("real", float(val.replace("d", "e")))
        if val in (".true.", ".false."):
            self.take()
            return ("log", val == ".true.")
        if val == "(":
            self.take()
            e = self.expr()
            self.take(")")
            return e
        if val in ("(/", "["):
            close = "/)" if val == "(/" else "]"
            self.take()
            items = []
            while not self.at(close):
                items.append(self.ac_item())
                if self.at(","):
                    self.take()
            self.take(close)
            return ("acon", items)
        if val in ("+", "-"):
            self.take()
            e = self.p_primary()
            return ("neg", e) if val == "-" else e
        if kind == "name":
            return self.designator()
        raise F90Error(f"unexpected token {val!r}")

    def ac_item(self):
        # implied do: ( expr, var = lo, hi [, step] )
        if self.at("("):
            save = self.i
            self.take()
            e = self.expr()
            if self.at(","):
                self.take()
                if self.peek()[0] == "name" and self.t[self.i + 1][1] == "=":
                    var = self.take()[1]
                    self.take("=")
                    lo = self.expr()
                    self.take(",")
                    hi = self.expr()
                    step = None
                    if self.at(","):
                        self.take()
                        step = self.expr()
                    self.take(")")
                    return ("implied", e, var, lo, hi, step)
            self.i = save
        return self.expr()

    def designator(self):
        parts = []
        while True:
            name = self.take()[1]
            args = None
            if self.at("("):
                self.take()
                args = []
                while not self.at(")"):
                    args.append(self.subscript())
                    if self.at(","):
                        self.take()
                self.take(")")
            parts.append((name, args))
            if self.at("%"):
                self.take()
                continue
            break
        return ("des", parts)

    def subscript(self):
        lo = hi = step = None
        if not self.at(":"):
            lo = self.expr()
            if not self.at(":"):
                return lo
        self.take(":")
        if not (self.at(",") or self.at(")") or self.at(":")):
            hi = self.expr()
        if self.at(":"):
            self.take()
            step = self.expr()
        return ("slice", lo, hi, step)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.peek()[0] != "end":
        raise F90Error(f"trailing tokens in expression {s!r}")
    return e


def split_top(s, sep=","):
    """split on `sep` outside parentheses / brackets"""
    parts, depth, cur = [], 0, ""
    i = 0
    while i < len(s):
        ch = s[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
        i += 1
    if cur.strip():
        parts.append(cur.strip())
    return parts


def matching_paren(s, start):
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise F90Error(f"unbalanced parentheses in {s!r}")


# --------------------------------------------------------------------------------------------- program units
TYPE_RE = re.compile(r"^(integer|real\s*\(\s*8\s*\)|real|logical|type\s*\(\s*([a-z_0-9]+)\s*\))\s*(.*)$")


class Decl:
    __slots__ = ("name", "base", "tname", "dims", "pointer", "allocatable", "init", "intent")

    def __init__(self):
        self.dims = None
        self.pointer = self.allocatable = False
        self.init = None
        self.intent = None


def parse_decl(line):
    """-> list of Decl or None if `line` is not a type declaration"""
    m = TYPE_RE.match(line)
    if not m:
        return None
    base = m.group(1)
    rest = m.group(3)
    if base.startswith("type") and rest.startswith("::") is False and not rest.startswith(","):
        # "type :: name" (a definition header) is handled elsewhere; "type(x) ..." is a declaration
        pass
    tname = m.group(2)
    base = "real8" if base.startswith("real") and "8" in base else ("type" if tname else base)
    attrs, ents = "", rest
    if "::" in rest:
        attrs, ents = rest.split("::", 1)
    dims_attr, pointer, alloc, intent = None, False, False, None
    for a in split_top(attrs.strip().lstrip(",")):
        a = a.strip()
        if a.startswith("dimension"):
            dims_attr = split_top(a[a.index("(") + 1:matching_paren(a, a.index("("))])
        elif a == "pointer":
            pointer = True
        elif a == "allocatable":
            alloc = True
        elif a.startswith("intent"):
            intent = a[a.index("(") + 1:-1].strip()
    out = []
    for ent in split_top(ents):
        d = Decl()
        d.base, d.tname, d.pointer, d.allocatable, d.intent = base, tname, pointer, alloc, intent
        d.dims = dims_attr
        if "=>" in ent:
            ent, init = ent.split("=>", 1)
            d.init = ("null",)
        elif "=" in ent:
            ent, init = ent.split("=", 1)
            d.init = parse_expr(init.strip())
        ent = ent.strip()
        if "(" in ent:
            p0 = ent.index("(")
            d.dims = split_top(ent[p0 + 1:matching_paren(ent, p0)])
            ent = ent[:p0].strip()
        d.name = ent
        out.append(d)
    return out


class Proc:
    def __init__(self, name, kind, args, result):
        self.name, self.kind, self.args, self.result = name, kind, args, result
        self.decls = {}
        self.decl_order = []
        self.body = []


PROC_RE = re.compile(r"^(?:(?:recursive|pure|elemental)\s+)*(subroutine|function)\s+([a-z_0-9]+)\s*(\(.*?\))?\s*"
                     r"(?:result\s*\(\s*([a-z_0-9]+)\s*\))?$")


class Program:
    """All modules of the given source files: derived types + procedures."""

    def __init__(self, texts):
        self.types, self.procs = {}, {}
        for text in texts:
            self._parse_file(logical_lines(text))

    def _parse_file(self, lines):
        i = 0
        while i < len(lines):
            ln = lines[i]
            m = re.match(r"^type\s*(?:::)?\s*([a-z_0-9]+)$", ln)
            if m and not ln.startswith("type("):
                fields = []
                i += 1
                while not lines[i].startswith("end type"):
                    fields.extend(parse_decl(lines[i]))
                    i += 1
                self.types[m.group(1)] = fields
                i += 1
                continue
            pm = PROC_RE.match(ln)
            if pm:
                args = [a.strip() for a in pm.group(3)[1:-1].split(",")] if pm.group(3) and pm.group(3) != "()" else []
                proc = Proc(pm.group(2), pm.group(1), args, pm.group(4) or pm.group(2))
                i += 1
                body_lines = []
                while not re.match(r"^end\s*(subroutine|function)\b", lines[i]):
                    body_lines.append(lines[i])
                    i += 1
                i += 1
                self._parse_proc(proc, body_lines)
                self.procs[proc.name] = proc
                continue
            i += 1  # module / use / implicit / private / public / contains / end module

    def _parse_proc(self, proc, lines):
        k = 0
        while k < len(lines):
            ln = lines[k]
            if re.match(r"^(implicit|use)\b", ln):
                k += 1
                continue
            decls = parse_decl(ln) if not re.match(r"^[a-z_0-9]+\s*(\(.*\))?\s*(%.*)?=", ln) else None
            if decls is None:
                break
            for d in decls:
                proc.decls[d.name] = d
                proc.decl_order.append(d.name)
            k += 1
        proc.body, rest = self._parse_block(lines, k, ())
        if rest != len(lines):
            raise F90Error(f"unparsed lines in {proc.name}: {lines[rest]!r}")

    def _parse_block(self, lines, k, enders):
        """parse statements until a line starting with one of `enders`; returns (stmts, index of the ender)"""
        stmts = []
        while k < len(lines):
            ln = lines[k]
            if any(re.match(e, ln) for e in enders):
                return stmts, k
            st, k = self._parse_stmt(lines, k)
            if st is not None:
                stmts.append(st)
        return stmts, k

    def _parse_stmt(self, lines, k):
        ln = lines[k]
        if re.match(r"^(print\b|write\s*\()", ln) or ln == "continue":
            return None, k + 1
        if ln == "exit":
            return ("exit",), k + 1
        if ln == "cycle":
            return ("cycle",), k + 1
        if ln == "return":
            return ("return",), k + 1
        if ln.startswith("if") and re.match(r"^if\s*\(", ln):
            p0 = ln.index("(")
            p1 = matching_paren(ln, p0)
            cond = parse_expr(ln[p0 + 1:p1])
            tail = ln[p1 + 1:].strip()
            if tail != "then":  # one-line IF
                st, _ = self._parse_stmt([tail], 0)
                return ("if", [(cond, [st] if st else [])], []), k + 1
            branches, else_block = [], []
            block, k = self._parse_block(lines, k + 1, (r"^else\b", r"^end\s*if\b"))
            branches.append((cond, block))
            while True:
                ln = lines[k]
                if re.match(r"^end\s*if\b", ln):
                    return ("if", branches, else_block), k + 1
                m = re.match(r"^else\s*if\s*\(", ln)
                if m:
                    p0 = ln.index("(")
                    p1 = matching_paren(ln, p0)
                    cond = parse_expr(ln[p0 + 1:p1])
                    block, k = self._parse_block(lines, k + 1, (r"^else\b", r"^end\s*if\b"))
                    branches.append((cond, block))
                else:  # else
                    else_block, k = self._parse_block(lines, k + 1, (r"^end\s*if\b",))
        if re.match(r"^do\b", ln):
            head = ln[2:].strip()
            enders = (r"^end\s*do\b",)
            if not head:
                block, k = self._parse_block(lines, k + 1, enders)
                return ("doforever", block), k + 1
            if head.startswith("while"):
                p0 = head.index("(")
                cond = parse_expr(head[p0 + 1:matching_paren(head, p0)])
                block, k = self._parse_block(lines, k + 1, enders)
                return ("dowhile", cond, block), k + 1
            var, ctl = head.split("=", 1)
            parts = split_top(ctl)
            step = parse_expr(parts[2]) if len(parts) > 2 else None
            block, k = self._parse_block(lines, k + 1, enders)
            return ("do", var.strip(), parse_expr(parts[0]), parse_expr(parts[1]), step, block), k + 1
        if ln.startswith("call "):
            rest = ln[5:].strip()
            if "(" in rest:
                p0 = rest.index("(")
                name = rest[:p0].strip()
                args = [parse_expr(a) for a in split_top(rest[p0 + 1:matching_paren(rest, p0)])]
            else:
                name, args = rest, []
            return ("call", name, args), k + 1
        m = re.match(r"^(allocate|deallocate|nullify)\s*\(", ln)
        if m:
            p0 = ln.index("(")
            items = [parse_expr(a) for a in split_top(ln[p0 + 1:matching_paren(ln, p0)])]
            return (m.group(1), items), k + 1
        # assignment / pointer assignment: find the top-level '=' or '=>'
        depth = 0
        for pos, ch in enumerate(ln):
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "=" and depth == 0 and ln[pos - 1] not in "<>/=" and ln[pos + 1:pos + 2] != "=":
                lhs = parse_expr(ln[:pos])
                if ln[pos + 1:pos + 2] == ">":
                    rhs_txt = ln[pos + 2:].strip()
                    rhs = ("null",) if rhs_txt.startswith("null") else parse_expr(rhs_txt)
                    return ("ptrassign", lhs, rhs), k + 1
                return ("assign", lhs, parse_expr(ln[pos + 1:])), k + 1
        raise F90Error(f"cannot parse statement {ln!r}")


# ------------------------------------------------------------------------------------------------- runtime
class FObj:
    """an instance of a derived type"""
    __slots__ = ("tname", "f")

    def __init__(self, tname):
        self.tname, self.f = tname, {}

    def copy(self):
        o = FObj(self.tname)
        for k, v in self.f.items():
            o.f[k] = v.copy() if isinstance(v, np.ndarray) else v
        return o


class Cell:
    """a variable: something that can be read, assigned and passed by reference"""
    __slots__ = ("v", "decl")

    def __init__(self, v=None, decl=None):
        self.v, self.decl = v, decl

    def get(self):
        return self.v

    def set(self, v):
        self.v = v


class FieldCell(Cell):
    __slots__ = ("obj", "name")

    def __init__(self, obj, name, decl=None):
        self.obj, self.name, self.decl = obj, name, decl

    def get(self):
        return self.obj.f[self.name]

    def set(self, v):
        self.obj.f[self.name] = v


class ElemCell(Cell):
    __slots__ = ("arr", "idx")

    def __init__(self, arr, idx, decl=None):
        self.arr, self.idx, self.decl = arr, idx, decl

    def get(self):
        v = self.arr[self.idx]
        return int(v) if isinstance(v, np.integer) else (bool(v) if isinstance(v, np.bool_) else v)

    def set(self, v):
        self.arr[self.idx] = v


class _Exit(Exception):
    pass


class _Cycle(Exception):
    pass


class _Return(Exception):
    pass


class Interpreter:
    def __init__(self, sources, real=np.float64):
        self.prog = Program(sources)
        self.real = np.dtype(real).type
        self.calls = 0

    # ---- values
    def _dtype(self, d):
        return {"integer": np.int64, "real": self.real, "real8": np.float64, "logical": np.bool_}.get(d.base, object)

    def _new_obj(self, tname, frame):
        o = FObj(tname)
        for fd in self.prog.types[tname]:
            o.f[fd.name] = self._default(fd, frame)
        return o

    def _default(self, d, frame):
        if d.pointer or d.allocatable:
            return None
        if d.dims is not None:
            if any(x.strip() == ":" for x in d.dims):
                return None
            shape = tuple(int(self.eval(parse_expr(x), frame)) for x in d.dims)
            if d.base == "type":
                a = np.empty(shape, dtype=object, order="F")
                for idx in np.ndindex(*shape):
                    a[idx] = self._new_obj(d.tname, frame)
                return a
            return np.zeros(shape, dtype=self._dtype(d), order="F")
        if d.base == "type":
            return self._new_obj(d.tname, frame)
        if d.init is not None and d.init != ("null",):
            return self._convert(self.eval(d.init, frame), d)
        return {"integer": 0, "logical": False}.get(d.base, self.real(0.0) if d.base == "real" else np.float64(0.0))

    def _convert(self, v, d):
        if d is None or isinstance(v, (np.ndarray, FObj)) or v is None:
            return v
        if d.base == "integer":
            return int(v)  # truncation toward zero for reals
        if d.base == "real":
            return self.real(v)
        if d.base == "real8":
            return np.float64(v)
        if d.base == "logical":
            return bool(v)
        return v

    # ---- expressions
    def eval(self, e, fr):
        k = e[0]
        if k == "int":
            return e[1]
        if k == "real":
            return self.real(e[1])
        if k == "log":
            return e[1]
        if k == "neg":
            return -self.eval(e[1], fr)
        if k == "not":
            v = self.eval(e[1], fr)
            return ~v if isinstance(v, np.ndarray) else (not v)
        if k == "bin":
            return self._binop(e[1], self.eval(e[2], fr), self.eval(e[3], fr))
        if k == "acon":
            vals = []
            for it in e[1]:
                if it[0] == "implied":
                    lo, hi = int(self.eval(it[3], fr)), int(self.eval(it[4], fr))
                    st = int(self.eval(it[5], fr)) if it[5] is not None else 1
                    saved = fr.get(it[2])
                    fr[it[2]] = Cell(0)
                    for i in range(lo, hi + (1 if st > 0 else -1), st):
                        fr[it[2]].set(i)
                        vals.append(self.eval(it[1], fr))
                    if saved is not None:
                        fr[it[2]] = saved
                else:
                    vals.append(self.eval(it, fr))
            if all(isinstance(v, (int, np.integer)) and not isinstance(v, bool) for v in vals):
                return np.array(vals, dtype=np.int64)
            return np.array([self.real(v) for v in vals], dtype=self.real)
        if k == "des":
            return self._des_value(e, fr)
        if k == "null":
            return None
        raise F90Error(f"cannot evaluate {e!r}")

    def _binop(self, op, a, b):
        if op == ".and.":
            return (a & b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else (a and b)
        if op == ".or.":
            return (a | b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else (a or b)
        if op == ".eqv.":
            return bool(a) == bool(b)
        if op == ".neqv.":
            return bool(a) != bool(b)
        a_int = isinstance(a, (int, np.integer)) and not isinstance(a, bool)
        b_int = isinstance(b, (int, np.integer)) and not isinstance(b, bool)
        if op == "**":
            if b_int:  # integer power = repeated multiplication (x**2 is x*x exactly)
                n = int(b)
                if n < 0:
                    raise F90Error("negative integer exponent not supported")
                res = 1 if a_int else (a * 0 + 1)
                for _ in range(n):
                    res = res * a
                return res
            return self.real(a) ** self.real(b)
        # mixed integer / real: the integer operand is converted to the real kind first
        if a_int and not b_int and not (isinstance(b, np.ndarray) and b.dtype.kind == "i"):
            a = self._to_real_like(a, b)
        if b_int and not a_int and not (isinstance(a, np.ndarray) and a.dtype.kind == "i"):
            b = self._to_real_like(b, a)
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if a_int and b_int:
                q = abs(int(a)) // abs(int(b))  # truncation toward zero
                return q if (a >= 0) == (b >= 0) else -q
            return a / b
        if op == "<":
            return a < b
        if op == "<=":
            return a <= b
        if op == ">":
            return a > b
        if op == ">=":
            return a >= b
        if op == "==":
            return a == b
        if op == "/=":
            return a != b
        raise F90Error(f"operator {op}")

    def _to_real_like(self, i, other):
        if isinstance(other, np.ndarray):
            return other.dtype.type(i) if other.dtype.kind == "f" else i
        if isinstance(other, np.floating):
            return type(other)(i)
        return self.real(i)

    # ---- designators
    def _index(self, arr, args, fr):
        """-> (numpy index tuple, all_scalar)"""
        idx, scalar = [], True
        for a in args:
            if isinstance(a, tuple) and a[0] == "slice":
                lo = self.eval(a[1], fr) if a[1] is not None else 1
                hi = self.eval(a[2], fr) if a[2] is not None else arr.shape[len(idx)]
                st = self.eval(a[3], fr) if a[3] is not None else 1
                idx.append(slice(int(lo) - 1, int(hi), int(st)) if st > 0 else slice(int(lo) - 1, int(hi) - 2 if hi > 1 else None, int(st)))
                scalar = False
            else:
                v = self.eval(a, fr)
                if isinstance(v, np.ndarray):
                    idx.append(v.astype(np.int64) - 1)
                    scalar = False
                else:
                    i = int(v)
                    if not 1 <= i <= arr.shape[len(idx)]:
                        raise F90Error(f"index {i} out of bounds 1..{arr.shape[len(idx)]}")
                    idx.append(i - 1)
        return tuple(idx), scalar

    def _des_cell(self, e, fr):
        """designator -> Cell (reference), or ('view', ndarray) for an array section"""
        parts = e[1]
        name, args = parts[0]
        if name not in fr:
            raise F90Error(f"unknown variable {name!r}")
        cell = fr[name]
        i = 0
        while True:
            if args is not None:
                arr = cell.get()
                if not isinstance(arr, np.ndarray):
                    raise F90Error(f"{name} is not an array")
                idx, scalar = self._index(arr, args, fr)
                if scalar:
                    cell = ElemCell(arr, idx, cell.decl)
                else:
                    view = arr[idx]
                    if i + 1 < len(parts):
                        raise F90Error("component of an array section is not supported")
                    return ("view", view)
            i += 1
            if i == len(parts):
                return cell
            obj = cell.get()
            if not isinstance(obj, FObj):
                raise F90Error(f"component access on a non-object ({parts[i - 1][0]} is {type(obj).__name__})")
            name, args = parts[i]
            fdecl = next(fd for fd in self.prog.types[obj.tname] if fd.name == name)
            cell = FieldCell(obj, name, fdecl)

    def _des_value(self, e, fr):
        parts = e[1]
        name, args = parts[0]
        if name not in fr:
            if len(parts) == 1 and (name in self.prog.procs or name in INTRINSICS):
                return self._call_function(name, args or [], fr)
            raise F90Error(f"unknown name {name!r}")
        c = self._des_cell(e, fr)
        if isinstance(c, tuple):
            return c[1]
        return c.get()

    # ---- calls
    def _call_function(self, name, args, fr):
        if name in self.prog.procs:
            return self._invoke(self.prog.procs[name], args, fr)
        return INTRINSICS[name](self, args, fr)

    def _actual(self, a, fr):
        """actual argument -> Cell (by reference where possible)"""
        if a[0] == "des" and a[1][0][0] in fr:
            c = self._des_cell(a, fr)
            if isinstance(c, tuple):
                return Cell(c[1])
            return c
        return Cell(self.eval(a, fr))

    def _invoke(self, proc, args, fr, cells=None):
        self.calls += 1
        if cells is None:
            cells = [self._actual(a, fr) for a in args]
        if len(cells) != len(proc.args):
            raise F90Error(f"{proc.name}: expected {len(proc.args)} arguments, got {len(cells)}")
        new = {}
        for dummy, cell in zip(proc.args, cells):
            d = proc.decls.get(dummy)
            if d is not None and d.dims is not None and not d.pointer:
                v = cell.get()
                if isinstance(v, np.ndarray):
                    new[dummy] = Cell(v, d)  # same storage (assumed shape: a view with lower bound 1)
                    continue
            if d is not None and cell.decl is None:
                cell.decl = d
            new[dummy] = cell
        for name in proc.decl_order:
            if name in new:
                continue
            d = proc.decls[name]
            new[name] = Cell(self._default(d, new), d)
        try:
            self._run(proc.body, new)
        except _Return:
            pass
        if proc.kind == "function":
            return new[proc.result].get()
        return None

    # ---- statements
    def _assign(self, lhs, val, fr):
        c = self._des_cell(lhs, fr)
        if isinstance(c, tuple):
            view = c[1]
            if view.dtype == object:
                src = val if isinstance(val, np.ndarray) else None
                for n_, idx in enumerate(np.ndindex(*view.shape)):
                    view[idx] = (src[idx] if src is not None else val).copy()
            else:
                view[...] = val.copy() if isinstance(val, np.ndarray) else val
            return
        cur = c.get()
        if isinstance(cur, np.ndarray) and not isinstance(c, ElemCell):
            if cur.dtype == object and isinstance(val, np.ndarray):
                for idx in np.ndindex(*cur.shape):
                    cur[idx] = val[idx].copy()
            else:
                cur[...] = val.copy() if isinstance(val, np.ndarray) else val
            return
        if isinstance(val, FObj):
            val = val.copy()
        c.set(self._convert(val, c.decl))

    def _run(self, stmts, fr):
        for st in stmts:
            k = st[0]
            if k == "assign":
                self._assign(st[1], self.eval(st[2], fr), fr)
            elif k == "if":
                for cond, block in st[1]:
                    if self.eval(cond, fr):
                        self._run(block, fr)
                        break
                else:
                    self._run(st[2], fr)
            elif k == "do":
                var = fr[st[1]]
                lo, hi = int(self.eval(st[2], fr)), int(self.eval(st[3], fr))
                step = int(self.eval(st[4], fr)) if st[4] is not None else 1
                trips = max((hi - lo + step) // step, 0)
                var.set(lo)
                try:
                    for _ in range(trips):
                        try:
                            self._run(st[5], fr)
                        except _Cycle:
                            pass
                        var.set(var.get() + step)
                except _Exit:
                    pass
            elif k == "dowhile":
                try:
                    while self.eval(st[1], fr):
                        try:
                            self._run(st[2], fr)
                        except _Cycle:
                            pass
                except _Exit:
                    pass
            elif k == "doforever":
                try:
                    while True:
                        try:
                            self._run(st[1], fr)
                        except _Cycle:
                            pass
                except _Exit:
                    pass
            elif k == "call":
                if st[1] in self.prog.procs:
                    self._invoke(self.prog.procs[st[1]], st[2], fr)
                elif st[1] not in ("cpu_time", "system_clock"):
                    raise F90Error(f"unknown subroutine {st[1]!r}")
            elif k == "ptrassign":
                c = self._des_cell(st[1], fr)
                c.set(None if st[2] == ("null",) else self.eval(st[2], fr))
            elif k == "allocate":
                for item in st[1]:
                    parts = item[1]
                    name, args = parts[-1]
                    if args is None:  # ALLOCATE(pointer to a derived type)
                        c = self._des_cell(item, fr)
                        c.set(self._new_obj(c.decl.tname, fr))
                    else:
                        c = self._des_cell(("des", parts[:-1] + [(name, None)]), fr)
                        shape = tuple(int(self.eval(a, fr)) for a in args)
                        d = c.decl
                        if d.base == "type":
                            arr = np.empty(shape, dtype=object, order="F")
                            for idx in np.ndindex(*shape):
                                arr[idx] = self._new_obj(d.tname, fr)
                        else:
                            arr = np.zeros(shape, dtype=self._dtype(d), order="F")
                        c.set(arr)
            elif k in ("deallocate", "nullify"):
                for item in st[1]:
                    self._des_cell(item, fr).set(None)
            elif k == "exit":
                raise _Exit()
            elif k == "cycle":
                raise _Cycle()
            elif k == "return":
                raise _Return()
            else:
                raise F90Error(f"statement {k}")

    # ---- Python-facing entry point
    def call(self, name, *pyargs):
        """Call a subroutine / function of the loaded sources.  numpy arrays are passed by reference (give them
        the right dtype: the interpreter's REAL kind, or int64 for INTEGER arrays, Fortran order, bool for LOGICAL)."""
        proc = self.prog.procs[name]
        cells = []
        for dummy, v in zip(proc.args, pyargs):
            d = proc.decls.get(dummy)
            cells.append(v if isinstance(v, Cell) else Cell(self._convert(v, d) if not isinstance(v, np.ndarray) else v, d))
        return self._invoke(proc, None, {}, cells)


# ------------------------------------------------------------------------------------------------ intrinsics
def _vals(itp, args, fr):
    return [itp.eval(a, fr) for a in args]


def _i_size(itp, args, fr):
    v = _vals(itp, args, fr)
    return int(v[0].size) if len(v) == 1 else int(v[0].shape[int(v[1]) - 1])


def _i_real(itp, args, fr):
    v = itp.eval(args[0], fr)
    return v.astype(itp.real) if isinstance(v, np.ndarray) else itp.real(v)


def _i_int(itp, args, fr):
    v = itp.eval(args[0], fr)
    return int(np.trunc(v)) if not isinstance(v, np.ndarray) else np.trunc(v).astype(np.int64)


def _i_floor(itp, args, fr):
    return int(np.floor(itp.eval(args[0], fr)))


def _i_sqrt(itp, args, fr):
    return np.sqrt(itp.eval(args[0], fr))


def _i_abs(itp, args, fr):
    return abs(itp.eval(args[0], fr))


def _i_mod(itp, args, fr):
    a, b = _vals(itp, args, fr)
    r = abs(int(a)) % abs(int(b))  # sign of the dividend (integers only here)
    return r if a >= 0 else -r


def _i_modulo(itp, args, fr):
    a, b = _vals(itp, args, fr)
    return int(a) % int(b)  # sign of the divisor


def _i_min(itp, args, fr):
    v = _vals(itp, args, fr)
    m = v[0]
    for x in v[1:]:
        m = x if x < m else m
    return m


def _i_max(itp, args, fr):
    v = _vals(itp, args, fr)
    m = v[0]
    for x in v[1:]:
        m = x if x > m else m
    return m


def _i_maxval(itp, args, fr):
    v = itp.eval(args[0], fr)
    m = v.max()
    return int(m) if v.dtype.kind == "i" else m


def _i_minval(itp, args, fr):
    v = itp.eval(args[0], fr)
    m = v.min()
    return int(m) if v.dtype.kind == "i" else m


def _i_sum(itp, args, fr):
    v = itp.eval(args[0], fr)
    acc = v.dtype.type(0)
    for x in v.ravel(order="F"):  # sequential, array element order (no pairwise summation)
        acc = acc + x
    return int(acc) if v.dtype.kind == "i" else acc


def _i_count(itp, args, fr):
    return int(np.count_nonzero(itp.eval(args[0], fr)))


def _i_huge(itp, args, fr):
    v = itp.eval(args[0], fr)
    return np.iinfo(np.int32).max if isinstance(v, int) else type(v)(np.finfo(type(v)).max)


def _i_tiny(itp, args, fr):
    v = itp.eval(args[0], fr)
    return type(v)(np.finfo(type(v)).tiny)


def _i_associated(itp, args, fr):
    return itp.eval(args[0], fr) is not None


INTRINSICS = {
    "size": _i_size, "real": _i_real, "int": _i_int, "floor": _i_floor, "sqrt": _i_sqrt, "abs": _i_abs,
    "mod": _i_mod, "modulo": _i_modulo, "min": _i_min, "max": _i_max, "maxval": _i_maxval, "minval": _i_minval,
    "sum": _i_sum, "count": _i_count, "huge": _i_huge, "tiny": _i_tiny, "associated": _i_associated,
    "omp_get_wtime": lambda itp, args, fr: np.float64(0.0),
    "null": lambda itp, args, fr: None,
}

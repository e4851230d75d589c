"""TEST INFRASTRUCTURE ONLY -- a small interpreter for the Fortran-90 subset the reference is written in.

Why it exists: the image has no Fortran compiler, so the reference cannot be built into ``oracle/_ref``.
This module instead EXECUTES THE REFERENCE'S OWN SOURCE TEXT, unmodified, where it lies under
``/root/reference`` (nothing is copied): ``tests/golden/make_ref_golden.py`` loads the ``.f90`` files,
sets the sizes in the parameter modules the way the reference's author does (by changing the values of
the ``integer,parameter`` constants -- here through ``Interp.override``, without touching the files),
calls the hot-path subroutines (``compute_update_exact``, ``compute_update``, ``apply_limiter``,
``evolve``-equivalent loops ...) on seeded inputs and stores inputs and outputs as golden vectors.
The C restatement in ``oracle/*.c`` (the oracle the CUDA path is checked against) is then pinned
against those vectors in ``tests/test_reference_pins.py``.

Semantics implemented (what gfortran does on x86-64 without FMA contraction, ``-O0`` .. ``-O2``):
  * default ``real`` and un-suffixed real literals are IEEE binary32, ``real(kind=8)`` / ``1d0`` binary64,
    mixed-mode arithmetic converts to the wider operand first (so ``gamma=1.4`` stores (double)1.4f);
  * ``integer/integer`` truncates; ``x**2`` is ``x*x``; other ``real**integer`` follow libgcc's
    ``__powidf2`` (binary method, ``1/y`` for negative exponents); ``real**real`` is libm ``pow``;
  * ``exp``/``pow``/``sin``... go to the C library (``math`` module = glibc), never to numpy's SIMD
    approximations; ``sqrt`` and ``/`` are correctly rounded; ``sum`` adds in array-element order;
  * arguments are passed by reference with sequence association (an element or a contiguous section
    passed to an explicit-shape dummy aliases the storage that follows it; a non-contiguous section
    is copied in and out); functions may modify their actual arguments (``2d/legendre.f90`` clamps its
    ``x`` in place);
  * a ``do`` variable keeps ``last+step`` after the loop; locals are re-created on every call
    (filled with NaN / a sentinel so that a use of an undefined value is visible);
  * ``write``/``print``/``open``/``close``/``pause`` are no-ops (the only statements skipped);
  * array subscripts are bounds checked: an out-of-range element raises ``FortranBoundsError`` unless
    ``oob='nan'`` (reads give NaN, writes are dropped, both are counted in ``Interp.oob_count``).

Nothing under ``fvm-source-wb_b200/`` imports this file.
"""
import ctypes
import ctypes.util
import math
import re
import sys

import numpy as np

__all__ = ["Interp", "FortranError", "FortranBoundsError", "FortranStop"]

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("expf", "logf", "sinf", "cosf", "atanf", "acosf", "tanf", "asinf", "log10f", "tanhf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]

F32, F64, I64 = np.float32, np.float64, np.int64


class FortranError(Exception):
    pass


class FortranBoundsError(FortranError):
    pass


class FortranStop(Exception):
    pass


class _Exit(Exception):
    pass


class _Cycle(Exception):
    pass


class _Return(Exception):
    pass


# ------------------------------------------------------------------------------------------------ lexer
_TOK = re.compile(r"""
  (?P<ws>\s+)
 |(?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
 |(?P<num>(?:\d+\.(?![a-z]+\.)\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)
 |(?P<dot>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
 |(?P<name>[a-z_]\w*)
 |(?P<op>\*\*|//|==|/=|<=|>=|=>|::|\(/|/\)|[-+*/(),=<>:%\[\]])
""", re.X | re.I)

_DOTREL = {".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


def _strip_comment(line):
    q = None
    for k, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:k]
    return line


def _logical_lines(text):
    """(first physical line number, joined text) for every statement; '&' continuations joined, ';' split."""
    out, cur, cur_no = [], "", None
    for no, raw in enumerate(text.split("\n"), 1):
        s = _strip_comment(raw).strip()
        if not s:
            continue
        if cur:
            if s.startswith("&"):
                s = s[1:]
        else:
            cur_no = no
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        cur += s
        for part in _split_semicolons(cur):
            if part.strip():
                out.append((cur_no, part.strip()))
        cur = ""
    if cur:
        out.append((cur_no, cur))
    return out


def _split_semicolons(s):
    if ";" not in s:
        return [s]
    parts, q, start = [], None, 0
    for k, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == ";":
            parts.append(s[start:k])
            start = k + 1
    parts.append(s[start:])
    return parts


def _tokenize(s, where):
    toks, pos = [], 0
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m:
            raise FortranError(f"{where}: cannot tokenize {s[pos:pos + 20]!r}")
        pos = m.end()
        k = m.lastgroup
        v = m.group(k)
        if k == "ws":
            continue
        if k == "str":
            q = v[0]
            toks.append(("str", v[1:-1].replace(q + q, q)))
        elif k == "num":
            toks.append(("num", v.lower()))
        elif k == "dot":
            v = v.lower()
            if v in (".true.", ".false."):
                toks.append(("log", v == ".true."))
            else:
                toks.append(("op", _DOTREL.get(v, v)))
        elif k == "name":
            toks.append(("name", v.lower()))
        else:
            toks.append(("op", v))
    return toks


def _number(v):
    kind = None
    if "_" in v:
        v, kind = v.split("_", 1)
    if re.fullmatch(r"\d+", v):
        return int(v)
    if "d" in v:
        return F64(float(v.replace("d", "e")))
    if kind in ("8", "dp"):
        return F64(float(v))
    return F32(float(v))     # decimal -> binary32, correctly rounded (what gfortran's front end does)


# ------------------------------------------------------------------------------------------------ parser
class _P:
    """Recursive-descent expression parser over one statement's tokens."""

    def __init__(self, toks, where):
        self.t, self.i, self.where = toks, 0, where

    def peek(self, k=0):
        j = self.i + k
        return self.t[j] if j < len(self.t) else ("eof", None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def at_op(self, v):
        return self.peek() == ("op", v)

    def at_name(self, v):
        return self.peek() == ("name", v)

    def accept(self, v):
        if self.at_op(v):
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise FortranError(f"{self.where}: expected {v!r}, got {self.peek()!r}")

    def done(self):
        return self.i >= len(self.t)

    # precedence (low -> high): .eqv. | .or. | .and. | .not. | relational | // | + - | * / | ** | primary
    def expr(self):
        return self.p_or()

    def p_or(self):
        l = self.p_and()
        while self.at_op(".or."):
            self.next()
            l = ("or", l, self.p_and())
        return l

    def p_and(self):
        l = self.p_not()
        while self.at_op(".and."):
            self.next()
            l = ("and", l, self.p_not())
        return l

    def p_not(self):
        if self.at_op(".not."):
            self.next()
            return ("not", self.p_not())
        return self.p_rel()

    def p_rel(self):
        l = self.p_cat()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("==", "/=", "<", "<=", ">", ">="):
            self.next()
            return ("rel", tok[1], l, self.p_cat())
        return l

    def p_cat(self):
        l = self.p_add()
        while self.at_op("//"):
            self.next()
            l = ("cat", l, self.p_add())
        return l

    def p_add(self):
        if self.at_op("-"):
            self.next()
            l = ("neg", self.p_mul())
        elif self.at_op("+"):
            self.next()
            l = self.p_mul()
        else:
            l = self.p_mul()
        while self.at_op("+") or self.at_op("-"):
            op = self.next()[1]
            l = ("bin", op, l, self.p_mul())
        return l

    def p_mul(self):
        l = self.p_pow()
        while self.at_op("*") or self.at_op("/"):
            op = self.next()[1]
            l = ("bin", op, l, self.p_pow())
        return l

    def p_pow(self):
        base = self.p_primary()
        if self.at_op("**"):
            self.next()
            # right associative; a unary minus may follow ** directly (gfortran extension)
            if self.at_op("-"):
                self.next()
                ex = ("neg", self.p_pow())
            else:
                ex = self.p_pow()
            return ("pow", base, ex)
        return base

    def p_primary(self):
        k, v = self.next()
        if k == "num":
            return ("const", _number(v))
        if k == "str":
            return ("const", v)
        if k == "log":
            return ("const", bool(v))
        if k == "op" and v == "(":
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if k == "op" and v in ("(/", "["):
            close = "/)" if v == "(/" else "]"
            items = []
            while not self.at_op(close):
                items.append(self.expr())
                if not self.accept(","):
                    break
            self.expect(close)
            return ("acons", items)
        if k == "op" and v == "-":          # e.g. a*-b (extension)
            return ("neg", self.p_primary())
        if k == "name":
            if self.at_op("("):
                self.next()
                return ("ref", v, self.subscripts())
            return ("name", v)
        raise FortranError(f"{self.where}: unexpected token {(k, v)!r}")

    def subscripts(self):
        """After '(' : list of expr | ('slice', lo, hi, step) | ('kw', name, expr); consumes ')'."""
        subs = []
        if self.accept(")"):
            return subs
        while True:
            if self.peek()[0] == "name" and self.peek(1) == ("op", "=") :
                nm = self.next()[1]
                self.next()
                subs.append(("kw", nm, self.expr()))
            else:
                lo = hi = st = None
                if not self.at_op(":"):
                    lo = self.expr()
                if self.accept(":"):
                    if not (self.at_op(",") or self.at_op(")") or self.at_op(":")):
                        hi = self.expr()
                    if self.accept(":"):
                        st = self.expr()
                    subs.append(("slice", lo, hi, st))
                else:
                    subs.append(lo)
            if self.accept(","):
                continue
            self.expect(")")
            return subs


_TYPEKW = ("real", "integer", "logical", "character", "double")
_NOPS = ("write", "print", "open", "close", "pause", "read", "format", "flush", "rewind")


class _Unit:
    def __init__(self, kind, name, args, result, file, line):
        self.kind, self.name, self.args, self.result = kind, name, args, result
        self.file, self.line = file, line
        self.uses, self.decls, self.body = [], [], []
        self.static = {}          # saved (initialised) locals
        self.plan = None          # cached declaration plan


def _parse_typespec(p):
    """real | real(kind=8) | real(8) | real*8 | double precision | integer | logical | character(len=..)"""
    base = p.next()[1]
    kind, clen = None, None
    if base == "double":
        p.next()
        return "real", 8, None
    if p.accept("*"):
        kind = int(p.next()[1])
    elif base == "character" and p.at_op("("):
        depth, j = 0, p.i
        while True:
            if p.t[j] == ("op", "("):
                depth += 1
            elif p.t[j] == ("op", ")"):
                depth -= 1
                if depth == 0:
                    break
            j += 1
        inner = p.t[p.i + 1:j]
        p.i = j + 1
        if ("op", "*") not in inner and ("op", ":") not in inner:
            if len(inner) > 1 and inner[1] == ("op", "="):
                inner = inner[2:]
            clen = _P(inner, p.where).expr()
    elif p.at_op("("):
        p.next()
        for s in p.subscripts():
            if isinstance(s, tuple) and s[0] == "kw":
                if s[1] == "kind":
                    kind = s[2]
                elif s[1] == "len":
                    clen = s[2]
            elif isinstance(s, tuple) and s[0] == "slice":
                clen = None
            elif base == "character":
                clen = s
            else:
                kind = s
    if isinstance(kind, tuple):
        if kind[0] != "const":
            raise FortranError(f"{p.where}: non-literal kind")
        kind = int(kind[1])
    if base == "real":
        kind = kind or 4
    return base, kind, clen


def _parse_decl(p):
    base, kind, clen = _parse_typespec(p)
    attrs = {}
    while p.accept(","):
        a = p.next()[1]
        if a == "dimension":
            p.expect("(")
            attrs["dimension"] = p.subscripts()
        elif a == "intent":
            p.expect("(")
            p.subscripts()
        else:
            attrs[a] = True
    p.accept("::")
    ents = []
    while True:
        nm = p.next()[1]
        dims = None
        if p.at_op("("):
            p.next()
            dims = p.subscripts()
        if p.accept("*"):            # character name*len
            p.next()
        init = None
        if p.accept("="):
            init = p.expr()
        ents.append((nm, dims, init))
        if not p.accept(","):
            break
    return ("decl", base, kind, clen, attrs, ents)


class _Parser:
    def __init__(self, text, fname):
        self.lines = _logical_lines(text)
        self.fname = fname
        self.k = 0

    def where(self, no):
        return f"{self.fname}:{no}"

    def units(self):
        out = []
        while self.k < len(self.lines):
            no, s = self.lines[self.k]
            toks = _tokenize(s, self.where(no))
            self.k += 1
            head = [v for _, v in toks]
            if head[0] in ("module", "program"):
                u = _Unit(head[0], head[1], [], None, self.fname, no)
            elif "function" in head and toks[head.index("function")][0] == "name" and head[0] != "end":
                f = head.index("function")
                p = _P(toks[f + 1:], self.where(no))
                name = p.next()[1]
                args = []
                if p.accept("("):
                    args = [a[1] for a in p.subscripts()]
                res = name
                if p.at_name("result"):
                    p.next()
                    p.expect("(")
                    res = p.subscripts()[0][1]
                u = _Unit("function", name, args, res, self.fname, no)
                if f > 0:       # typed header: real(kind=8) function f(...)
                    tp = _P(toks[:f], self.where(no))
                    base, kind, clen = _parse_typespec(tp)
                    u.decls.append(("decl", base, kind, clen, {}, [(res, None, None)]))
            elif head[0] == "subroutine":
                p = _P(toks[1:], self.where(no))
                name = p.next()[1]
                args = []
                if p.accept("("):
                    args = [a[1] for a in p.subscripts()]
                u = _Unit("subroutine", name, args, None, self.fname, no)
            else:
                raise FortranError(f"{self.where(no)}: expected a program unit, got {s!r}")
            u.body = self.block(u, ("end",))
            self.k += 1     # the end line
            out.append(u)
        return out

    def block(self, u, terms):
        """Parse statements until a line whose first word(s) is in terms; leaves self.k AT that line."""
        body = []
        while True:
            if self.k >= len(self.lines):
                raise FortranError(f"{self.fname}: unexpected end of file in {u.name}")
            no, s = self.lines[self.k]
            toks = _tokenize(s, self.where(no))
            w0 = toks[0][1] if toks[0][0] == "name" else None
            w1 = toks[1][1] if len(toks) > 1 and toks[1][0] == "name" else None
            key = w0
            if w0 == "end" and w1 in ("do", "if", "select"):
                key = "end" + w1
            elif w0 == "else" and w1 == "if":
                key = "elseif"
            if w0 == "end" and w1 in (None, "subroutine", "function", "program", "module"):
                key = "end"
            if key in terms and not self._is_assignment(toks):
                return body
            self.k += 1
            st = self.statement(u, toks, no, key)
            if st is not None:
                body.append(st)

    @staticmethod
    def _is_assignment(toks):
        """name [ (...) ] = ... at depth 0 (so that variables called 'end', 'case' ... would still work)"""
        if toks[0][0] != "name":
            return False
        j, depth = 1, 0
        if j < len(toks) and toks[j] == ("op", "("):
            depth = 1
            j += 1
            while j < len(toks) and depth:
                if toks[j] == ("op", "("):
                    depth += 1
                elif toks[j] == ("op", ")"):
                    depth -= 1
                j += 1
        return j < len(toks) and toks[j] == ("op", "=")

    def statement(self, u, toks, no, key):
        where = self.where(no)
        w0 = toks[0][1] if toks[0][0] == "name" else None
        if self._is_assignment(toks) and not (w0 == "do" and self._looks_like_do(toks)):
            p = _P(toks, where)
            lhs = p.p_primary()
            p.expect("=")
            rhs = p.expr()
            if not p.done():
                raise FortranError(f"{where}: trailing tokens in assignment")
            return ("assign", lhs, rhs, no)
        if w0 in _TYPEKW and not (len(toks) > 1 and toks[1] == ("op", "=")):
            u.decls.append(_parse_decl(_P(toks, where)))
            return None
        if w0 == "use":
            u.uses.append(toks[1][1])
            return None
        if w0 in ("implicit", "external", "contains", "continue", "intrinsic"):
            return None
        if w0 in _NOPS:
            return None
        if w0 == "call":
            p = _P(toks[1:], where)
            name = p.next()[1]
            args = []
            if p.accept("("):
                args = p.subscripts()
            return ("call", name, args, no)
        if key in ("return", "exit", "cycle", "stop"):
            return (key, no)
        if w0 == "do":
            if len(toks) == 1:
                body = self.block(u, ("enddo",))
                self.k += 1
                return ("dowhile", ("const", True), body, no)
            if toks[1] == ("name", "while"):
                p = _P(toks[2:], where)
                p.expect("(")
                cond = p.expr()
                p.expect(")")
                body = self.block(u, ("enddo",))
                self.k += 1
                return ("dowhile", cond, body, no)
            p = _P(toks[1:], where)
            var = p.next()[1]
            p.expect("=")
            a = p.expr()
            p.expect(",")
            b = p.expr()
            c = p.expr() if p.accept(",") else None
            body = self.block(u, ("enddo",))
            self.k += 1
            return ("do", var, a, b, c, body, no)
        if w0 == "if":
            p = _P(toks[1:], where)
            p.expect("(")
            cond = p.expr()
            p.expect(")")
            if p.at_name("then") and p.i == len(p.t) - 1:
                arms, other = [], None
                while True:
                    body = self.block(u, ("elseif", "else", "endif"))
                    arms.append((cond, body))
                    no2, s2 = self.lines[self.k]
                    t2 = _tokenize(s2, self.where(no2))
                    self.k += 1
                    if t2[0][1] == "endif" or (t2[0][1] == "end"):
                        break
                    if t2[0][1] == "elseif" or (t2[0][1] == "else" and len(t2) > 1 and t2[1] == ("name", "if")):
                        off = 1 if t2[0][1] == "elseif" else 2
                        p2 = _P(t2[off:], self.where(no2))
                        p2.expect("(")
                        cond = p2.expr()
                        p2.expect(")")
                        continue
                    other = self.block(u, ("endif",))
                    self.k += 1
                    break
                return ("if", arms, other, no)
            rest = toks[1 + p.i:]
            st = self.statement(u, rest, no, rest[0][1] if rest[0][0] == "name" else None)
            return ("if", [(cond, [st] if st is not None else [])], None, no)
        if w0 == "select":
            p = _P(toks[2:], where)
            p.expect("(")
            sel = p.expr()
            p.expect(")")
            cases, default = [], None
            # skip to the first 'case'
            while True:
                no2, s2 = self.lines[self.k]
                t2 = _tokenize(s2, self.where(no2))
                self.k += 1
                if t2[0][1] == "endselect" or (t2[0][1] == "end" and len(t2) > 1 and t2[1][1] == "select"):
                    break
                if t2[0][1] != "case":
                    raise FortranError(f"{self.where(no2)}: expected case, got {s2!r}")
                if len(t2) > 1 and t2[1] == ("name", "default"):
                    default = self.block(u, ("case", "endselect"))
                else:
                    p2 = _P(t2[1:], self.where(no2))
                    p2.expect("(")
                    vals = p2.subscripts()
                    cases.append((vals, self.block(u, ("case", "endselect"))))
            return ("select", sel, cases, default, no)
        if w0 in ("enddo", "endif", "endselect", "end", "else", "elseif", "case"):
            raise FortranError(f"{where}: unbalanced {w0!r}")
        raise FortranError(f"{where}: cannot parse statement {' '.join(str(v) for _, v in toks)!r}")

    @staticmethod
    def _looks_like_do(toks):
        # do i = a , b   has a top-level comma after '='
        depth = 0
        for k, v in toks[3:]:
            if (k, v) == ("op", "("):
                depth += 1
            elif (k, v) == ("op", ")"):
                depth -= 1
            elif (k, v) == ("op", ",") and depth == 0:
                return True
        return False


# ------------------------------------------------------------------------------------------------ values
class Var:
    """A Fortran variable: storage (numpy array, 0-d for scalars, Fortran order) + lower bounds."""
    __slots__ = ("a", "lb", "clen")

    def __init__(self, a, lb=None, clen=None):
        self.a = a
        self.lb = lb if lb is not None else (1,) * a.ndim
        self.clen = clen


_TCLS = {np.float64: "r8", float: "r8", np.float32: "r4", int: "i", np.int64: "i", np.int32: "i", bool: "l", np.bool_: "l",
         str: "c", np.str_: "c"}
_ADT = {"f8": "r8", "f4": "r4", "i8": "i", "i4": "i", "b1": "l"}


def _tcls(x):
    t = _TCLS.get(type(x))
    if t is not None:
        return t
    if isinstance(x, np.ndarray):
        return _ADT.get(x.dtype.str[1:], "c")
    if isinstance(x, np.integer):
        return "i"
    if isinstance(x, np.floating):
        return "r8"
    raise FortranError(f"unknown value type {type(x)}")


def _conv(x, t):
    if t == "r8":
        return x.astype(F64) if isinstance(x, np.ndarray) else F64(x)
    if t == "r4":
        return x.astype(F32) if isinstance(x, np.ndarray) else F32(x)
    return x


def _promote(a, b):
    if type(a) is np.float64 and type(b) is np.float64:
        return a, b, "r8"
    ta, tb = _tcls(a), _tcls(b)
    if ta == tb:
        return a, b, ta
    if "c" in (ta, tb) or "l" in (ta, tb):
        raise FortranError(f"bad operand types {ta},{tb}")
    t = "r8" if "r8" in (ta, tb) else "r4"
    return (a if ta == t else _conv(a, t)), (b if tb == t else _conv(b, t)), t


def _idiv(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        q = np.abs(a) // np.abs(b)
        return (q * np.sign(a) * np.sign(b)).astype(I64)
    a, b = int(a), int(b)
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _powi(x, m):
    """libgcc __powidf2 / __powisf2 (what gfortran calls for real**integer, except **2 -> x*x)."""
    m = int(m)
    n = -m if m < 0 else m
    one = x.dtype.type(1) if isinstance(x, np.ndarray) else type(x)(1)
    y = x if n % 2 else one
    n >>= 1
    while n:
        x = x * x
        if n % 2:
            y = y * x
        n >>= 1
    return one / y if m < 0 else y


def _map1(f64, f32name):
    f32 = getattr(_libm, f32name)

    def g(x):
        t = _tcls(x)
        if t == "i":
            raise FortranError("integer argument to a real intrinsic")
        if isinstance(x, np.ndarray):
            flat = x.ravel(order="K")
            if t == "r8":
                out = np.array([f64(float(v)) for v in flat], dtype=F64)
            else:
                out = np.array([f32(float(v)) for v in flat], dtype=F32)
            return out.reshape(x.shape, order="F" if x.flags.f_contiguous else "C")
        return F64(f64(float(x))) if t == "r8" else F32(f32(float(x)))
    return g


def _c_exp(v):
    try:
        return math.exp(v)
    except OverflowError:
        return math.inf


def _c_log(v):
    if v > 0:
        return math.log(v)
    return -math.inf if v == 0 else math.nan


def _c_acos(v):
    return math.acos(v) if -1 <= v <= 1 else math.nan


def _c_pow(a, b):
    try:
        return math.pow(a, b)
    except (ValueError, ZeroDivisionError):
        return math.nan if a < 0 else math.inf
    except OverflowError:
        return math.inf


def _pow_real(a, b):
    a, b, t = _promote(a, b)
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.broadcast_arrays(a, b)
        f = _c_pow if t == "r8" else (lambda p, q: _libm.powf(p, q))
        out = np.array([f(float(p), float(q)) for p, q in zip(a.ravel(), b.ravel())],
                       dtype=F64 if t == "r8" else F32)
        return np.asfortranarray(out.reshape(a.shape))
    if t == "r8":
        return F64(_c_pow(float(a), float(b)))
    return F32(_libm.powf(float(a), float(b)))


def _minmax(args, is_max):
    """MAX / MIN as gfortran expands them on x86-64 (trans-intrinsic.c, gfc_conv_intrinsic_minmax):
    mvar = a1;  if (a2 .op. mvar || isnan(mvar)) mvar = a2;  ...   -- a NaN is dropped whichever side it comes from."""
    r = args[0]
    for b in args[1:]:
        r, b, _ = _promote(r, b)
        if isinstance(r, np.ndarray) or isinstance(b, np.ndarray):
            nan_r = np.isnan(r) if np.asarray(r).dtype.kind == "f" else False
            r = np.where((b > r) | nan_r, b, r) if is_max else np.where((b < r) | nan_r, b, r)
        else:
            nan_r = (r != r)
            if is_max:
                r = b if (b > r or nan_r) else r
            else:
                r = b if (b < r or nan_r) else r
    return r


def _seq_sum(x):
    flat = x.ravel(order="F")
    s = flat.dtype.type(0)
    for v in flat:
        s = s + v
    return s


def _sign(a, b):
    a, b, _ = _promote(a, b)
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.where(np.signbit(b) if _tcls(b) != "i" else b < 0, -np.abs(a), np.abs(a))
    neg = (b < 0) if _tcls(b) == "i" else bool(np.signbit(b))
    return -abs(a) if neg else abs(a)


def _mod(a, b):
    a, b, t = _promote(a, b)
    if t == "i":
        return a - _idiv(a, b) * b
    return np.fmod(a, b)


def _modulo(a, b):
    a, b, t = _promote(a, b)
    if t == "i":
        return int(a) % int(b)
    return a - np.floor(a / b) * b


def _matmul(a, b):
    a, b, t = _promote(a, b)
    dt = {"r8": F64, "r4": F32, "i": I64}[t]
    if a.ndim == 2 and b.ndim == 2:
        out = np.zeros((a.shape[0], b.shape[1]), dtype=dt, order="F")
        for i in range(a.shape[0]):
            for j in range(b.shape[1]):
                s = dt(0)
                for k in range(a.shape[1]):
                    s = s + a[i, k] * b[k, j]
                out[i, j] = s
        return out
    if a.ndim == 2 and b.ndim == 1:
        out = np.zeros(a.shape[0], dtype=dt)
        for i in range(a.shape[0]):
            s = dt(0)
            for k in range(a.shape[1]):
                s = s + a[i, k] * b[k]
            out[i] = s
        return out
    if a.ndim == 1 and b.ndim == 2:
        out = np.zeros(b.shape[1], dtype=dt)
        for j in range(b.shape[1]):
            s = dt(0)
            for k in range(a.shape[0]):
                s = s + a[k] * b[k, j]
            out[j] = s
        return out
    raise FortranError("matmul rank")


_INTRINSICS = {
    "exp": _map1(_c_exp, "expf"), "log": _map1(_c_log, "logf"),
    "sin": _map1(math.sin, "sinf"), "cos": _map1(math.cos, "cosf"), "tan": _map1(math.tan, "tanf"),
    "atan": _map1(math.atan, "atanf"), "acos": _map1(_c_acos, "acosf"), "asin": _map1(math.asin, "asinf"),
    "tanh": _map1(math.tanh, "tanhf"), "log10": _map1(math.log10, "log10f"),
    "dexp": _map1(_c_exp, "expf"), "dsqrt": np.sqrt,
    "sqrt": lambda x: np.sqrt(x),
    "abs": lambda x: abs(x) if not isinstance(x, np.ndarray) else np.abs(x),
    "dabs": lambda x: abs(x),
    "max": lambda *a: _minmax(a, True), "min": lambda *a: _minmax(a, False),
    "dmax1": lambda *a: _minmax(a, True), "dmin1": lambda *a: _minmax(a, False),
    "dble": lambda x: _conv(x, "r8"),
    "int": lambda x, kind=None: (np.trunc(x).astype(I64) if isinstance(x, np.ndarray) else int(x)),
    "nint": lambda x: (np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5)).astype(I64)
                       if isinstance(x, np.ndarray) else int(math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5))),
    "floor": lambda x: (np.floor(x).astype(I64) if isinstance(x, np.ndarray) else int(math.floor(x))),
    "sign": _sign, "mod": _mod, "modulo": _modulo,
    "sum": _seq_sum,
    "trim": lambda s: s.rstrip(), "adjustl": lambda s: s.lstrip().ljust(len(s)), "len": lambda s: len(s),
    "len_trim": lambda s: len(s.rstrip()),
    "transpose": lambda a: np.asfortranarray(a.T.copy()),
    "matmul": _matmul,
    "epsilon": lambda x: np.finfo(F64 if _tcls(x) == "r8" else F32).eps.astype(F64 if _tcls(x) == "r8" else F32),
    "huge": lambda x: (np.finfo(F64).max if _tcls(x) == "r8" else np.finfo(F32).max if _tcls(x) == "r4" else 2147483647),
    "tiny": lambda x: (np.finfo(F64).tiny if _tcls(x) == "r8" else np.finfo(F32).tiny),
    "isnan": lambda x: np.isnan(x),
}


def _real(x, kind=None):
    if kind is None:
        kind = 4
    return _conv(x, "r8" if int(kind) == 8 else "r4")


_INTRINSICS["real"] = _real


class _Actual:
    """An actual argument before it is associated with a dummy."""
    __slots__ = ("kind", "var", "idx", "value")

    def __init__(self, kind, var=None, idx=None, value=None):
        self.kind, self.var, self.idx, self.value = kind, var, idx, value


_DTYPES = {("real", 4): F32, ("real", 8): F64, ("integer", None): I64, ("integer", 4): I64, ("integer", 8): I64,
           ("logical", None): np.bool_, ("character", None): object}
_INT_SENTINEL = -777777777


class Interp:
    """Loads reference .f90 files and calls their procedures on numpy arrays.

    override(module, name=value, ...) replaces the initial value of module entities (the reference is configured
    by editing its parameter modules; this is the same act without touching the files)."""

    def __init__(self, oob="raise", undefined="nan", trace=None):
        self.units = {}           # procedures and programs by name
        self.modules = {}         # name -> _Unit
        self.modvars = {}         # name -> {entity: Var} once elaborated
        self.overrides = {}
        self.oob = oob
        self.undefined = undefined
        self.oob_count = 0
        self.calls = {}           # procedure name -> number of calls (coverage evidence for the goldens)
        self.hooks = {}           # procedure name -> python callable(interp, frame) replacing the body (unused by default)
        self.files = []
        self.probes = {}          # line number -> callable(dict of the running unit's variables): test harness taps
        self.depth = 0
        self.trace = trace

    # ---- loading
    def load(self, path):
        with open(path, "r", errors="replace") as f:
            text = f.read()
        for u in _Parser(text, path).units():
            if u.kind == "module":
                self.modules[u.name] = u
            else:
                self.units[u.name] = u
        self.files.append(path)
        return self

    def override(self, module, **kv):
        self.overrides.setdefault(module, {}).update(kv)
        self.modvars.pop(module, None)
        for u in self.units.values():
            u.static.clear()
        return self

    def module(self, name):
        """Elaborated variables of a module (dict name -> Var)."""
        if name not in self.modvars:
            u = self.modules[name]
            fr = {}
            self.modvars[name] = fr
            for m in u.uses:
                fr.update(self.module(m))
            ov = self.overrides.get(name, {})
            for d in u.decls:
                _, base, kind, clen, attrs, ents = d
                for nm, dims, init in ents:
                    dims = dims if dims is not None else attrs.get("dimension")
                    v = self._alloc(fr, base, kind, clen, dims, f"{u.file}:{nm}")
                    fr[nm] = v
                    if nm in ov:
                        self._store_whole(v, ov[nm])
                    elif init is not None:
                        self._store_whole(v, self.ev(init, fr))
            unknown = set(ov) - set(fr)
            if unknown:
                raise FortranError(f"override of unknown entities {unknown} in module {name}")
        return self.modvars[name]

    def get(self, module, name):
        v = self.module(module)[name]
        return v.a[()] if v.a.ndim == 0 else v.a

    # ---- storage
    def _alloc(self, fr, base, kind, clen, dims, where, fill=True):
        dt = _DTYPES.get((base, kind)) or _DTYPES.get((base, None))
        if dt is None:
            raise FortranError(f"{where}: unsupported type {base}({kind})")
        cl = None
        if base == "character":
            cl = None if clen is None else int(self.ev(clen, fr))
        if dims is None:
            a = np.zeros((), dtype=dt)
            lb = ()
        else:
            lbs, shape = [], []
            for d in dims:
                if isinstance(d, tuple) and d[0] == "slice":
                    lo = 1 if d[1] is None else int(self.ev(d[1], fr))
                    hi = int(self.ev(d[2], fr))
                else:
                    lo, hi = 1, int(self.ev(d, fr))
                lbs.append(lo)
                shape.append(max(hi - lo + 1, 0))
            a = np.zeros(tuple(shape), dtype=dt, order="F")
            lb = tuple(lbs)
        if fill:
            if dt in (F32, F64):
                a[...] = np.nan if self.undefined == "nan" else 0.0
            elif dt is I64:
                a[...] = _INT_SENTINEL if self.undefined == "nan" else 0
            elif dt is object:
                a[...] = ""
        return Var(a, lb, cl)

    @staticmethod
    def _store_whole(v, val):
        if v.a.dtype == object:
            s = str(val)
            if v.clen is not None:
                s = s[:v.clen].ljust(v.clen)
            v.a[...] = s
        else:
            v.a[...] = val

    # ---- expression evaluation
    def ev(self, e, fr):
        k = e[0]
        if k == "const":
            return e[1]
        if k == "name":
            v = fr.get(e[1])
            if v is None:
                if e[1] in self.units and self.units[e[1]].kind == "function":
                    return self.call_function(e[1], [], fr)
                raise FortranError(f"undefined name {e[1]!r}")
            return v.a[()] if v.a.ndim == 0 else v.a
        if k == "paren":
            return self.ev(e[1], fr)
        if k == "bin":
            op = e[1]
            a, b, t = _promote(self.ev(e[2], fr), self.ev(e[3], fr))
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if t == "i":
                return _idiv(a, b)
            with np.errstate(divide="ignore", invalid="ignore"):
                return a / b
        if k == "neg":
            return -self.ev(e[1], fr)
        if k == "pow":
            base = self.ev(e[1], fr)
            ex = self.ev(e[2], fr)
            tb, te = _tcls(base), _tcls(ex)
            if te == "i":
                if tb == "i":
                    return int(base) ** int(ex) if int(ex) >= 0 else (1 if int(base) == 1 else 0)
                if isinstance(ex, np.ndarray):
                    raise FortranError("array integer exponent")
                if int(ex) == 2:
                    return base * base
                return _powi(base, ex)
            if tb == "i":
                base = _conv(base, te)
            return _pow_real(base, ex)
        if k == "ref":
            return self.ev_ref(e, fr)
        if k == "rel":
            op = e[1]
            a, b = self.ev(e[2], fr), self.ev(e[3], fr)
            if isinstance(a, str) or isinstance(b, str):
                a, b = str(a).rstrip(), str(b).rstrip()
            else:
                a, b, _ = _promote(a, b)
            if op == "==":
                return a == b
            if op == "/=":
                return a != b
            if op == "<":
                return a < b
            if op == "<=":
                return a <= b
            if op == ">":
                return a > b
            return a >= b
        if k == "and":
            a = self.ev(e[1], fr)
            if not isinstance(a, np.ndarray) and not a:
                return False
            b = self.ev(e[2], fr)
            return np.logical_and(a, b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else bool(a) and bool(b)
        if k == "or":
            a = self.ev(e[1], fr)
            if not isinstance(a, np.ndarray) and a:
                return True
            b = self.ev(e[2], fr)
            return np.logical_or(a, b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else bool(a) or bool(b)
        if k == "not":
            a = self.ev(e[1], fr)
            return np.logical_not(a) if isinstance(a, np.ndarray) else (not a)
        if k == "cat":
            return str(self.ev(e[1], fr)) + str(self.ev(e[2], fr))
        if k == "acons":
            items = [self.ev(x, fr) for x in e[1]]
            flat = []
            for it in items:
                if isinstance(it, np.ndarray):
                    flat.extend(it.ravel(order="F"))
                else:
                    flat.append(it)
            ts = {_tcls(x) for x in flat}
            dt = F64 if "r8" in ts else F32 if "r4" in ts else I64
            return np.array(flat, dtype=dt)
        raise FortranError(f"cannot evaluate {e!r}")

    def _index(self, v, subs, fr, name):
        """numpy index tuple for Fortran subscripts; (index, all_scalar, in_bounds)."""
        a = v.a
        if len(subs) != a.ndim:
            raise FortranError(f"rank mismatch in reference to {name}: {len(subs)} subscripts, rank {a.ndim}")
        idx, scalar, ok = [], True, True
        for d, s in enumerate(subs):
            n, lb = a.shape[d], v.lb[d]
            if isinstance(s, tuple) and s[0] == "slice":
                scalar = False
                lo = lb if s[1] is None else int(self.ev(s[1], fr))
                hi = lb + n - 1 if s[2] is None else int(self.ev(s[2], fr))
                st = 1 if s[3] is None else int(self.ev(s[3], fr))
                cnt = max((hi - lo + st) // st, 0)
                if cnt > 0:
                    last = lo + (cnt - 1) * st
                    if min(lo, last) < lb or max(lo, last) > lb + n - 1:
                        raise FortranBoundsError(f"section {name} dim {d + 1}: {lo}:{hi}:{st} outside {lb}:{lb + n - 1}")
                    stop = last - lb + (1 if st > 0 else -1)
                    idx.append(slice(lo - lb, None if stop < 0 else stop, st))
                else:
                    idx.append(slice(0, 0))
            else:
                val = self.ev(s, fr)
                if isinstance(val, np.ndarray):       # vector subscript
                    scalar = False
                    idx.append(val.astype(I64) - lb)
                    continue
                if _tcls(val) != "i":
                    raise FortranError(f"non-integer subscript of {name}")
                k = int(val) - lb
                if k < 0 or k >= n:
                    ok = False
                idx.append(k)
        return tuple(idx), scalar, ok

    def ev_ref(self, e, fr):
        name, subs = e[1], e[2]
        v = fr.get(name)
        if v is not None and v.a.ndim > 0:
            idx, scalar, ok = self._index(v, subs, fr, name)
            if not ok:
                self.oob_count += 1
                if self.oob == "raise":
                    raise FortranBoundsError(f"{name}{tuple(i + l for i, l in zip(idx, v.lb)) if scalar else ''} out of bounds "
                                             f"(shape {v.a.shape}, lower bounds {v.lb})")
                if scalar:
                    return v.a.dtype.type(np.nan) if v.a.dtype.kind == "f" else v.a.dtype.type(_INT_SENTINEL)
                raise FortranBoundsError(f"section of {name} out of bounds")
            return v.a[idx]
        if v is not None and v.a.dtype == object and len(subs) == 1 and isinstance(subs[0], tuple) and subs[0][0] == "slice":
            s = v.a[()]
            lo = 1 if subs[0][1] is None else int(self.ev(subs[0][1], fr))
            hi = len(s) if subs[0][2] is None else int(self.ev(subs[0][2], fr))
            return s[lo - 1:hi]
        u = self.units.get(name)
        if u is not None and u.kind == "function":
            return self.call_function(name, subs, fr)
        f = _INTRINSICS.get(name)
        if f is not None:
            args, kw = [], {}
            for s in subs:
                if isinstance(s, tuple) and s[0] == "kw":
                    kw[s[1]] = self.ev(s[2], fr)
                else:
                    args.append(self.ev(s, fr))
            return f(*args, **kw)
        if name in ("maxval", "minval"):
            x = self.ev(subs[0], fr)
            return (x.max() if name == "maxval" else x.min()) if x.size else (-np.inf if name == "maxval" else np.inf)
        if name in ("maxloc", "minloc"):
            x = self.ev(subs[0], fr)
            k = np.unravel_index((np.argmax if name == "maxloc" else np.argmin)(x.ravel(order="F")), x.shape, order="F")
            return np.array([i + 1 for i in k], dtype=I64)
        if name == "size":
            x = self.ev(subs[0], fr)
            if len(subs) > 1:
                d = subs[1][2] if isinstance(subs[1], tuple) and subs[1][0] == "kw" else subs[1]
                return int(x.shape[int(self.ev(d, fr)) - 1])
            return int(x.size)
        if name == "shape":
            return np.array(self.ev(subs[0], fr).shape, dtype=I64)
        if name == "reshape":
            x = self.ev(subs[0], fr)
            shp = subs[1][2] if isinstance(subs[1], tuple) and subs[1][0] == "kw" else subs[1]
            shp = tuple(int(s) for s in self.ev(shp, fr))
            return np.asfortranarray(np.asarray(x).ravel(order="F")[:int(np.prod(shp))].reshape(shp, order="F"))
        raise FortranError(f"unknown array or function {name!r} (at depth {self.depth})")

    # ---- argument association
    def make_actual(self, e, fr):
        if isinstance(e, tuple) and e[0] == "kw":
            raise FortranError("keyword arguments to user procedures are not supported")
        if e[0] == "name":
            v = fr.get(e[1])
            if v is not None:
                return _Actual("var", var=v)
        elif e[0] == "ref":
            v = fr.get(e[1])
            if v is not None and v.a.ndim > 0:
                idx, scalar, ok = self._index(v, e[2], fr, e[1])
                if scalar:
                    if not ok:
                        self.oob_count += 1
                        if self.oob == "raise":
                            raise FortranBoundsError(f"actual argument {e[1]}{tuple(i + l for i, l in zip(idx, v.lb))} out of bounds")
                        return _Actual("value", value=(v.a.dtype.type(np.nan) if v.a.dtype.kind == "f" else _INT_SENTINEL))
                    return _Actual("elem", var=v, idx=idx)
                if not ok:
                    self.oob_count += 1
                    if self.oob == "raise" or v.a.dtype.kind != "f":
                        raise FortranBoundsError(f"actual argument: section of {e[1]} with a subscript out of bounds "
                                                 f"(0-based {idx}, shape {v.a.shape})")
                    # memory outside the array: undefined values in, stores lost
                    shp = tuple(len(range(*i.indices(n))) for i, n in zip(idx, v.a.shape) if isinstance(i, slice))
                    return _Actual("value", value=np.full(shp, np.nan, dtype=v.a.dtype, order="F"))
                return _Actual("section", var=v, idx=idx)
        return _Actual("value", value=self.ev(e, fr))

    def associate(self, act, dt, shape, lb, where):
        """Storage for a dummy of dtype dt and explicit shape (None = scalar, () entries None = assumed).
        Returns (Var, copy_out or None)."""
        if act.kind == "value":
            val = act.value
            if isinstance(val, np.ndarray):
                src = np.asfortranarray(val)
            else:
                t = _tcls(val)
                src = np.array(val, dtype={"r8": F64, "r4": F32, "i": I64, "l": np.bool_, "c": object}[t])
            parent, off, writeback = src, 0, None
        elif act.kind == "var":
            parent, off, writeback = act.var.a, 0, None
        elif act.kind == "elem":
            parent = act.var.a
            off = int(np.ravel_multi_index(act.idx, parent.shape, order="F"))
            writeback = None
        else:
            sec = act.var.a[act.idx]
            flat = sec.reshape(-1, order="F")
            if np.shares_memory(flat, sec) or sec.size == 0:
                # contiguous section: gfortran passes the address, the dummy may run on into the parent
                if all(isinstance(i, (slice, int)) and (not isinstance(i, slice) or (i.step or 1) == 1) for i in act.idx) and sec.size:
                    first = tuple((i.start or 0) if isinstance(i, slice) else int(i) for i in act.idx)
                    parent = act.var.a
                    off = int(np.ravel_multi_index(first, parent.shape, order="F"))
                    writeback = None
                else:
                    parent, off, writeback = sec, 0, None
            else:
                tmp = np.array(sec, order="F", copy=True)
                parent, off = tmp, 0

                def writeback(sec=sec, tmp=tmp):
                    sec[...] = tmp
        if parent.dtype != dt:
            if parent.dtype == object or dt == object:
                pass
            else:
                raise FortranError(f"{where}: argument type mismatch (actual {parent.dtype}, dummy {np.dtype(dt)}) -- "
                                   f"Fortran would reinterpret the bits")
        if shape is None:                      # scalar dummy
            if parent.ndim == 0:
                return Var(parent, ()), writeback
            flat = parent.reshape(-1, order="F")
            if not np.shares_memory(flat, parent):
                raise FortranError(f"{where}: non-contiguous storage")
            return Var(flat[off:off + 1].reshape(()), ()), writeback
        if any(s is None for s in shape):      # assumed shape: take the actual's
            src = parent if act.kind != "elem" else None
            if src is None or src.ndim != len(shape):
                raise FortranError(f"{where}: assumed-shape dummy needs an array actual of the same rank")
            return Var(src, lb), writeback
        n = 1
        for s in shape:
            n *= s
        flat = parent.reshape(-1, order="F") if parent.ndim else parent.reshape(1)
        if parent.ndim and not np.shares_memory(flat, parent) and parent.size:
            raise FortranError(f"{where}: non-contiguous storage")
        if off + n > flat.size:
            self.oob_count += 1
            if self.oob == "raise":
                raise FortranBoundsError(f"{where}: dummy of {n} elements, only {flat.size - off} left in the actual")
            tmp = np.full(n, np.nan if flat.dtype.kind == "f" else 0, dtype=flat.dtype)
            avail = max(flat.size - off, 0)
            tmp[:avail] = flat[off:off + avail]
            prev = writeback

            def writeback(flat=flat, tmp=tmp, off=off, avail=avail, prev=prev):
                flat[off:off + avail] = tmp[:avail]
                if prev:
                    prev()
            return Var(tmp.reshape(shape, order="F"), lb), writeback
        return Var(flat[off:off + n].reshape(shape, order="F"), lb), writeback

    # ---- procedure calls
    def _plan(self, u):
        """Per-unit declaration plan: [(name, base, kind, clen, dims, init, role)]."""
        if u.plan is None:
            plan = []
            for d in u.decls:
                _, base, kind, clen, attrs, ents = d
                for nm, dims, init in ents:
                    dims = dims if dims is not None else attrs.get("dimension")
                    if nm in u.args:
                        role = "dummy"
                    elif nm == u.result:
                        role = "result"
                    elif dims is None and init is None and nm in self.units and self.units[nm].kind == "function":
                        role = "external"
                    elif init is not None or attrs.get("save"):
                        role = "static"
                    else:
                        role = "local"
                    plan.append((nm, base, kind, clen, dims, init, role))
            declared = {p[0] for p in plan}
            for a in u.args:
                if a not in declared:     # implicit typing is not used by the reference
                    raise FortranError(f"{u.file}:{u.line}: dummy {a} of {u.name} has no declaration")
            # scalars first so that array bounds can use them
            u.plan = ([p for p in plan if p[6] == "dummy" and p[4] is None] +
                      [p for p in plan if not (p[6] == "dummy" and p[4] is None)])
        return u.plan

    def invoke(self, u, actuals):
        self.calls[u.name] = self.calls.get(u.name, 0) + 1
        if len(actuals) != len(u.args):
            raise FortranError(f"{u.name}: {len(actuals)} actual arguments for {len(u.args)} dummies")
        fr = {}
        for m in u.uses:
            fr.update(self.module(m))
        amap = dict(zip(u.args, actuals))
        outs = []
        where = f"{u.file}:{u.line} {u.name}"
        for nm, base, kind, clen, dims, init, role in self._plan(u):
            if role == "external":
                continue
            if role == "dummy":
                dt = _DTYPES.get((base, kind)) or _DTYPES[(base, None)]
                if dims is None:
                    shape, lb = None, ()
                else:
                    shape, lbs = [], []
                    for d in dims:
                        if isinstance(d, tuple) and d[0] == "slice":
                            lo = 1 if d[1] is None else int(self.ev(d[1], fr))
                            hi = None if d[2] is None else int(self.ev(d[2], fr))
                        else:
                            lo, hi = 1, int(self.ev(d, fr))
                        lbs.append(lo)
                        shape.append(None if hi is None else max(hi - lo + 1, 0))
                    shape, lb = tuple(shape), tuple(lbs)
                v, wb = self.associate(amap[nm], dt, shape, lb, f"{where} dummy {nm}")
                if base == "character":
                    v.clen = None
                fr[nm] = v
                if wb:
                    outs.append(wb)
            elif role == "static":
                if nm not in u.static:
                    v = self._alloc(fr, base, kind, clen, dims, where)
                    if init is not None:
                        self._store_whole(v, self.ev(init, fr))
                    u.static[nm] = v
                fr[nm] = u.static[nm]
            else:
                fr[nm] = self._alloc(fr, base, kind, clen, dims, where)
        self.depth += 1
        try:
            hook = self.hooks.get(u.name)
            if hook is not None:
                hook(self, fr)
            else:
                self.run(u.body, fr)
        except _Return:
            pass
        finally:
            self.depth -= 1
        for wb in outs:
            wb()
        return fr

    def call_function(self, name, subs, fr):
        u = self.units[name]
        callee = self.invoke(u, [self.make_actual(s, fr) for s in subs])
        r = callee[u.result]
        return r.a[()] if r.a.ndim == 0 else r.a.copy()

    def call(self, name, *args):
        """Call a subroutine/function from Python.  numpy arrays are passed by reference (they must be
        Fortran-ordered to be updated in place), Python/numpy scalars by value (wrap in a 0-d array to get them back)."""
        u = self.units[name]
        acts = []
        for a in args:
            if isinstance(a, Var):
                acts.append(_Actual("var", var=a))
            elif isinstance(a, np.ndarray):
                if a.ndim > 1 and not a.flags.f_contiguous:
                    raise FortranError("pass Fortran-ordered arrays")
                acts.append(_Actual("var", var=Var(a)))
            elif isinstance(a, str):
                acts.append(_Actual("value", value=a))
            elif isinstance(a, (int, np.integer)) and not isinstance(a, (bool, np.bool_)):
                acts.append(_Actual("value", value=int(a)))
            else:
                acts.append(_Actual("value", value=a))
        fr = self.invoke(u, acts)
        if u.kind == "function":
            r = fr[u.result]
            return r.a[()] if r.a.ndim == 0 else r.a.copy()
        return {k: (v.a[()] if v.a.ndim == 0 else v.a) for k, v in fr.items()}     # the callee's variables at return

    def run_program(self, name):
        """Execute a main program; returns its variables at the end."""
        fr = self.invoke(self.units[name], [])
        return {k: (v.a[()] if v.a.ndim == 0 else v.a) for k, v in fr.items()}

    # ---- statements
    def assign(self, lhs, val, fr):
        if lhs[0] == "name":
            v = fr.get(lhs[1])
            if v is None:
                raise FortranError(f"assignment to undeclared {lhs[1]!r}")
            if v.a.dtype == object:
                self._store_whole(v, val)
            elif v.a.ndim == 0:
                if isinstance(val, np.ndarray) and val.ndim:
                    raise FortranError(f"array assigned to scalar {lhs[1]}")
                v.a[()] = val
            else:
                if isinstance(val, np.ndarray) and val.shape != v.a.shape:
                    raise FortranError(f"shape mismatch assigning to {lhs[1]}: {val.shape} -> {v.a.shape}")
                v.a[...] = val
            return
        if lhs[0] != "ref":
            raise FortranError(f"bad assignment target {lhs!r}")
        name = lhs[1]
        v = fr.get(name)
        if v is None:
            raise FortranError(f"assignment to undeclared {name!r}")
        if v.a.ndim == 0 and v.a.dtype == object:
            return        # substring assignment: strings only feed file names
        idx, scalar, ok = self._index(v, lhs[2], fr, name)
        if not ok:
            self.oob_count += 1
            if self.oob == "raise":
                raise FortranBoundsError(f"store to {name}{tuple(i + l for i, l in zip(idx, v.lb))} out of bounds (shape {v.a.shape})")
            return
        if not scalar and isinstance(val, np.ndarray) and val.ndim and val.shape != v.a[idx].shape:
            raise FortranError(f"shape mismatch assigning to section of {name}: {val.shape} -> {v.a[idx].shape}")
        v.a[idx] = val

    def run(self, body, fr):
        for st in body:
            k = st[0]
            if self.probes and st[-1] in self.probes:
                self.probes[st[-1]]({n: v.a for n, v in fr.items()})
            try:
                if k == "assign":
                    self.assign(st[1], self.ev(st[2], fr), fr)
                elif k == "call":
                    u = self.units.get(st[1])
                    if u is None:
                        raise FortranError(f"call of unknown subroutine {st[1]!r}")
                    self.invoke(u, [self.make_actual(a, fr) for a in st[2]])
                elif k == "if":
                    for cond, blk in st[1]:
                        c = self.ev(cond, fr)
                        if isinstance(c, np.ndarray):
                            raise FortranError("array-valued IF condition")
                        if c:
                            self.run(blk, fr)
                            break
                    else:
                        if st[2] is not None:
                            self.run(st[2], fr)
                elif k == "do":
                    _, var, a, b, c, blk, _no = st
                    v = fr[var]
                    lo, hi = int(self.ev(a, fr)), int(self.ev(b, fr))
                    step = 1 if c is None else int(self.ev(c, fr))
                    n = max((hi - lo + step) // step, 0)
                    i = lo
                    v.a[()] = i
                    try:
                        for _ in range(n):
                            try:
                                self.run(blk, fr)
                            except _Cycle:
                                pass
                            i += step
                            v.a[()] = i
                    except _Exit:
                        pass
                elif k == "dowhile":
                    try:
                        while self.ev(st[1], fr):
                            try:
                                self.run(st[2], fr)
                            except _Cycle:
                                pass
                    except _Exit:
                        pass
                elif k == "select":
                    sel = self.ev(st[1], fr)
                    if isinstance(sel, str):
                        sel = sel.rstrip()
                    hit = None
                    for vals, blk in st[2]:
                        for cv in vals:
                            if isinstance(cv, tuple) and cv[0] == "slice":
                                lo = None if cv[1] is None else self.ev(cv[1], fr)
                                hi = None if cv[2] is None else self.ev(cv[2], fr)
                                if (lo is None or sel >= lo) and (hi is None or sel <= hi):
                                    hit = blk
                            else:
                                c = self.ev(cv, fr)
                                if isinstance(c, str):
                                    c = c.rstrip()
                                if sel == c:
                                    hit = blk
                            if hit is not None:
                                break
                        if hit is not None:
                            break
                    if hit is None:
                        hit = st[3]
                    if hit is not None:
                        self.run(hit, fr)
                elif k == "return":
                    raise _Return()
                elif k == "exit":
                    raise _Exit()
                elif k == "cycle":
                    raise _Cycle()
                elif k == "stop":
                    raise FortranStop()
                else:
                    raise FortranError(f"unknown statement kind {k}")
            except FortranError as ex:
                if not getattr(ex, "_located", False):
                    ex._located = True
                    ex.args = (f"{ex.args[0]}  [at line {st[-1]}]",) + ex.args[1:]
                raise


if __name__ == "__main__":
    it = Interp()
    for f in sys.argv[1:]:
        it.load(f)
    print("modules:", sorted(it.modules))
    print("procedures:", sorted(it.units))

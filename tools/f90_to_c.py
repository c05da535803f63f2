#!/usr/bin/env python3
"""f90_to_c.py -- source-to-source translator from the Fortran 90 subset RRTMG is written in to C (gnu11).

TEST INFRASTRUCTURE.  The image has no Fortran compiler, so the reference's own RRTMG sources cannot be compiled as
they are.  This tool reads them where they lie (under /root/reference, at build time, driven by oracle/Makefile) and
writes an equivalent C translation unit per source file into oracle/_ref/src/ -- statement by statement, expression
by expression, nothing reordered -- which gcc then compiles into oracle/_ref/librrtmg_ref.so.  The hand-written
oracle (oracle/*.c) is validated bit for bit against that library (tests/test_ref_translation.py).  Nothing under
mima_b200/ uses either.  No reference source is copied into the repository: the generated C lives only under the
git-ignored oracle/_ref/.

The dialect handled is what the non-McICA RRTMG_LW / RRTMG_SW sources use:
  * modules with `use ..., only: a => b`, `implicit none`, `save`, `public`, `equivalence`, parameters, module arrays;
  * module subroutines and internal subroutines (`contains` inside a subroutine: emitted as GCC nested functions, so
    host association needs no rewriting);
  * integer / real(kind=rb) / logical scalars and arrays: explicit-shape, automatic, assumed-shape `(:)`, `(0:)`;
    `intent`, `optional` + `present()`, `parameter`, `dimension(...)`;
  * assignment, whole-array and array-section assignment from a scalar or an array constructor `(/ ... /)`;
  * `do` (with step), `do while`, `if/else if/else`, one-line `if`, `goto` + labels, `call`, `return`, `exit`, `cycle`,
    `stop 'message'`;
  * expressions with `**`, dotted and symbolic relational operators, `.and. .or. .not.`, kind-suffixed literals,
    and the intrinsics int, real, dble, float, nint, aint, min, max, mod, abs, sign, exp, log, log10, sqrt, sin, cos, tan,
    asin, acos, atan, present, epsilon, tiny, huge.
Semantics kept: column-major storage with declared lower bounds; integer division and real->integer assignment truncate
toward zero (same as C); `x**n` with an integer n is evaluated by repeated multiplication (square and multiply), with
a real exponent by pow(); every real literal is a C double, which is what the reference's build makes of them
(`-r8`, CMakeLists.txt:96: default real = 8 bytes); scalars with intent(in) are passed by value, all other
arguments by reference.  Automatic arrays become C variable-length arrays; local variables are zero-initialised (Fortran leaves them
undefined).
Anything outside the subset stops the translation with the file and line number.

usage: f90_to_c.py -o OUTDIR file.f90 [file.f90 ...]
"""
from __future__ import annotations

import os
import re
import sys

C_RESERVED = {
    "auto", "break", "case", "char", "const", "continue", "default", "do", "double", "else", "enum", "extern", "float",
    "for", "goto", "if", "inline", "int", "long", "register", "restrict", "return", "short", "signed", "sizeof", "static",
    "struct", "switch", "typedef", "union", "unsigned", "void", "volatile", "while", "main", "index", "pi", "abs", "exp",
    "log", "pow", "sqrt", "min", "max", "y0", "y1", "j0", "j1", "jn", "yn", "gamma", "time", "signal", "div", "remainder",
}


class TranslateError(Exception):
    pass


# ------------------------------------------------------------------------------------------------ reading
def _strip_comment(line: str) -> str:
    q = None
    for i, c in enumerate(line):
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "!":
            return line[:i]
    return line


def _lower_outside_strings(s: str) -> str:
    out, q = [], None
    for c in s:
        if q:
            out.append(c)
            if c == q:
                q = None
        else:
            if c in "'\"":
                q = c
                out.append(c)
            else:
                out.append(c.lower())
    return "".join(out)


def read_statements(path: str):
    """Logical statements of a free-form source file: [(line number, label or None, text)]."""
    stmts = []
    cur, cur_line = "", 0
    with open(path, encoding="latin-1") as f:
        for n, raw in enumerate(f, 1):
            line = _strip_comment(raw.rstrip("\n").replace("\t", " ")).strip()
            if not line or line.startswith("#"):
                continue
            if cur:
                if line.startswith("&"):
                    line = line[1:].lstrip()
                cur += " " + line
            else:
                cur, cur_line = line, n
            if cur.endswith("&"):
                cur = cur[:-1].rstrip()
                continue
            text = _lower_outside_strings(cur)
            cur = ""
            # split on ';' outside strings
            parts, q, b = [], None, 0
            for i, c in enumerate(text):
                if q:
                    if c == q:
                        q = None
                elif c in "'\"":
                    q = c
                elif c == ";":
                    parts.append(text[b:i])
                    b = i + 1
            parts.append(text[b:])
            for p in parts:
                p = p.strip()
                if not p:
                    continue
                m = re.match(r"^(\d+)\s+(.*)$", p)
                label = None
                if m:
                    label, p = m.group(1), m.group(2)
                stmts.append((cur_line, label, p))
    return stmts


# ------------------------------------------------------------------------------------------------ expressions
DOTOPS = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "/=", ".and.": ".and.",
          ".or.": ".or.", ".not.": ".not.", ".true.": ".true.", ".false.": ".false.", ".eqv.": ".eqv.", ".neqv.": ".neqv."}
_DOT_RE = re.compile(r"\.[a-z]+\.")


def tokenize(s: str):
    toks, i, n = [], 0, len(s)
    while i < n:
        c = s[i]
        if c.isspace():
            i += 1
            continue
        if c in "'\"":
            j = s.index(c, i + 1)
            toks.append(("str", s[i + 1:j]))
            i = j + 1
            continue
        if c.isdigit() or (c == "." and i + 1 < n and s[i + 1].isdigit()):
            j = i
            while j < n and s[j].isdigit():
                j += 1
            real = False
            if j < n and s[j] == ".":
                m = _DOT_RE.match(s, j)
                if not (m and m.group(0) in DOTOPS):
                    real = True
                    j += 1
                    while j < n and s[j].isdigit():
                        j += 1
            if j < n and s[j] in "ed":
                k = j + 1
                if k < n and s[k] in "+-":
                    k += 1
                if k < n and s[k].isdigit():
                    real = True
                    while k < n and s[k].isdigit():
                        k += 1
                    j = k
            text = s[i:j].replace("d", "e")
            if j < n and s[j] == "_":           # kind suffix
                k = j + 1
                while k < n and (s[k].isalnum() or s[k] == "_"):
                    k += 1
                j = k
            toks.append(("real" if real else "int", text))
            i = j
            continue
        if c == ".":
            m = _DOT_RE.match(s, i)
            if not m or m.group(0) not in DOTOPS:
                raise TranslateError("bad token at " + s[i:i + 10])
            toks.append(("op", DOTOPS[m.group(0)]))
            i = m.end()
            continue
        if c.isalpha() or c == "_":
            j = i
            while j < n and (s[j].isalnum() or s[j] == "_"):
                j += 1
            toks.append(("name", s[i:j]))
            i = j
            continue
        two = s[i:i + 2]
        if two in ("**", "==", "/=", "<=", ">=", "(/", "/)", "=>", "::"):
            toks.append(("op", two))
            i += 2
            continue
        if c in "+-*/(),=<>:%":
            toks.append(("op", c))
            i += 1
            continue
        raise TranslateError("bad character %r in %r" % (c, s))
    toks.append(("end", ""))
    return toks


class Parser:
    """Recursive descent over a token list; produces tuples:
    ('int', text) ('real', text) ('str', s) ('bool', 0/1) ('name', id) ('call', id, [args]) ('un', op, e)
    ('bin', op, a, b) ('sec', lo, hi, step) ('kw', name, e) ('ctor', [items]) ('ido', [items], var, a, b, c)"""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, kind, val=None):
        k, v = self.t[self.i]
        if k == kind and (val is None or v == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        if not self.accept(kind, val):
            raise TranslateError("expected %s %s, got %s" % (kind, val, self.t[self.i],))

    def expr(self):
        left = self.or_()
        while self.peek() in (("op", ".eqv."), ("op", ".neqv.")):
            op = self.next()[1]
            left = ("bin", op, left, self.or_())
        return left

    def or_(self):
        left = self.and_()
        while self.accept("op", ".or."):
            left = ("bin", ".or.", left, self.and_())
        return left

    def and_(self):
        left = self.not_()
        while self.accept("op", ".and."):
            left = ("bin", ".and.", left, self.not_())
        return left

    def not_(self):
        if self.accept("op", ".not."):
            return ("un", ".not.", self.not_())
        return self.rel()

    def rel(self):
        left = self.add()
        k, v = self.peek()
        if k == "op" and v in ("<", "<=", ">", ">=", "==", "/="):
            self.next()
            return ("bin", v, left, self.add())
        return left

    def add(self):
        k, v = self.peek()
        if k == "op" and v in "+-" and v:
            self.next()
            left = ("un", v, self.mul())
        else:
            left = self.mul()
        while True:
            k, v = self.peek()
            if k == "op" and v in ("+", "-"):
                self.next()
                left = ("bin", v, left, self.mul())
            else:
                return left

    def mul(self):
        left = self.pow_()
        while True:
            k, v = self.peek()
            if k == "op" and v in ("*", "/"):
                self.next()
                left = ("bin", v, left, self.pow_())
            else:
                return left

    def pow_(self):
        base = self.primary()
        if self.accept("op", "**"):
            k, v = self.peek()
            if k == "op" and v in ("+", "-"):
                self.next()
                return ("bin", "**", base, ("un", v, self.pow_()))
            return ("bin", "**", base, self.pow_())
        return base

    def arg(self):
        # section, keyword argument or expression
        if self.peek() == ("op", ":"):
            lo = None
        else:
            if self.peek()[0] == "name" and self.t[self.i + 1] == ("op", "="):
                name = self.next()[1]
                self.next()
                return ("kw", name, self.expr())
            lo = self.expr()
            if self.peek() != ("op", ":"):
                return lo
        self.expect("op", ":")
        hi = step = None
        if self.peek()[0] != "end" and self.peek() not in (("op", ","), ("op", ")"), ("op", ":")):
            hi = self.expr()
        if self.accept("op", ":"):
            step = self.expr()
        return ("sec", lo, hi, step)

    def primary(self):
        k, v = self.next()
        if k in ("int", "real", "str"):
            return (k, v)
        if k == "op" and v in (".true.", ".false."):
            return ("bool", 1 if v == ".true." else 0)
        if k == "name":
            if self.accept("op", "("):
                args = []
                if not self.accept("op", ")"):
                    while True:
                        args.append(self.arg())
                        if self.accept("op", ")"):
                            break
                        self.expect("op", ",")
                return ("call", v, args)
            return ("name", v)
        if k == "op" and v == "(":
            e = self.expr()
            self.expect("op", ")")
            return ("par", e)
        if k == "op" and v == "(/":
            items = []
            while True:
                items.append(self.ctor_item())
                if self.accept("op", "/)"):
                    break
                self.expect("op", ",")
            return ("ctor", items)
        raise TranslateError("unexpected token %s %r" % (k, v))

    def ctor_item(self):
        # implied do: ( item, ..., i = a, b [, c] )
        if self.peek() == ("op", "("):
            save = self.i
            try:
                self.next()
                items = [self.expr()]
                while self.accept("op", ","):
                    if self.peek()[0] == "name" and self.t[self.i + 1] == ("op", "="):
                        var = self.next()[1]
                        self.next()
                        a = self.expr()
                        self.expect("op", ",")
                        b = self.expr()
                        c = self.expr() if self.accept("op", ",") else None
                        self.expect("op", ")")
                        return ("ido", items, var, a, b, c)
                    items.append(self.expr())
            except TranslateError:
                pass
            self.i = save
        return self.expr()


def parse_expr(s: str):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.peek()[0] != "end":
        raise TranslateError("trailing tokens in expression %r" % s)
    return e


# ------------------------------------------------------------------------------------------------ program structure
class Var:
    def __init__(self, name):
        self.name = name
        self.cname = name + "_" if name in C_RESERVED else name
        self.ctype = "double"
        self.dims = None          # list of (lo ast or None, hi ast or None) ; hi None = assumed
        self.intent = None
        self.optional = False
        self.param = None         # ast of the parameter value
        self.init = None
        self.where = "local"      # 'local' | 'dummy' | 'module'
        self.module = None
        self.alias_of = None      # equivalence
        self.ibounds = None       # module arrays: [(lo, n)] evaluated
        self.ival = None          # integer parameter value
        self.is_char = False

    @property
    def byval(self):
        return self.where == "dummy" and self.dims is None and self.intent == "in" and not self.optional

    def is_array(self):
        return self.dims is not None


class Unit:
    def __init__(self, kind, name, parent=None):
        self.kind, self.name, self.parent = kind, name, parent      # kind: 'module' | 'sub'
        self.vars = {}
        self.uses = []           # (module, None | {local: remote})
        self.args = []
        self.subs = {}
        self.sub_order = []
        self.body = []           # executable statements (line, label, text)
        self.decl_order = []
        self.path = ""
        self.equiv = []
        self.fn_type = None      # C type of the result when the unit is a function
        self.data = []           # data statements: (name, [value asts])
        self.line = 0

    def module_of(self):
        u = self
        while u and u.kind != "module":
            u = u.parent
        return u

    def cname(self):
        if self.kind == "sub" and self.parent is not None and self.parent.kind == "sub":
            return self.name + "_" if self.name in C_RESERVED else self.name       # nested function
        m = self.module_of()
        return (m.name + "__" if m else "") + self.name


DECL_RE = re.compile(r"^(integer|real|logical|character|double\s+precision)\b")


def split_top(s: str, sep=","):
    """Split on `sep` outside parentheses, constructors and strings."""
    out, depth, q, b = [], 0, None, 0
    i = 0
    while i < len(s):
        c = s[i]
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
        elif c == sep and depth == 0:
            out.append(s[b:i])
            b = i + 1
        i += 1
    out.append(s[b:])
    return [x.strip() for x in out]


def match_paren(s: str, i: int) -> int:
    """Index of the ')' matching the '(' at s[i]."""
    depth, q = 0, None
    for j in range(i, len(s)):
        c = s[j]
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return j
    raise TranslateError("unbalanced parentheses in %r" % s)


def parse_dims(text: str):
    dims = []
    for d in split_top(text):
        if ":" in d and not d.startswith("("):
            parts = split_top(d, ":")
            lo = parse_expr(parts[0]) if parts[0] else None
            hi = parse_expr(parts[1]) if len(parts) > 1 and parts[1] else None
            if lo is None and hi is None:
                dims.append((None, None))
            else:
                dims.append((lo, hi))
        elif d == "*":
            dims.append((None, None))
        else:
            dims.append((None, parse_expr(d)))
    return dims


class Program:
    def __init__(self):
        self.modules = {}
        self.order = []
        self.externals = set()

    # -------------------------------------------------------------------------------------------- parsing
    def parse_file(self, path):
        stmts = read_statements(path)
        stack = []
        for (ln, label, text) in stmts:
            try:
                self._stmt(path, stack, ln, label, text)
            except TranslateError as e:
                raise TranslateError("%s:%d: %s" % (path, ln, e))
        if stack:
            raise TranslateError("%s: unterminated %s %s" % (path, stack[-1].kind, stack[-1].name))

    def _stmt(self, path, stack, ln, label, text):
        cur = stack[-1] if stack else None
        m = re.match(r"^module\s+(\w+)$", text)
        if m and not text.startswith("module procedure"):
            u = Unit("module", m.group(1))
            u.path = path
            if u.name in self.modules:          # parkind.f90 exists once per code (LW, SW): the later copy replaces the earlier
                self.order.remove(self.modules[u.name])
            self.modules[u.name] = u
            self.order.append(u)
            stack.append(u)
            return
        m = re.match(r"^(?:recursive\s+)?subroutine\s+(\w+)\s*(?:\((.*)\))?$", text)
        if m:
            u = Unit("sub", m.group(1), cur)
            u.path = path
            u.line = ln
            u.args = [a for a in split_top(m.group(2) or "") if a]
            if cur is None:
                # external subroutine (not in a module): collected in a pseudo-module, C name f90ext__<name>
                cur = self.modules.get("f90ext")
                if cur is None:
                    cur = Unit("module", "f90ext")
                    self.modules["f90ext"] = cur
                    self.order.append(cur)
                cur.path = path
                u.parent = cur
            cur.subs[u.name] = u
            cur.sub_order.append(u)
            stack.append(u)
            return
        m = re.match(r"^(?:(integer|real|logical|double\s+precision)\s*(?:\([^)]*\))?\s+)?function\s+(\w+)\s*\((.*)\)$", text)
        if m:
            u = Unit("sub", m.group(2), cur)
            u.path, u.line = path, ln
            u.args = [a for a in split_top(m.group(3) or "") if a]
            u.fn_type = "int" if (m.group(1) or "real").startswith(("integer", "logical")) else "double"
            res = Var(u.name)
            res.ctype = u.fn_type
            u.vars[u.name] = res
            u.decl_order.append(u.name)
            if cur is None:
                raise TranslateError("function outside a module")
            cur.subs[u.name] = u
            cur.sub_order.append(u)
            stack.append(u)
            return
        if re.match(r"^end\s*(module|subroutine|function)\b", text) or text == "end":
            stack.pop()
            return
        if re.match(r"^(end\s*)?(program|interface|type\b)", text) and not DECL_RE.match(text):
            raise TranslateError("unsupported program unit: " + text)
        if cur is None:
            raise TranslateError("statement outside a program unit: " + text)
        if text in ("implicit none", "save", "contains", "private", "public") or text.startswith("implicit "):
            return
        if re.match(r"^(public|private)\b", text):
            return
        m = re.match(r"^use\s+(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", text)
        if m:
            only = None
            if m.group(2) is not None:
                only = {}
                for item in split_top(m.group(2)):
                    if not item:
                        continue
                    if "=>" in item:
                        a, b = [x.strip() for x in item.split("=>")]
                        only[a] = b
                    else:
                        only[item] = item
            cur.uses.append((m.group(1), only))
            return
        m = re.match(r"^data\s+(\w+)\s*/(.*)/$", text)
        if m:
            vals = []
            for item in split_top(m.group(2)):
                rep = re.match(r"^(\d+)\s*\*\s*(.*)$", item)
                if rep:
                    vals += [parse_expr(rep.group(2))] * int(rep.group(1))
                else:
                    vals.append(parse_expr(item))
            cur.data.append((m.group(1), vals))
            return
        if text.startswith("equivalence"):
            body = text[len("equivalence"):].strip()
            for grp in split_top(body):
                inner = grp.strip()[1:-1]
                a, b = split_top(inner)
                cur.equiv.append((parse_expr(a), parse_expr(b)))
            return
        if DECL_RE.match(text) and ("::" in text or re.match(r"^(integer|real|logical)\s+[a-z_]", text)):
            self._decl(cur, text)
            return
        cur.body.append((ln, label, text))

    def _decl(self, unit, text):
        if "::" in text:
            left, right = text.split("::", 1)
        else:
            m = re.match(r"^(integer|real|logical)\s+(.*)$", text)
            left, right = m.group(1), m.group(2)
        attrs = split_top(left)
        t = attrs[0]
        ctype = "double"
        is_char = False
        if t.startswith("integer") or t.startswith("logical"):
            ctype = "int"
        elif t.startswith("character"):
            is_char = True
        intent, optional, param, dim = None, False, False, None
        for a in attrs[1:]:
            a = a.strip()
            m = re.match(r"^intent\s*\(\s*(\w+)\s*\)$", a)
            if m:
                intent = m.group(1)
            elif a == "optional":
                optional = True
            elif a == "parameter":
                param = True
            elif a.startswith("dimension"):
                dim = a[a.index("(") + 1:match_paren(a, a.index("("))]
            elif a in ("save", "public", "private", "target"):
                pass
            else:
                raise TranslateError("unsupported attribute %r" % a)
        for ent in split_top(right):
            m = re.match(r"^(\w+)\s*(\(.*?\))?\s*(?:=\s*(.*))?$", ent)
            if not m:
                raise TranslateError("cannot parse declaration entity %r" % ent)
            name = m.group(1)
            dims_txt = None
            rest = ent[len(name):].strip()
            init = None
            if rest.startswith("("):
                j = match_paren(rest, 0)
                dims_txt = rest[1:j]
                rest = rest[j + 1:].strip()
            if rest.startswith("="):
                init = rest[1:].strip()
            elif rest:
                raise TranslateError("cannot parse declaration entity %r" % ent)
            v = unit.vars.get(name) or Var(name)
            v.ctype, v.is_char = ctype, is_char
            v.intent, v.optional = intent, optional
            dd = dims_txt if dims_txt is not None else dim
            if dd is not None:
                v.dims = parse_dims(dd)
            if init is not None and not is_char:
                if param:
                    v.param = parse_expr(init)
                else:
                    v.init = parse_expr(init)
            if unit.kind == "module":
                v.where, v.module = "module", unit.name
            elif name in unit.args:
                v.where = "dummy"
            unit.vars[name] = v
            unit.decl_order.append(name)

    # -------------------------------------------------------------------------------------------- name resolution
    def lookup(self, unit, name, _seen=None):
        """Variable `name` as seen from `unit` (host association, then use association)."""
        u = unit
        while u is not None:
            if name in u.vars:
                return u.vars[name]
            for (mod, only) in u.uses:
                m = self.modules.get(mod)
                if m is None:
                    continue
                if only is None:
                    v = self.lookup_module(m, name)
                    if v is not None:
                        return v
                elif name in only:
                    v = self.lookup_module(m, only[name])
                    if v is not None:
                        return v
            u = u.parent
        return None

    def lookup_module(self, m, name, depth=0):
        if name in m.vars:
            return m.vars[name]
        if depth > 4:
            return None
        for (mod, only) in m.uses:
            mm = self.modules.get(mod)
            if mm is None:
                continue
            if only is None:
                v = self.lookup_module(mm, name, depth + 1)
                if v is not None:
                    return v
            elif name in only:
                v = self.lookup_module(mm, only[name], depth + 1)
                if v is not None:
                    return v
        return None

    def find_sub(self, unit, name):
        u = unit
        while u is not None:
            if name in u.subs:
                return u.subs[name]
            for (mod, only) in u.uses:
                m = self.modules.get(mod)
                if m is None:
                    continue
                if only is None and name in m.subs:
                    return m.subs[name]
                if only is not None and name in only and only[name] in m.subs:
                    return m.subs[only[name]]
            u = u.parent
        cands = [m.subs[name] for m in self.order if name in m.subs]
        if len(cands) == 1:
            return cands[0]
        return None

    # -------------------------------------------------------------------------------------------- constants
    def const_int(self, unit, e):
        k = e[0]
        if k == "int":
            return int(e[1])
        if k == "par":
            return self.const_int(unit, e[1])
        if k == "un":
            v = self.const_int(unit, e[2])
            return -v if e[1] == "-" else v
        if k == "bin":
            a, b = self.const_int(unit, e[2]), self.const_int(unit, e[3])
            op = e[1]
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "/":
                return int(a / b)
            if op == "**":
                return a ** b
        if k == "name":
            v = self.lookup(unit, e[1])
            if v is not None and v.param is not None and v.ctype == "int":
                if v.ival is None:
                    owner = self.modules[v.module] if v.module else unit
                    v.ival = self.const_int(owner, v.param)
                return v.ival
        if k == "call" and e[1] in ("selected_int_kind", "selected_real_kind", "kind"):
            return 0
        raise TranslateError("not an integer constant expression: %r" % (e,))


# ------------------------------------------------------------------------------------------------ C emission
PRELUDE = r"""/* generated by tools/f90_to_c.py -- do not edit, do not commit (oracle/_ref/ is git-ignored) */
#ifndef F90REF_RT_H
#define F90REF_RT_H
#include <math.h>
#include <float.h>
#include <limits.h>
#include <string.h>
#include <stdlib.h>
static inline int f_powii(int a, int n) { int r = 1; if (n < 0) return a == 1 ? 1 : (a == -1 ? ((n & 1) ? -1 : 1) : 0); while (n) { if (n & 1) r *= a; a *= a; n >>= 1; } return r; }
static inline double f_powdi(double a, int n)
{
    unsigned m = n < 0 ? (unsigned)(-(long)n) : (unsigned)n;
    double r = 1.0, p = a;
    int first = 1;
    while (m) { if (m & 1u) { r = first ? p : r * p; first = 0; } m >>= 1; if (m) p = p * p; }
    return n < 0 ? 1.0 / r : r;
}
static inline int f_modi(int a, int b) { return a % b; }
static inline int f_signi(int a, int b) { int m = a < 0 ? -a : a; return b >= 0 ? m : -m; }
static inline double f_signd(double a, double b) { double m = fabs(a); return signbit(b) ? -m : m; }
static inline int f_absi(int a) { return a < 0 ? -a : a; }
#define F_POW(a, b) _Generic((b), int: _Generic((a), int: f_powii, default: f_powdi), default: pow)((a), (b))
#define F_MOD(a, b) _Generic((a) + (b), int: f_modi, default: fmod)((a), (b))
#define F_ABS(a) _Generic((a), int: f_absi, default: fabs)(a)
#define F_SIGN(a, b) _Generic((a) + (b), int: f_signi, default: f_signd)((a), (b))
#define F_MIN(a, b) ((a) < (b) ? (a) : (b))
#define F_MAX(a, b) ((a) > (b) ? (a) : (b))
void f90_stop(const char *msg);
#endif
"""


class Emitter:
    def __init__(self, prog: Program):
        self.p = prog
        self.tmp = 0

    # -------------------------------------------------------------------------------------------- variables
    def var_bounds(self, unit, v):
        """[(lo C text, extent C text)] of an array variable as seen inside `unit`."""
        if v.where == "module":
            if v.ibounds is None:
                owner = self.p.modules[v.module]
                b = []
                for (lo, hi) in v.dims:
                    l = self.p.const_int(owner, lo) if lo is not None else 1
                    h = self.p.const_int(owner, hi)
                    b.append((l, h - l + 1))
                v.ibounds = b
            return [(str(l), str(n)) for (l, n) in v.ibounds]
        return [("%s_l%d" % (v.cname, k + 1), "%s_n%d" % (v.cname, k + 1)) for k in range(len(v.dims))]

    def var_ref(self, v):
        if v.where == "module":
            return "%s__%s" % (v.module, v.name)
        return v.cname

    def scalar(self, unit, v):
        if v.where == "dummy" and not v.byval:
            return "(*%s)" % v.cname
        return self.var_ref(v)

    def index(self, unit, v, args):
        b = self.var_bounds(unit, v)
        if len(args) != len(b):
            raise TranslateError("rank mismatch indexing %s" % v.name)
        s = ""
        for k in reversed(range(len(b))):
            term = "((%s) - %s)" % (self.expr(unit, args[k]), b[k][0])
            s = term if not s else "(%s + %s * %s)" % (term, b[k][1], s)
        return "%s[%s]" % (self.var_ref(v), s)

    # -------------------------------------------------------------------------------------------- expressions
    FUN1 = {"exp": "exp", "log": "log", "alog": "log", "log10": "log10", "alog10": "log10", "sqrt": "sqrt", "sin": "sin", "cos": "cos", "tan": "tan",
            "asin": "asin", "acos": "acos", "atan": "atan", "aint": "trunc", "dexp": "exp", "dlog": "log", "dsqrt": "sqrt"}

    def expr(self, unit, e):
        k = e[0]
        if k == "int":
            return e[1]
        if k == "real":
            t = e[1]
            if t.endswith("."):
                t += "0"
            if t.startswith("."):
                t = "0" + t
            t = t.replace(".e", ".0e")
            return t
        if k == "bool":
            return str(e[1])
        if k == "par":
            return "(%s)" % self.expr(unit, e[1])
        if k == "un":
            if e[1] == ".not.":
                return "(!%s)" % self.expr(unit, e[2])
            return "(%s%s)" % (e[1], self.expr(unit, e[2]))
        if k == "bin":
            op = e[1]
            a, b = self.expr(unit, e[2]), self.expr(unit, e[3])
            if op == "**":
                return "F_POW(%s, %s)" % (a, b)
            cop = {"/=": "!=", ".and.": "&&", ".or.": "||", ".eqv.": "==", ".neqv.": "!="}.get(op, op)
            return "(%s %s %s)" % (a, cop, b)
        if k == "name":
            v = self.p.lookup(unit, e[1])
            if v is None:
                raise TranslateError("unknown name %r" % e[1])
            if v.is_array():
                raise TranslateError("whole array %r in a scalar expression" % e[1])
            return self.scalar(unit, v)
        if k == "call":
            name, args = e[1], e[2]
            v = self.p.lookup(unit, name)
            if v is not None and v.is_array():
                return self.index(unit, v, args)
            pos = [a for a in args if a[0] != "kw"]
            if name in ("int", "ifix", "idint"):
                return "((int)(%s))" % self.expr(unit, pos[0])
            if name in ("real", "dble", "float", "dfloat"):
                return "((double)(%s))" % self.expr(unit, pos[0])
            if name == "nint":
                return "((int)lround(%s))" % self.expr(unit, pos[0])
            if name in ("min", "max", "amin1", "amax1", "dmin1", "dmax1", "min0", "max0"):
                mac = "F_MIN" if "min" in name else "F_MAX"
                s = self.expr(unit, pos[0])
                for a in pos[1:]:
                    s = "%s(%s, %s)" % (mac, s, self.expr(unit, a))
                return s
            if name in ("mod", "amod", "dmod"):
                return "F_MOD(%s, %s)" % (self.expr(unit, pos[0]), self.expr(unit, pos[1]))
            if name in ("abs", "dabs", "iabs"):
                return "F_ABS(%s)" % self.expr(unit, pos[0])
            if name in ("sign", "dsign", "isign"):
                return "F_SIGN(%s, %s)" % (self.expr(unit, pos[0]), self.expr(unit, pos[1]))
            if name in self.FUN1:
                return "%s(%s)" % (self.FUN1[name], self.expr(unit, pos[0]))
            if name == "present":
                pv = self.p.lookup(unit, pos[0][1])
                return "(%s != 0)" % pv.cname
            if name == "epsilon":
                return "DBL_EPSILON"
            if name == "tiny":
                return "DBL_MIN"
            if name == "huge":
                return "DBL_MAX"
            fn = self.p.find_sub(unit, name)
            if fn is not None and fn.fn_type:
                return "%s(%s)" % (fn.cname(), ", ".join(self.call_args(unit, fn, args)))
            raise TranslateError("unknown function or array %r" % name)
        raise TranslateError("cannot translate expression node %r" % (e,))

    # -------------------------------------------------------------------------------------------- statements
    def lvalue(self, unit, e):
        if e[0] == "name":
            v = self.p.lookup(unit, e[1])
            if v is None:
                raise TranslateError("unknown name %r" % e[1])
            return self.scalar(unit, v)
        if e[0] == "call":
            v = self.p.lookup(unit, e[1])
            if v is None or not v.is_array():
                raise TranslateError("assignment to non-array %r" % e[1])
            return self.index(unit, v, e[2])
        raise TranslateError("bad assignment target")

    def assign(self, unit, lhs, rhs, ind, out):
        # character variables (version strings) are dropped
        base = lhs[1] if lhs[0] in ("name", "call") else None
        v = self.p.lookup(unit, base) if base else None
        if v is None and rhs[0] == "str":
            return                      # version strings of modules that are not translated
        if v is None:
            raise TranslateError("unknown assignment target %r" % (base,))
        if v.is_char:
            return
        sections = None
        if v.is_array():
            if lhs[0] == "name":
                sections = [("sec", None, None, None)] * len(v.dims)
            elif any(a[0] == "sec" for a in lhs[2]):
                sections = lhs[2]
        if sections is None:
            out.append("%s%s = %s;" % (ind, self.lvalue(unit, lhs), self.expr(unit, rhs)))
            return
        # array / section assignment: loops in column-major order, rhs scalar or constructor
        b = self.var_bounds(unit, v)
        self.tmp += 1
        t = self.tmp
        loops, idx = [], []
        for k, a in enumerate(sections):
            if a[0] == "sec":
                if a[3] is not None:
                    raise TranslateError("strided section")
                lo = self.expr(unit, a[1]) if a[1] is not None else b[k][0]
                hi = self.expr(unit, a[2]) if a[2] is not None else "(%s + %s - 1)" % (b[k][0], b[k][1])
                iv = "f90_i%d_%d" % (t, k)
                loops.append((iv, lo, hi))
                idx.append(("cname", iv))
            else:
                idx.append(a)
        out.append(ind + "{")
        rhs_c = None
        if rhs[0] == "ctor":
            vals = []
            for it in rhs[1]:
                if it[0] == "ido":
                    raise TranslateError("implied do in a constructor")
                vals.append(self.expr(unit, it))
            out.append("%s  static const %s f90_v%d[] = {%s};" % (ind, v.ctype, t, ", ".join(vals)))
            out.append("%s  int f90_k%d = 0;" % (ind, t))
            rhs_c = "f90_v%d[f90_k%d++]" % (t, t)
        else:
            rhs_c = self.expr(unit, rhs)
        for (iv, lo, hi) in reversed(loops):
            out.append("%s  for (int %s = %s; %s <= %s; ++%s)" % (ind, iv, lo, iv, hi, iv))
        bb = self.var_bounds(unit, v)
        s = ""
        for k in reversed(range(len(bb))):
            a = idx[k]
            ex = a[1] if a[0] == "cname" else self.expr(unit, a)
            term = "((%s) - %s)" % (ex, bb[k][0])
            s = term if not s else "(%s + %s * %s)" % (term, bb[k][1], s)
        out.append("%s    %s[%s] = %s;" % (ind, self.var_ref(v), s, rhs_c))
        out.append(ind + "}")

    def call(self, unit, text, ind, out):
        m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", text)
        if not m:
            raise TranslateError("cannot parse call: " + text)
        name = m.group(1)
        args = [parse_arg(a) for a in split_top(m.group(2))] if m.group(2) and m.group(2).strip() else []
        sub = self.p.find_sub(unit, name)
        if sub is None:
            if args:
                raise TranslateError("call to unknown subroutine %s with arguments" % name)
            self.p.externals.add(name)
            out.append("%sf90ext__%s();" % (ind, name))
            return
        out.append("%s%s(%s);" % (ind, sub.cname(), ", ".join(self.call_args(unit, sub, args))))

    def call_args(self, unit, sub, args):
        name = sub.name
        cargs = []
        if len(args) > len(sub.args):
            raise TranslateError("too many arguments calling %s" % name)
        for k, dname in enumerate(sub.args):
            d = sub.vars.get(dname)
            if d is None:
                raise TranslateError("dummy %s of %s is not declared" % (dname, name))
            a = args[k] if k < len(args) else None
            if a is not None and a[0] == "kw":
                raise TranslateError("keyword argument in call to %s" % name)
            if a is None:
                if not d.optional:
                    raise TranslateError("missing argument %s calling %s" % (dname, name))
                cargs.append("0")
                if d.is_array():
                    cargs += ["0" for (lo, hi) in d.dims if hi is None]
                continue
            if d.is_array():
                av = self.p.lookup(unit, a[1]) if a[0] in ("name", "call") else None
                if a[0] == "name" and av is not None and av.is_array():
                    cargs.append(self.var_ref(av))
                    ab = self.var_bounds(unit, av)
                    assumed = [j for j, (lo, hi) in enumerate(d.dims) if hi is None]
                    if assumed:
                        if len(ab) != len(d.dims):
                            raise TranslateError("rank mismatch passing %s to %s of %s" % (a[1], dname, name))
                        if av.optional:
                            cargs += ["(%s ? %s : 0)" % (av.cname, ab[j][1]) for j in assumed]
                        else:
                            cargs += [ab[j][1] for j in assumed]
                elif a[0] == "call" and av is not None and av.is_array() and not any(x[0] == "sec" for x in a[2]) \
                        and all(hi is not None for (lo, hi) in d.dims):
                    cargs.append("&" + self.index(unit, av, a[2]))       # sequence association
                else:
                    raise TranslateError("unsupported actual argument for array dummy %s of %s" % (dname, name))
            elif d.byval:
                cargs.append(self.expr(unit, a))
            else:
                # by reference: an lvalue, else a temporary
                av = self.p.lookup(unit, a[1]) if a[0] in ("name", "call") else None
                if a[0] == "name" and av is not None and not av.is_array():
                    if av.where == "dummy" and not av.byval:
                        cargs.append(av.cname)
                    else:
                        cargs.append("&" + self.var_ref(av))
                elif a[0] == "call" and av is not None and av.is_array():
                    cargs.append("&" + self.index(unit, av, a[2]))
                else:
                    cargs.append("&(%s){%s}" % (d.ctype, self.expr(unit, a)))
        return cargs

    def body(self, unit, out, ind="  "):
        stack = []        # 'do' | 'if'
        for (ln, label, text) in unit.body:
            try:
                if label:
                    out.append("L%s:;" % label)
                ind_now = ind + "  " * len(stack)
                self.statement(unit, text, stack, ind_now, out)
            except TranslateError as e:
                raise TranslateError("%s:%d: %s   [%s]" % (unit.path, ln, e, text))
        if stack:
            raise TranslateError("%s: unterminated block in %s" % (unit.path, unit.name))

    def statement(self, unit, text, stack, ind, out):
        m = re.match(r"^do\s+while\s*\((.*)\)$", text)
        if m:
            out.append("%swhile (%s) {" % (ind, self.expr(unit, parse_expr(m.group(1)))))
            stack.append("do1")
            return
        m = re.match(r"^do\s+(?:(\d+)\s+)?(\w+)\s*=\s*(.*)$", text)
        if m:
            if m.group(1):
                raise TranslateError("labelled do")
            parts = split_top(m.group(3))
            var = self.lvalue(unit, ("name", m.group(2)))
            a = self.expr(unit, parse_expr(parts[0]))
            b = self.expr(unit, parse_expr(parts[1]))
            self.tmp += 1
            t = self.tmp
            if len(parts) > 2:
                ce = parse_expr(parts[2])
                try:
                    step = self.p.const_int(unit, ce)
                except TranslateError:
                    step = None
                if step is None:
                    c = self.expr(unit, ce)
                    out.append("%s{ const int f90_s%d = %s; int f90_n%d = ((%s) - (%s) + f90_s%d) / f90_s%d; for (%s = %s; f90_n%d > 0; --f90_n%d, %s += f90_s%d) {"
                               % (ind, t, c, t, b, a, t, t, var, a, t, t, var, t))
                    stack.append("do")
                    return
            else:
                step = 1
            cmp_ = "<=" if step > 0 else ">="
            out.append("%s{ const int f90_h%d = %s; for (%s = %s; %s %s f90_h%d; %s += %d) {" % (ind, t, b, var, a, var, cmp_, t, var, step))
            stack.append("do")
            return
        if text == "do":
            out.append(ind + "for (;;) {")
            stack.append("do1")
            return
        if re.match(r"^end\s*do$", text):
            k = stack.pop()
            if k not in ("do", "do1"):
                raise TranslateError("enddo closes %s" % k)
            out.append(ind[:-2] + ("}}" if k == "do" else "}"))
            return
        m = re.match(r"^if\s*\(", text)
        if m:
            j = match_paren(text, text.index("("))
            cond = self.expr(unit, parse_expr(text[text.index("(") + 1:j]))
            rest = text[j + 1:].strip()
            if rest == "then":
                out.append("%sif (%s) {" % (ind, cond))
                stack.append("if")
                return
            out.append("%sif (%s) {" % (ind, cond))
            self.statement(unit, rest, stack, ind + "  ", out)
            out.append(ind + "}")
            return
        m = re.match(r"^else\s*if\s*\(", text)
        if m:
            j = match_paren(text, text.index("("))
            cond = self.expr(unit, parse_expr(text[text.index("(") + 1:j]))
            if text[j + 1:].strip() != "then" or not stack or stack[-1] != "if":
                raise TranslateError("bad else if")
            out.append("%s} else if (%s) {" % (ind[:-2], cond))
            return
        if text == "else":
            if not stack or stack[-1] != "if":
                raise TranslateError("else without if")
            out.append(ind[:-2] + "} else {")
            return
        if re.match(r"^end\s*if$", text):
            if not stack or stack.pop() != "if":
                raise TranslateError("endif closes a do")
            out.append(ind[:-2] + "}")
            return
        m = re.match(r"^go\s*to\s+(\d+)$", text)
        if m:
            out.append("%sgoto L%s;" % (ind, m.group(1)))
            return
        if text == "continue":
            out.append(ind + ";")
            return
        if text == "return":
            out.append(ind + ("return %s;" % unit.vars[unit.name].cname if unit.fn_type else "return;"))
            return
        if text == "exit":
            out.append(ind + "break;")
            return
        if text == "cycle":
            out.append(ind + "continue;")
            return
        m = re.match(r"^stop\b\s*(.*)$", text)
        if m:
            msg = m.group(1).strip().strip("'\"")
            out.append('%sf90_stop("%s");' % (ind, msg.replace('"', "'")))
            return
        if text.startswith("call "):
            self.call(unit, text, ind, out)
            return
        # assignment
        eq = find_assign(text)
        if eq < 0:
            raise TranslateError("unsupported statement")
        lhs = parse_arg(text[:eq])
        rhs = parse_expr(text[eq + 1:])
        self.assign(unit, lhs, rhs, ind, out)

    # -------------------------------------------------------------------------------------------- units
    def signature(self, sub):
        if sub.parent is not None and sub.parent.kind == "sub":
            if sub.args:
                raise TranslateError("internal subroutine %s with arguments" % sub.name)
            return "%s %s(void)" % (sub.fn_type or "void", sub.cname())
        ps = []
        for a in sub.args:
            v = sub.vars.get(a)
            if v is None:
                raise TranslateError("%s: dummy %s of %s is not declared" % (sub.path, a, sub.name))
            if v.is_array():
                ps.append("%s *%s" % (v.ctype, v.cname))
                for k, (lo, hi) in enumerate(v.dims):
                    if hi is None:
                        ps.append("int %s_n%d" % (v.cname, k + 1))
            elif v.byval:
                ps.append("%s %s" % (v.ctype, v.cname))
            else:
                ps.append("%s *%s" % (v.ctype, v.cname))
        return "%s %s(%s)" % (sub.fn_type or "void", sub.cname(), ", ".join(ps) if ps else "void")

    def sub_def(self, sub, out, ind=""):
        out.append("")
        out.append("%s/* %s:%d  subroutine %s */" % (ind, os.path.basename(sub.path), sub.line, sub.name))
        nested = sub.parent is not None and sub.parent.kind == "sub"
        out.append(ind + ("auto " if nested else "") + self.signature(sub))
        out.append(ind + "{")
        i2 = ind + "  "
        # bounds of dummy arrays, then locals
        for name in dict.fromkeys(sub.decl_order):
            v = sub.vars[name]
            if v.is_char:
                continue
            if v.param is not None:
                if v.is_array():
                    raise TranslateError("array parameter %s" % name)
                out.append("%sconst %s %s = %s;" % (i2, v.ctype, v.cname, self.expr(sub, v.param)))
                continue
            if v.where == "dummy":
                if v.is_array():
                    for k, (lo, hi) in enumerate(v.dims):
                        l = self.expr(sub, lo) if lo is not None else "1"
                        out.append("%sconst int %s_l%d = %s;" % (i2, v.cname, k + 1, l))
                        if hi is not None:
                            out.append("%sconst int %s_n%d = (%s) - %s_l%d + 1;" % (i2, v.cname, k + 1, self.expr(sub, hi), v.cname, k + 1))
                continue
            if v.is_array():
                total = []
                for k, (lo, hi) in enumerate(v.dims):
                    if hi is None:
                        raise TranslateError("assumed-shape local %s" % name)
                    l = self.expr(sub, lo) if lo is not None else "1"
                    out.append("%sconst int %s_l%d = %s;" % (i2, v.cname, k + 1, l))
                    out.append("%sconst int %s_n%d = (%s) - %s_l%d + 1;" % (i2, v.cname, k + 1, self.expr(sub, hi), v.cname, k + 1))
                    total.append("%s_n%d" % (v.cname, k + 1))
                sz = " * ".join("(long)" + t for t in total)
                out.append("%s%s %s[(%s) > 0 ? (%s) : 1];" % (i2, v.ctype, v.cname, sz, sz))
                out.append("%smemset(%s, 0, sizeof %s);" % (i2, v.cname, v.cname))
            else:
                init = self.expr(sub, v.init) if v.init is not None else "0"
                out.append("%s%s %s = %s;" % (i2, v.ctype, v.cname, init))
        for (dn, vals) in sub.data:
            v = sub.vars[dn]
            cv = ", ".join(self.expr(sub, x) for x in vals)
            if v.is_array():
                out.append("%s{ static const %s f90_d[] = {%s}; memcpy(%s, f90_d, sizeof f90_d); }" % (i2, v.ctype, cv, v.cname))
            else:
                out.append("%s%s = %s;" % (i2, v.cname, cv))
        for s in sub.sub_order:
            out.append("%sauto %s;" % (i2, self.signature(s)))
        body = []
        self.body(sub, body, i2)
        out += body
        out.append(i2 + ("return %s;" % sub.vars[sub.name].cname if sub.fn_type else "return;"))     # also keeps a trailing label legal
        for s in sub.sub_order:
            self.sub_def(s, out, i2)
        out.append(ind + "}")


def find_assign(text: str) -> int:
    depth, q = 0, None
    for i, c in enumerate(text):
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
        elif c == "=" and depth == 0:
            if text[i + 1:i + 2] == "=" or text[i - 1] in "<>/=":
                continue
            return i
    return -1


def parse_arg(s: str):
    p = Parser(tokenize(s))
    e = p.arg()
    if p.peek()[0] != "end":
        raise TranslateError("trailing tokens in %r" % s)
    return e


# ------------------------------------------------------------------------------------------------ driver
def translate(paths, outdir):
    prog = Program()
    for p in paths:
        prog.parse_file(p)
    em = Emitter(prog)
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "f90ref_rt.h"), "w") as f:
        f.write(PRELUDE)
    # apply equivalences: the first name becomes an alias of the second
    for m in prog.order:
        for (a, b) in m.equiv:
            va, vb = m.vars[a[1]], m.vars[b[1]]
            va.alias_of = vb
    hdr = ["/* generated by tools/f90_to_c.py -- module variables and subroutine prototypes */", "#ifndef F90REF_H", "#define F90REF_H",
           '#include "f90ref_rt.h"',
           ]
    glob = ['#include "f90ref.h"', ""]
    reg = []
    for m in prog.order:
        hdr.append("/* module %s (%s) */" % (m.name, os.path.basename(m.path)))
        for name in dict.fromkeys(m.decl_order):
            v = m.vars[name]
            if v.is_char:
                continue
            cn = "%s__%s" % (m.name, name)
            if v.param is not None:
                if v.is_array():
                    raise TranslateError("%s: array parameter %s" % (m.path, name))
                if v.ctype == "int":
                    try:
                        hdr.append("enum { %s = %d };" % (cn, prog.const_int(m, ("name", name))))
                    except TranslateError:
                        hdr.append("/* %s: kind parameter, not needed */" % cn)
                else:
                    hdr.append("static const double %s = %s;" % (cn, em.expr(m, v.param)))
                continue
            if v.is_array():
                n = 1
                for (lo, hi) in v.dims:
                    l = prog.const_int(m, lo) if lo is not None else 1
                    n *= prog.const_int(m, hi) - l + 1
                if v.alias_of is not None:
                    hdr.append("#define %s %s__%s" % (cn, m.name, v.alias_of.name))
                else:
                    hdr.append("extern %s %s[%d];" % (v.ctype, cn, n))
                    glob.append("%s %s[%d];" % (v.ctype, cn, n))
                    reg.append('  {"%s.%s", %s, %d, %d},' % (m.name, name, cn, n, 1 if v.ctype == "int" else 0))
            else:
                hdr.append("extern %s %s;" % (v.ctype, cn))
                init = " = " + em.expr(m, v.init) if v.init is not None else ""
                glob.append("%s %s%s;" % (v.ctype, cn, init))
                reg.append('  {"%s.%s", &%s, 1, %d},' % (m.name, name, cn, 1 if v.ctype == "int" else 0))
        for s in m.sub_order:
            hdr.append(em.signature(s) + ";")
    files = {}
    for m in prog.order:
        for s in m.sub_order:
            base = os.path.basename(s.path)
            cfile = base[:-4].replace(".", "_") + ".c"
            out = files.setdefault(cfile, ['#include "f90ref.h"', "/* translated from %s */" % s.path])
            em.sub_def(s, out)
    for ext in sorted(prog.externals):
        hdr.append("void f90ext__%s(void);   /* not among the translated sources: supplied by the harness */" % ext)
    hdr += ["struct f90_var { const char *name; void *ptr; long n; int is_int; };", "extern const struct f90_var f90_vars[];", "#endif"]
    glob += ["", "const struct f90_var f90_vars[] = {"] + reg + ["  {0, 0, 0, 0}", "};"]
    with open(os.path.join(outdir, "f90ref.h"), "w") as f:
        f.write("\n".join(hdr) + "\n")
    with open(os.path.join(outdir, "f90ref_globals.c"), "w") as f:
        f.write("\n".join(glob) + "\n")
    for name, lines in files.items():
        with open(os.path.join(outdir, name), "w") as f:
            f.write("\n".join(lines) + "\n")
    return prog, sorted(files)


def main(argv):
    if len(argv) < 4 or argv[1] != "-o":
        sys.stderr.write(__doc__)
        return 2
    try:
        prog, files = translate(argv[3:], argv[2])
    except TranslateError as e:
        sys.stderr.write("f90_to_c: %s\n" % e)
        return 1
    nsub = sum(len(m.sub_order) + sum(len(s.sub_order) for s in m.sub_order) for m in prog.order)
    print("f90_to_c: %d modules, %d subroutines -> %s (%s)" % (len(prog.order), nsub, argv[2], ", ".join(files)))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))

"""f90cxx.py — a translator from the Fortran 90 subset the Noah-MP physics is written in to C++ (TEST INFRASTRUCTURE).

Why: the build image has no Fortran compiler, so the reference cannot be compiled as it is.  This script reads the
reference's *.F90 files WHERE THEY LIE (/root/reference/phys, never copied into the repository), translates them
statement by statement into C++ that keeps the Fortran evaluation rules (operator precedence and association, integer
division, x**n by repeated multiplication as libgcc's __powisf2, MIN/MAX as gfortran expands them, DO trip counts
fixed at loop entry, array lower bounds as declared, by-reference arguments), and writes the result under
oracle/_ref/ (git-ignored).  Compiled with g++ -O2 -ffp-contract=off it plays the part `gfortran -O2` would play:
an executable form of the reference's own text, made without anybody re-typing a formula.  tests/test_reference_pin.py
compares the hand-written oracle with it bit for bit.

Nothing here is physics; the translator knows the language, not the model.

usage: python oracle/ref/f90cxx.py OUT.cpp FILE.F90[:skip=SUB1,SUB2][:only=...] ...
"""
import re
import sys

# ------------------------------------------------------------------------------------------------ source reading

OPS_DOT = ("AND", "OR", "NOT", "EQV", "NEQV", "EQ", "NE", "LT", "LE", "GT", "GE", "TRUE", "FALSE")


def cpp_filter(lines, defined=()):
    """#ifdef / #ifndef / #if / #else / #endif with every macro undefined unless listed."""
    out, stack = [], []
    for ln in lines:
        s = ln.strip()
        if s.startswith("#"):
            d = s[1:].strip()
            if d.startswith("ifdef"):
                stack.append(d.split()[1] in defined)
            elif d.startswith("ifndef"):
                stack.append(d.split()[1] not in defined)
            elif d.startswith("if"):
                e = d[2:]
                e = re.sub(r"defined\s*\(\s*(\w+)\s*\)", lambda m: "1" if m.group(1) in defined else "0", e)
                e = re.sub(r"[A-Za-z_]\w*", lambda m: "1" if m.group(0) in defined else "0", e)
                e = e.replace("&&", " and ").replace("||", " or ").replace("!", " not ")
                stack.append(bool(eval(e)))
            elif d.startswith("else"):
                stack[-1] = not stack[-1]
            elif d.startswith("endif"):
                stack.pop()
            else:
                pass  # #define / #include: none that matter
            out.append("")
            continue
        out.append(ln if all(stack) else "")
    return out


def strip_comment(line):
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


def upcase(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
                out.append(ch)
            else:
                out.append(ch.upper())
    return "".join(out)


def logical_lines(path, defined=()):
    """[(first source line number, statement text upper-cased outside strings)]"""
    raw = open(path, errors="replace").read().split("\n")
    raw = cpp_filter(raw, defined)
    res, cur, start = [], "", 0
    for no, ln in enumerate(raw, 1):
        ln = strip_comment(ln.replace("\t", " ")).rstrip()
        if not ln.strip():
            continue
        s = ln.strip()
        if cur:
            if s.startswith("&"):
                s = s[1:]
            cur += " " + s
        else:
            cur, start = s, no
        if cur.endswith("&"):
            cur = cur[:-1]
            continue
        for part in split_semicolon(cur):
            if part.strip():
                res.append((start, upcase(part.strip())))
        cur = ""
    return res


def split_semicolon(s):
    parts, q, cur = [], None, ""
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == ";":
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


# ------------------------------------------------------------------------------------------------------ tokens

TOK = re.compile(r"""
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<dotop>\.(?:AND|OR|NOT|EQV|NEQV|EQ|NE|LT|LE|GT|GE|TRUE|FALSE)\.)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ED][+-]?\d+)?(?:_\w+)?)
  | (?P<id>[A-Z_]\w*)
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|\(/|/\)|[-+*/(),:=<>%])
  | (?P<ws>\s+)
""", re.X)


def tokenize(s):
    toks, i = [], 0
    while i < len(s):
        m = TOK.match(s, i)
        if not m:
            raise SyntaxError("cannot tokenize %r at %r" % (s, s[i:i + 20]))
        k = m.lastgroup
        t = m.group(k)
        if k == "num":
            # "1.EQ." : the dot belongs to the operator;  "1.E5" is a number
            m2 = re.match(r"(\d+)\.(?=(?:AND|OR|NOT|EQV|NEQV|EQ|NE|LT|LE|GT|GE)\.)", s[i:])
            if m2:
                t = m2.group(1)
                toks.append(("num", t))
                i += len(t)
                continue
        if k != "ws":
            toks.append((k, t))
        i = m.end()
    return toks


# --------------------------------------------------------------------------------------------------------- AST


class Node:
    pass


class Num(Node):
    def __init__(self, text):
        t = text
        kind = None
        if "_" in t:
            t, kind = t.split("_", 1)
        self.text = t
        if re.fullmatch(r"\d+", t):
            self.ty = "int"
        elif "D" in t or kind in ("8", "DP", "R8"):
            self.ty = "double"
        else:
            self.ty = "real"


class Str(Node):
    ty = "char"

    def __init__(self, text):
        self.text = text


class Log(Node):
    ty = "logical"

    def __init__(self, v):
        self.v = v


class Name(Node):
    def __init__(self, name):
        self.name = name


class Rng(Node):
    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi


class Ref(Node):
    """NAME(args): array element / section / function call / intrinsic; args may be Rng or ('kw', name, expr)"""

    def __init__(self, name, args):
        self.name, self.args = name, args


class Bin(Node):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b


class Un(Node):
    def __init__(self, op, a):
        self.op, self.a = op, a


class ArrCons(Node):
    def __init__(self, items):
        self.items = items


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        x = self.peek()
        self.i += 1
        return x

    def accept(self, text):
        if self.peek()[1] == text and self.peek()[0] != "str":
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            raise SyntaxError("expected %r, found %r in %r" % (text, self.peek(), self.t))

    def at_end(self):
        return self.i >= len(self.t)

    # precedence climbing, Fortran levels
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        a = self.p_or()
        while self.peek()[1] in (".EQV.", ".NEQV."):
            op = self.next()[1]
            a = Bin(op, a, self.p_or())
        return a

    def p_or(self):
        a = self.p_and()
        while self.peek()[1] == ".OR.":
            self.next()
            a = Bin(".OR.", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek()[1] == ".AND.":
            self.next()
            a = Bin(".AND.", a, self.p_not())
        return a

    def p_not(self):
        if self.peek()[1] == ".NOT.":
            self.next()
            return Un(".NOT.", self.p_not())
        return self.p_rel()

    REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=", ".EQ.": "==", ".NE.": "!=", ".LT.": "<",
           ".LE.": "<=", ".GT.": ">", ".GE.": ">="}

    def p_rel(self):
        a = self.p_concat()
        if self.peek()[1] in self.REL and self.peek()[0] in ("op", "dotop"):
            op = self.REL[self.next()[1]]
            a = Bin(op, a, self.p_concat())
        return a

    def p_concat(self):
        a = self.p_add()
        while self.peek() == ("op", "//"):
            self.next()
            b = self.p_add()
            if isinstance(a, Str) and isinstance(b, Str):
                a = Str(a.text[:-1] + b.text[1:])
            else:
                a = Bin("//", a, b)
        return a

    def p_add(self):
        if self.peek()[0] == "op" and self.peek()[1] in "+-":
            op = self.next()[1]
            a = Un(op, self.p_mul())
        else:
            a = self.p_mul()
        while self.peek()[0] == "op" and self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = Bin(op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek()[0] == "op" and self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            a = Bin(op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_prim()
        if self.peek() == ("op", "**"):
            self.next()
            # right associative; the exponent may carry a sign:  X**-2 is an extension, X**(-2) the norm
            if self.peek()[0] == "op" and self.peek()[1] in "+-":
                op = self.next()[1]
                b = Un(op, self.p_pow())
            else:
                b = self.p_pow()
            a = Bin("**", a, b)
        return a

    def p_prim(self):
        k, t = self.next()
        if k == "num":
            return Num(t)
        if k == "str":
            return Str(t)
        if k == "dotop":
            if t == ".TRUE.":
                return Log(True)
            if t == ".FALSE.":
                return Log(False)
            raise SyntaxError("unexpected " + t)
        if k == "op" and t == "(":
            e = self.expr()
            self.expect(")")
            return Un("()", e)
        if k == "op" and t == "(/":
            items = [self.expr()]
            while self.accept(","):
                items.append(self.expr())
            self.expect("/)")
            return ArrCons(items)
        if k == "id":
            if self.peek() == ("op", "("):
                self.next()
                args = self.arglist()
                return Ref(t, args)
            return Name(t)
        raise SyntaxError("unexpected token %r in %r" % ((k, t), self.t))

    def arglist(self):
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.arg())
            if self.accept(")"):
                return args
            self.expect(",")

    def arg(self):
        # keyword argument
        if self.peek()[0] == "id" and self.peek(1) == ("op", "=") and self.peek(2) != ("op", "="):
            kw = self.next()[1]
            self.next()
            return ("kw", kw, self.expr())
        lo = None
        if self.peek() != ("op", ":"):
            lo = self.expr()
        if self.accept(":"):
            hi = None
            if self.peek()[1] not in (",", ")") or self.peek()[0] != "op":
                hi = self.expr()
            return Rng(lo, hi)
        return lo


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.at_end():
        raise SyntaxError("trailing tokens in %r" % s)
    return e


# ----------------------------------------------------------------------------------------------------- symbols

CTYPE = {"int": "int", "real": "float", "double": "double", "logical": "bool", "char": "const char*"}


class Sym:
    def __init__(self, name, ty, dims=None, intent=None, param=None, dummy=False, init=None, optional=False,
                 save=False):
        self.name, self.ty, self.dims, self.intent = name, ty, dims, intent
        self.param, self.dummy, self.init, self.optional, self.save = param, dummy, init, optional, save
        self.owner = None  # Module for module-level entities

    @property
    def rank(self):
        return len(self.dims) if self.dims else 0


class Scope:
    def __init__(self, name, kind, parent=None):
        self.name, self.kind, self.parent = name, kind, parent
        self.syms = {}      # declared here
        self.order = []     # declaration order
        self.uses = []      # (module name, only-list or None)
        self.subs = {}      # contained procedures
        self.args = []
        self.body = []      # [(lineno, label, text)]
        self.data = []
        self.data_stmts = []
        self.stmt_funcs = {}
        self.skip = False

    def declare(self, s):
        if s.name in self.syms:
            old = self.syms[s.name]
            # attributes spread over several statements
            if s.dims and not old.dims:
                old.dims = s.dims
            if s.ty and old.ty is None:
                old.ty = s.ty
            return old
        self.syms[s.name] = s
        self.order.append(s.name)
        return s


class Program:
    def __init__(self):
        self.modules = {}
        self.all_subs = {}  # NAME -> Scope (module procedures, for call signatures)
        self.skipped = set()
        self.noop = set()  # procedures whose calls are dropped (table readers: the harness fills the tables)


TYPE_RE = re.compile(r"^(INTEGER|REAL|DOUBLE\s*PRECISION|LOGICAL|CHARACTER)\b")


def split_top(s, sep=","):
    parts, depth, cur, q = [], 0, "", None
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == "(":
            depth += 1
            cur += ch
        elif ch == ")":
            depth -= 1
            cur += ch
        elif ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
        i += 1
    parts.append(cur)
    return [p.strip() for p in parts]


def match_paren(s, i):
    """index of the ')' matching the '(' at s[i]"""
    depth, q = 0, None
    for j in range(i, len(s)):
        ch = s[j]
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return j
    raise SyntaxError("unbalanced parentheses in %r" % s)


def parse_dims(text):
    dims = []
    for d in split_top(text):
        parts = split_top(d, ":")
        if d in (":", "*"):
            dims.append((None, None))
        elif len(parts) == 2:
            lo, hi = parts
            dims.append((parse_expr(lo) if lo else None, parse_expr(hi) if hi and hi != "*" else None))
        else:
            dims.append((Num("1"), parse_expr(d)))
    return dims


def split_colon_top(s):
    return split_top(s, ":")


def parse_decl(stmt, scope):
    """type declaration statement -> symbols declared into scope; returns True if it was one"""
    m = TYPE_RE.match(stmt)
    if not m:
        return False
    base = re.sub(r"\s+", "", m.group(1))
    rest = stmt[m.end():].lstrip()
    ty = {"INTEGER": "int", "REAL": "real", "DOUBLEPRECISION": "double", "LOGICAL": "logical", "CHARACTER": "char"}[base]
    # kind / len selector
    if rest.startswith("("):
        j = match_paren(rest, 0)
        sel = rest[1:j].replace(" ", "")
        rest = rest[j + 1:].lstrip()
        if ty == "real" and re.search(r"(KIND=)?8$", sel):
            ty = "double"
    elif rest.startswith("*"):
        m2 = re.match(r"\*\s*(\d+|\(\s*\*\s*\))", rest)
        if ty == "real" and m2.group(1) == "8":
            ty = "double"
        rest = rest[m2.end():].lstrip()
    attrs = {}
    if "::" in rest:
        a, ents = rest.split("::", 1)
        for at in split_top(a):
            if not at:
                continue
            u = at.replace(" ", "")
            if u.startswith("DIMENSION"):
                attrs["dims"] = parse_dims(at[at.index("(") + 1:match_paren(at, at.index("("))])
            elif u.startswith("INTENT"):
                attrs["intent"] = u[7:-1]
            elif u == "PARAMETER":
                attrs["param"] = True
            elif u == "OPTIONAL":
                attrs["optional"] = True
            elif u == "SAVE":
                attrs["save"] = True
            elif u in ("ALLOCATABLE", "TARGET", "PUBLIC", "PRIVATE", "EXTERNAL"):
                attrs[u.lower()] = True
            else:
                raise SyntaxError("attribute %r in %r" % (at, stmt))
    else:
        if rest.startswith(","):
            raise SyntaxError("attribute list without '::' in %r" % stmt)
        ents = rest
    for ent in split_top(ents):
        if not ent:
            continue
        init = None
        if "=" in ent:
            # NAME[(dims)] = init   (the '=' is at depth 0)
            pp = split_top(ent, "=")
            if len(pp) >= 2:
                ent, init = pp[0].strip(), "=".join(pp[1:]).strip()
        mm = re.match(r"^([A-Z_]\w*)\s*(\(.*\))?\s*(\*\s*\d+)?$", ent)
        if not mm:
            raise SyntaxError("entity %r in %r" % (ent, stmt))
        name = mm.group(1)
        dims = attrs.get("dims")
        if mm.group(2):
            dims = parse_dims(mm.group(2)[1:-1])
        s = Sym(name, ty, dims, attrs.get("intent"), None, False, None, attrs.get("optional", False),
                attrs.get("save", False))
        if init is not None:
            e = parse_expr(init)
            if attrs.get("param"):
                s.param = e
            else:
                s.init, s.save = e, True
        old = scope.syms.get(name)
        if old is not None:
            old.ty = ty
            if dims:
                old.dims = dims
            if s.intent:
                old.intent = s.intent
            if s.param is not None:
                old.param = s.param
            if s.init is not None:
                old.init, old.save = s.init, True
            old.optional |= s.optional
        else:
            scope.declare(s)
    return True


# ------------------------------------------------------------------------------------------------ file parsing

SUB_RE = re.compile(r"^(?:RECURSIVE\s+)?SUBROUTINE\s+([A-Z_]\w*)\s*(?:\((.*)\))?\s*$")
END_RE = re.compile(r"^END\s*(SUBROUTINE|MODULE|FUNCTION|PROGRAM)?\b\s*([A-Z_]\w*)?\s*$")
LABEL_RE = re.compile(r"^(\d+)\s+(.*)$")
DECL_START = ("IMPLICIT", "USE ", "NAMELIST", "SAVE", "DATA ", "DATA(", "PARAMETER", "DIMENSION", "EXTERNAL", "INTRINSIC",
              "PUBLIC", "PRIVATE", "INTERFACE", "COMMON", "EQUIVALENCE")


def parse_file(prog, path, skip=(), only=None, defined=()):
    lines = logical_lines(path, defined)
    stack = []  # scopes
    in_interface = 0
    in_type = False
    for no, st in lines:
        if in_interface:
            if re.match(r"^END\s*INTERFACE", st):
                in_interface -= 1
            continue
        if re.match(r"^INTERFACE\b", st):
            in_interface += 1
            continue
        if in_type:
            if re.match(r"^END\s*TYPE", st):
                in_type = False
            continue
        if re.match(r"^TYPE\s+[A-Z_]\w*\s*$", st) or re.match(r"^TYPE\s*,", st):
            in_type = True
            continue
        m = re.match(r"^MODULE\s+([A-Z_]\w*)\s*$", st)
        if m and not st.startswith("MODULE PROCEDURE"):
            mod = Scope(m.group(1), "module")
            mod.path = path
            prog.modules[mod.name] = mod
            stack = [mod]
            continue
        m = SUB_RE.match(st)
        if m:
            parent = stack[-1] if stack else None
            sub = Scope(m.group(1), "sub", parent)
            sub.args = [a.strip() for a in split_top(m.group(2) or "") if a.strip()]
            sub.line = no
            sub.path = path
            for a in sub.args:
                sub.declare(Sym(a, None, dummy=True))
            if parent is not None:
                parent.subs[sub.name] = sub
            if parent is None or parent.kind == "module":
                sub.skip = (sub.name in skip) or (only is not None and sub.name not in only)
                if not sub.skip:
                    prog.all_subs[sub.name] = sub
                else:
                    prog.skipped.add(sub.name)
            stack.append(sub)
            continue
        m = END_RE.match(st)
        if m and (m.group(1) or st.strip() == "END"):
            if m.group(1) in (None, "SUBROUTINE", "MODULE", "FUNCTION", "PROGRAM"):
                if stack:
                    stack.pop()
                continue
        if st == "CONTAINS":
            continue
        if not stack:
            continue
        sc = stack[-1]
        if sc.kind == "sub" and sc.skip_chain():
            continue
        label = None
        ml = LABEL_RE.match(st)
        if ml:
            label, st = ml.group(1), ml.group(2)
        sc.body.append((no, label, st))
    return prog


def _skip_chain(self):
    s = self
    while s is not None:
        if s.skip:
            return True
        s = s.parent
    return False


Scope.skip_chain = _skip_chain


def _split_decls_impl(scope):
    """separate the specification part from the executable part; fill the symbol table"""
    body = []
    for no, label, st in scope.body:
        if not body and label is None:
            if st.startswith("USE "):
                m = re.match(r"^USE\s+([A-Z_]\w*)\s*(?:,\s*ONLY\s*:\s*(.*))?$", st)
                only = None
                if m.group(2) is not None:
                    only = [x.strip() for x in m.group(2).split(",") if x.strip()]
                scope.uses.append((m.group(1), only))
                continue
            if st.startswith("IMPLICIT") or st.startswith("NAMELIST") or st.startswith("EXTERNAL") or \
               st.startswith("INTRINSIC") or st in ("SAVE", "PUBLIC", "PRIVATE") or st.startswith("PUBLIC ") or \
               st.startswith("PRIVATE "):
                continue
            if st.startswith("SAVE "):
                for n in st[5:].replace("::", "").split(","):
                    n = n.strip()
                    if n in scope.syms:
                        scope.syms[n].save = True
                    else:
                        scope.declare(Sym(n, None, save=True))
                continue
            if re.match(r"^DATA\b", st):
                scope.data.append(st[4:].strip())
                continue
            if re.match(r"^PARAMETER\s*\(", st):
                inner = st[st.index("(") + 1:match_paren(st, st.index("("))]
                for p in split_top(inner):
                    n, v = p.split("=", 1)
                    scope.syms[n.strip()].param = parse_expr(v.strip())
                continue
            if re.match(r"^DIMENSION\b", st):
                for ent in split_top(st[9:].replace("::", "")):
                    mm = re.match(r"^([A-Z_]\w*)\s*\((.*)\)$", ent)
                    s = scope.syms.get(mm.group(1)) or scope.declare(Sym(mm.group(1), None))
                    s.dims = parse_dims(mm.group(2))
                continue
            try:
                if parse_decl(st, scope):
                    continue
            except SyntaxError as e:
                raise SyntaxError("%s:%d: %s" % (getattr(scope, "path", "?"), no, e))
        body.append((no, label, st))
    # statement functions:  F(X) = expr  at the head of the executable part, F a declared scalar
    while body:
        no, label, st = body[0]
        m = re.match(r"^([A-Z_]\w*)\s*\(([A-Z_0-9, ]*)\)\s*=(?!=)(.*)$", st)
        if not m or m.group(1) not in scope.syms or scope.syms[m.group(1)].dims or scope.syms[m.group(1)].dummy:
            break
        scope.stmt_funcs[m.group(1)] = ([a.strip() for a in m.group(2).split(",") if a.strip()], m.group(3).strip())
        body.pop(0)
    scope.body = body
    # implicit typing never applies (IMPLICIT NONE everywhere); a dummy without a type is an error we want to see
    for d in scope.data:
        apply_data(scope, d)


def const_int(scope, text):
    """value of an integer constant expression made of literals and PARAMETERs"""
    text = text.strip()
    if re.fullmatch(r"[+-]?\d+", text):
        return int(text)
    sc = scope
    while sc is not None:
        if text in sc.syms and sc.syms[text].param is not None and isinstance(sc.syms[text].param, Num):
            return int(sc.syms[text].param.text)
        sc = sc.parent
    raise SyntaxError("not a constant integer: " + text)


def apply_data(scope, d):
    """DATA A /v/ ; DATA A /v1, v2/ (array, storage order) ; DATA (A(I,2),I=1,N) /.../ ; DATA A, B /1., 2./
    -> scope.data_stmts: [(C-ish target, value expression text)] with target either ('elem', fortran text) or
    ('lin', NAME, k)"""
    segs = re.findall(r"\s*,?\s*(.*?)/(.*?)/", d)
    for names, vals in segs:
        vals = [v.strip() for v in split_top(vals)]
        expanded = []
        for v in vals:
            if re.match(r"^\d+\s*\*", v):
                n, x = v.split("*", 1)
                expanded += [x.strip()] * int(n)
            else:
                expanded.append(v)
        names = names.strip()
        targets = []
        m = re.match(r"^\(\s*([A-Z_]\w*\s*\(.*\))\s*,\s*([A-Z_]\w*)\s*=\s*([^,]+),\s*([^,]+)\)$", names)
        if m:
            lo, hi = const_int(scope, m.group(3)), const_int(scope, m.group(4))
            for k in range(lo, hi + 1):
                targets.append(("elem", re.sub(r"\b%s\b" % m.group(2), str(k), m.group(1))))
            scope.syms[m.group(1).split("(")[0].strip()].save = True
        else:
            for tname in split_top(names):
                sy = scope.syms[tname]
                sy.save = True
                if sy.dims:
                    if len(split_top(names)) != 1:
                        raise SyntaxError("DATA with several arrays: " + d)
                    targets += [("lin", tname, k) for k in range(len(expanded))]
                else:
                    targets.append(("elem", tname))
        if len(targets) != len(expanded):
            raise SyntaxError("DATA count mismatch: " + d)
        scope.data_stmts += list(zip(targets, expanded))


# ---------------------------------------------------------------------------------------------------- emission

INTRINSIC_ELEMENTAL = {
    "ABS": "F_ABS", "SQRT": "F_SQRT", "EXP": "F_EXP", "LOG": "F_LOG", "ALOG": "F_LOG", "LOG10": "F_LOG10",
    "ALOG10": "F_LOG10", "SIN": "F_SIN", "COS": "F_COS", "TAN": "F_TAN", "ATAN": "F_ATAN", "ASIN": "F_ASIN",
    "ACOS": "F_ACOS", "TANH": "F_TANH", "SINH": "F_SINH", "COSH": "F_COSH", "ATAN2": "F_ATAN2", "SIGN": "F_SIGN",
    "MOD": "F_MOD", "NINT": "F_NINT", "INT": "F_INT", "IFIX": "F_INT", "REAL": "F_REAL", "FLOAT": "F_REAL",
    "DBLE": "F_DBLE", "AINT": "F_AINT", "ANINT": "F_ANINT", "ISNAN": "F_ISNAN", "CEILING": "F_CEILING", "FLOOR": "F_FLOOR",
    "DABS": "F_ABS", "DSQRT": "F_SQRT", "DEXP": "F_EXP", "DLOG": "F_LOG", "IABS": "F_ABS", "AMOD": "F_MOD",
}
MINMAX = {"MAX": "F_MAX", "MIN": "F_MIN", "AMAX1": "F_MAX", "AMIN1": "F_MIN", "MAX0": "F_MAX", "MIN0": "F_MIN",
          "DMAX1": "F_MAX", "DMIN1": "F_MIN"}
REDUCTIONS = ("MAXVAL", "MINVAL", "SUM", "ANY", "ALL", "COUNT")
DROP_CALLS = ("WRF_MESSAGE", "WRF_DEBUG", "WRF_DEBUG2", "FLUSH")
FATAL_CALLS = ("WRF_ERROR_FATAL", "WRF_ERROR_FATAL3", "ABORT")


class Emitter:
    def __init__(self, prog):
        self.prog = prog
        self.out = []
        self.uid = 0
        self.warnings = []
        self.files = []
        self.markers = []  # (file id * 100000 + line, enclosing module procedure) of every executable statement

    # ---- name resolution -----------------------------------------------------------------------------------
    def module_public(self, mname, seen=None):
        """{name: (Sym, owner module)} visible through USE mname (its own entities and what it re-exports)"""
        seen = seen or set()
        if mname in seen or mname not in self.prog.modules:
            return {}
        seen.add(mname)
        mod = self.prog.modules[mname]
        res = {}
        for um, only in mod.uses:
            for k, v in self.module_public(um, seen).items():
                if only is None or k in only:
                    res[k] = v
        for k, s in mod.syms.items():
            res[k] = (s, mod)
        return res

    def lookup(self, scope, name):
        """-> (Sym, qualifier string) or (None, None)"""
        s = scope
        while s is not None:
            if name in s.syms:
                return s.syms[name], ("M_%s::" % s.name if s.kind == "module" else "")
            for um, only in s.uses:
                if only is not None and name not in only:
                    continue
                pub = self.module_public(um)
                if name in pub:
                    sym, mod = pub[name]
                    return sym, "M_%s::" % mod.name
            s = s.parent
        return None, None

    @staticmethod
    def stmt_func(scope, name):
        s = scope
        while s is not None:
            if name in s.stmt_funcs:
                return s.stmt_funcs[name]
            s = s.parent
        return None

    def module_subs(self, mname, seen=None):
        seen = seen if seen is not None else set()
        if mname in seen or mname not in self.prog.modules:
            return {}
        seen.add(mname)
        mod = self.prog.modules[mname]
        res = {}
        for um, only in mod.uses:
            for k, v in self.module_subs(um, seen).items():
                if only is None or k in only:
                    res[k] = v
        for k, v in mod.subs.items():
            if not v.skip:
                res[k] = v
        return res

    def find_sub(self, scope, name):
        """-> (Scope of the callee, C++ qualifier) or (None, None)"""
        s = scope
        while s is not None:
            if name in s.subs and not s.subs[name].skip:
                return s.subs[name], ("M_%s::" % s.name if s.kind == "module" else "")
            for um, only in s.uses:
                if only is not None and name not in only:
                    continue
                ms = self.module_subs(um)
                if name in ms:
                    return ms[name], "M_%s::" % ms[name].parent.name
            s = s.parent
        return None, None

    # ---- expression typing -----------------------------------------------------------------------------------
    def typeof(self, e, sc):
        if isinstance(e, (Num, Str, Log)):
            return e.ty
        if isinstance(e, Name):
            s, _ = self.lookup(sc, e.name)
            if s is None:
                raise NameError("undeclared %s in %s" % (e.name, sc.name))
            return s.ty
        if isinstance(e, Ref):
            s, _ = self.lookup(sc, e.name)
            if s is not None and (s.dims or s.ty == "char"):
                return s.ty
            if self.stmt_func(sc, e.name):
                return s.ty
            n = e.name
            if n in MINMAX:
                return self.promote([self.typeof(a, sc) for a in e.args])
            if n in ("REAL", "FLOAT", "SNGL"):
                return "real"
            if n in ("DBLE",):
                return "double"
            if n in ("INT", "NINT", "IFIX", "CEILING", "FLOOR", "SIZE", "COUNT", "MAX0", "MIN0", "IABS", "LEN", "LEN_TRIM"):
                return "int"
            if n in ("ANY", "ALL", "PRESENT", "ISNAN"):
                return "logical"
            if n in ("MAXVAL", "MINVAL", "SUM"):
                return self.typeof(e.args[0], sc)
            if n in INTRINSIC_ELEMENTAL:
                if n in ("SIGN", "MOD", "ATAN2", "AMOD"):
                    return self.promote([self.typeof(a, sc) for a in e.args])
                return self.typeof(e.args[0], sc)
            if n in ("TRIM", "ADJUSTL"):
                return "char"
            if n in ("EPSILON", "TINY", "HUGE"):
                return self.typeof(e.args[0], sc)
            raise NameError("unknown function or array %s in %s" % (n, sc.name))
        if isinstance(e, Un):
            if e.op == ".NOT.":
                return "logical"
            return self.typeof(e.a, sc)
        if isinstance(e, Bin):
            if e.op in (".AND.", ".OR.", ".EQV.", ".NEQV.", "==", "!=", "<", "<=", ">", ">="):
                return "logical"
            if e.op == "//":
                return "char"
            ta, tb = self.typeof(e.a, sc), self.typeof(e.b, sc)
            if e.op == "**":
                return ta if tb == "int" else self.promote([ta, tb])
            return self.promote([ta, tb])
        if isinstance(e, ArrCons):
            return self.typeof(e.items[0], sc)
        if isinstance(e, Rng):
            return "int"
        raise TypeError(e)

    @staticmethod
    def promote(ts):
        if "double" in ts:
            return "double"
        if "real" in ts:
            return "real"
        if all(t == "int" for t in ts):
            return "int"
        if all(t == "logical" for t in ts):
            return "logical"
        raise TypeError("cannot promote %r" % (ts,))

    # ---- array-valuedness ---------------------------------------------------------------------------------------
    def shape_of(self, e, sc):
        """None for a scalar expression, else a list of (lo C++ text, extent C++ text) per array dimension, taken from
        the first array-valued operand (Fortran requires conformance)"""
        if isinstance(e, Name):
            s, _ = self.lookup(sc, e.name)
            if s is not None and s.dims:
                return [self.dim_lo_ext(s, d, sc) for d in range(s.rank)]
            return None
        if isinstance(e, Ref):
            s, _ = self.lookup(sc, e.name)
            if s is not None and s.dims:
                sh = []
                for d, a in enumerate(e.args):
                    if isinstance(a, Rng):
                        lo = self.ex(a.lo, sc) if a.lo is not None else self.dim_lo_ext(s, d, sc)[0]
                        if a.hi is not None:
                            hi = self.ex(a.hi, sc)
                        else:
                            l0, n0 = self.dim_lo_ext(s, d, sc)
                            hi = "((%s)+(%s)-1)" % (l0, n0)
                        sh.append((lo, "((%s)-(%s)+1)" % (hi, lo)))
                return sh or None
            if e.name in REDUCTIONS or e.name in ("SIZE",):
                return None
            for a in e.args:
                if isinstance(a, tuple):
                    a = a[2]
                sh = self.shape_of(a, sc)
                if sh:
                    return sh
            return None
        if isinstance(e, Un):
            return self.shape_of(e.a, sc)
        if isinstance(e, Bin):
            return self.shape_of(e.a, sc) or self.shape_of(e.b, sc)
        if isinstance(e, ArrCons):
            return [("1", str(len(e.items)))]
        return None

    def dim_lo_ext(self, s, d, sc):
        lo, hi = s.dims[d]
        lo_t = self.ex(lo, sc) if lo is not None else "1"
        if hi is None:
            return lo_t, "0x7fffffff"
        return lo_t, "((%s)-(%s)+1)" % (self.ex(hi, sc), lo_t)

    # ---- expressions --------------------------------------------------------------------------------------------
    def num(self, e):
        t = e.text
        if e.ty == "int":
            return t
        if e.ty == "double":
            t = t.replace("D", "E")
            if "." not in t and "E" not in t:
                t += ".0"
            return t
        if "." not in t and "E" not in t:
            t += ".0"
        return t + "f"

    def ex(self, e, sc, elem=None):
        """C++ text of expression e; elem = list of loop counter names when e is evaluated element-wise"""
        if isinstance(e, Num):
            return self.num(e)
        if isinstance(e, Str):
            body = e.text[1:-1].replace("\\", "\\\\").replace('"', '\\"')
            return '"%s"' % body
        if isinstance(e, Log):
            return "true" if e.v else "false"
        if isinstance(e, Name):
            s, q = self.lookup(sc, e.name)
            if s is None:
                raise NameError("undeclared %s in %s" % (e.name, sc.name))
            if s.dims and elem is not None:
                idx = ["((%s)+%s)" % (self.dim_lo_ext(s, d, sc)[0], elem[d]) for d in range(s.rank)]
                return "%s%s(%s)" % (q, s.name, ",".join(idx))
            return q + s.name
        if isinstance(e, Un):
            if e.op == "()":
                return "(" + self.ex(e.a, sc, elem) + ")"
            if e.op == ".NOT.":
                return "(!" + self.ex(e.a, sc, elem) + ")"
            return "(%s%s)" % (e.op, self.ex(e.a, sc, elem))
        if isinstance(e, Bin):
            if e.op == "//":
                return '""'
            a, b = self.ex(e.a, sc, elem), self.ex(e.b, sc, elem)
            if e.op == "**":
                ta, tb = self.typeof(e.a, sc), self.typeof(e.b, sc)
                if tb == "int":
                    if ta == "int":
                        return "F_IPOW(%s,%s)" % (a, b)
                    return "F_POWI(%s,%s)" % (a, b)
                t = CTYPE[self.promote([ta, tb])]
                return "F_POW((%s)(%s),(%s)(%s))" % (t, a, t, b)
            op = {".AND.": "&&", ".OR.": "||", ".EQV.": "==", ".NEQV.": "!="}.get(e.op, e.op)
            return "(%s %s %s)" % (a, op, b)
        if isinstance(e, Ref):
            return self.ref(e, sc, elem)
        if isinstance(e, ArrCons):
            if elem is not None:
                t = CTYPE[self.typeof(e, sc)]
                return "((const %s[]){%s})[%s]" % (t, ",".join(self.ex(i, sc) for i in e.items), elem[0])
            raise SyntaxError("array constructor outside an array assignment")
        raise TypeError(e)

    def ref(self, e, sc, elem):
        s, q = self.lookup(sc, e.name)
        if s is not None and s.dims:
            if len(e.args) != s.rank:
                raise SyntaxError("rank mismatch for %s in %s" % (e.name, sc.name))
            idx, k = [], 0
            for a in e.args:
                if isinstance(a, Rng):
                    if elem is None:
                        raise SyntaxError("array section of %s used as a scalar in %s" % (e.name, sc.name))
                    d = len(idx)
                    lo = self.ex(a.lo, sc) if a.lo is not None else self.dim_lo_ext(s, d, sc)[0]
                    idx.append("((%s)+%s)" % (lo, elem[k]))
                    k += 1
                else:
                    idx.append(self.ex(a, sc))  # subscripts are scalar
            return "%s%s(%s)" % (q, s.name, ",".join(idx))
        n = e.name
        args = [a for a in e.args]
        if self.stmt_func(sc, n):
            return "SF_%s(%s)" % (n, ", ".join(self.ex(a, sc, elem) for a in args))
        if n in MINMAX:
            f = MINMAX[n]
            t = CTYPE[self.promote([self.typeof(a, sc) for a in args])]
            parts = ["(%s)(%s)" % (t, self.ex(a, sc, elem)) for a in args]
            acc = parts[0]
            for p in parts[1:]:
                acc = "%s(%s,%s)" % (f, acc, p)
            return acc
        if n in ("REAL", "FLOAT", "SNGL"):
            if len(args) == 2:
                kind = args[1][2] if isinstance(args[1], tuple) else args[1]
                if self.ex(kind, sc) == "8":
                    return "((double)(%s))" % self.ex(args[0], sc, elem)
            return "((float)(%s))" % self.ex(args[0], sc, elem)
        if n == "DBLE":
            return "((double)(%s))" % self.ex(args[0], sc, elem)
        if n in ("INT", "IFIX"):
            return "((int)(%s))" % self.ex(args[0], sc, elem)
        if n in INTRINSIC_ELEMENTAL:
            if n in ("SIGN", "MOD", "ATAN2", "AMOD"):
                t = CTYPE[self.promote([self.typeof(a, sc) for a in args])]
                return "%s(%s)" % (INTRINSIC_ELEMENTAL[n], ",".join("(%s)(%s)" % (t, self.ex(a, sc, elem)) for a in args))
            return "%s(%s)" % (INTRINSIC_ELEMENTAL[n], ",".join(self.ex(a, sc, elem) for a in args))
        if n in REDUCTIONS:
            return self.reduction(e, sc)
        if n == "PRESENT":
            s2, _ = self.lookup(sc, args[0].name)
            return "(%s%s != nullptr)" % (args[0].name, "__p" if not s2.dims else ".p")
        if n == "SIZE":
            s2, _ = self.lookup(sc, args[0].name)
            if len(args) == 2:
                d = int(self.ex(args[1], sc)) - 1
                return self.dim_lo_ext(s2, d, sc)[1]
            return "*".join(self.dim_lo_ext(s2, d, sc)[1] for d in range(s2.rank))
        if n in ("EPSILON", "TINY", "HUGE"):
            t = CTYPE[self.typeof(args[0], sc)]
            return {"EPSILON": "std::numeric_limits<%s>::epsilon()", "TINY": "std::numeric_limits<%s>::min()",
                    "HUGE": "std::numeric_limits<%s>::max()"}[n] % t
        if n in ("TRIM", "ADJUSTL"):
            return self.ex(args[0], sc, elem)
        raise NameError("unknown function or array %s in %s" % (n, sc.name))

    def reduction(self, e, sc):
        arg = e.args[0]
        sh = self.shape_of(arg, sc)
        if not sh:
            raise SyntaxError("reduction over a scalar in " + sc.name)
        self.uid += 1
        ks = ["_r%d_%d" % (self.uid, d) for d in range(len(sh))]
        body = self.ex(arg, sc, ks)
        t = CTYPE[self.typeof(arg, sc)] if e.name in ("MAXVAL", "MINVAL", "SUM") else "bool"
        loops = "".join("for (int %s = 0; %s < %s; ++%s) " % (k, k, n, k) for k, (_, n) in zip(ks, sh))
        if e.name == "MAXVAL":
            return "[&]{ %s _m = -__builtin_inff(); %s{ %s _v = %s; if (_v > _m) _m = _v; } return _m; }()" % (t, loops, t, body)
        if e.name == "MINVAL":
            return "[&]{ %s _m = __builtin_inff(); %s{ %s _v = %s; if (_v < _m) _m = _v; } return _m; }()" % (t, loops, t, body)
        if e.name == "SUM":
            return "[&]{ %s _m = 0; %s{ _m = _m + %s; } return _m; }()" % (t, loops, body)
        if e.name == "ANY":
            return "[&]{ %s{ if (%s) return true; } return false; }()" % (loops, body)
        if e.name == "ALL":
            return "[&]{ %s{ if (!(%s)) return false; } return true; }()" % (loops, body)
        if e.name == "COUNT":
            return "[&]{ int _m = 0; %s{ if (%s) ++_m; } return _m; }()" % (loops, body)

    # ---- statements ---------------------------------------------------------------------------------------------
    def w(self, ind, text):
        self.out.append("  " * ind + text)

    def emit_assign(self, lhs_txt, rhs_txt, sc, ind):
        lhs, rhs = parse_expr(lhs_txt), parse_expr(rhs_txt)
        lsym = None
        if isinstance(lhs, (Name, Ref)):
            lsym, _ = self.lookup(sc, lhs.name)
        if lsym is None:
            raise NameError("assignment to undeclared %s in %s" % (lhs_txt, sc.name))
        if lsym.ty == "char":
            self.w(ind, "/* character assignment dropped */;")
            return
        sh = self.shape_of(lhs, sc)
        if sh is None:
            self.w(ind, "%s = %s;" % (self.ex(lhs, sc), self.ex(rhs, sc)))
            return
        self.uid += 1
        ks = ["_k%d_%d" % (self.uid, d) for d in range(len(sh))]
        rsh = self.shape_of(rhs, sc)
        rhs_c = self.ex(rhs, sc, ks if rsh is not None else None)
        lhs_c = self.ex(lhs, sc, ks)
        # Fortran evaluates the whole right-hand side before storing; the statements of this code base never overlap
        # (checked: when the same array appears on both sides, it does with the same subscripts)
        for d in reversed(range(len(sh))):
            k, n = ks[d], sh[d][1]
            self.w(ind, "for (int %s = 0, %s_n = %s; %s < %s_n; ++%s)" % (k, k, n, k, k, k))
            ind += 1
        self.w(ind, "%s = %s;" % (lhs_c, rhs_c))

    def call_args(self, callee, actuals, sc):
        """C++ actual argument list for CALL callee(actuals)"""
        pos, kw = [], {}
        for a in actuals:
            if isinstance(a, tuple):
                kw[a[1]] = a[2]
            else:
                pos.append(a)
        res = []
        for i, dname in enumerate(callee.args):
            d = callee.syms[dname]
            a = pos[i] if i < len(pos) else kw.get(dname)
            if a is None:
                if not d.optional:
                    raise SyntaxError("missing argument %s in call to %s from %s" % (dname, callee.name, sc.name))
                res.append("nullptr" if d.dims else "F_ABSENT<%s>()" % CTYPE[d.ty])
                continue
            res.append(self.one_arg(d, a, sc, callee))
        if len(pos) > len(callee.args):
            raise SyntaxError("too many arguments in call to %s from %s" % (callee.name, sc.name))
        return res

    def one_arg(self, d, a, sc, callee):
        t = CTYPE[d.ty]
        if d.ty == "char":
            return '""' if not isinstance(a, Str) else self.ex(a, sc)
        asym = None
        if isinstance(a, (Name, Ref)):
            asym, q = self.lookup(sc, a.name)
        if d.dims:
            # array dummy: pass the address of the first element
            if asym is None or not asym.dims:
                raise SyntaxError("scalar passed to array dummy %s of %s from %s" % (d.name, callee.name, sc.name))
            if asym.ty != d.ty:
                raise SyntaxError("type mismatch for array dummy %s of %s from %s" % (d.name, callee.name, sc.name))
            if isinstance(a, Name):
                return "%s%s.p" % (q, a.name)
            first = []
            nrng = 0
            for dd, x in enumerate(a.args):
                if isinstance(x, Rng):
                    nrng += 1
                    first.append(self.ex(x.lo, sc) if x.lo is not None else self.dim_lo_ext(asym, dd, sc)[0])
                else:
                    first.append(self.ex(x, sc))
            # a section is passed without a copy only when it is contiguous: ranges in the leading dimensions
            lead = [isinstance(x, Rng) for x in a.args]
            if nrng and any(lead[i] and not all(lead[:i]) for i in range(len(lead))):
                if not (nrng == 1 and asym.rank == 1):
                    raise SyntaxError("non-contiguous section %s passed to %s from %s" % (a.name, callee.name, sc.name))
            return "&%s%s(%s)" % (q, a.name, ",".join(first))
        # scalar dummy
        if asym is not None and not (asym.dims and isinstance(a, Name)):
            is_elem = isinstance(a, Ref) and asym.dims and not any(isinstance(x, Rng) for x in a.args)
            is_scalar_var = isinstance(a, Name) and not asym.dims and asym.param is None
            if (is_elem or is_scalar_var) and asym.ty == d.ty:
                if is_scalar_var and asym.optional and asym.dummy:
                    return "%s__p ? *%s__p : F_ABSENT<%s>()" % (a.name, a.name, t) if d.optional else self.ex(a, sc)
                return self.ex(a, sc)
        if d.intent in ("OUT", "INOUT"):
            self.warnings.append("expression passed to INTENT(%s) dummy %s of %s from %s" % (d.intent, d.name, callee.name, sc.name))
        return "F_TMP<%s>(%s)" % (t, self.ex(a, sc))

    def emit_body(self, sc, ind):
        """translate sc.body; returns nothing, appends to self.out"""
        stack = []  # open constructs: ('if',) ('do', name, has_named_jump) ...
        for no, label, st in sc.body:
            try:
                if label is not None:
                    self.w(ind, "L%s: ;" % label)
                if not re.match(r"^(ELSE|END|CASE)", st) and not (stack and stack[-1][0] == "where"):
                    self.markers.append((self.file_id(sc) * 100000 + no, top_name(sc)))
                    self.w(ind, "REF_LINE(%d);" % self.markers[-1][0])
                ind = self.emit_stmt(st, sc, ind, stack)
            except (SyntaxError, NameError, TypeError, KeyError, AttributeError, IndexError) as e:
                raise type(e)("%s:%d [%s] %s\n    statement: %s" % (getattr(sc, "path", "?"), no, sc.name, e, st))
        if stack:
            raise SyntaxError("unclosed construct in %s: %r" % (sc.name, stack))

    def emit_stmt(self, st, sc, ind, stack):
        # construct name prefix  NAME: DO ...
        cname = None
        m = re.match(r"^([A-Z_]\w*)\s*:\s*(DO\b.*|IF\s*\(.*THEN)$", st)
        if m and not m.group(1) in ("ELSE",):
            cname, st = m.group(1), m.group(2)
        if st == "CONTINUE":
            self.w(ind, ";")
            return ind
        if re.match(r"^(WRITE|PRINT|READ|OPEN|CLOSE|REWIND|FORMAT)\b", st) and not re.match(r"^(WRITE|PRINT|READ|OPEN|CLOSE|REWIND|FORMAT)\s*=", st):
            self.w(ind, "/* i/o dropped */;")
            return ind
        if st == "RETURN":
            self.w(ind, "return;")
            return ind
        if re.match(r"^STOP\b", st):
            self.w(ind, 'F_FATAL("STOP");')
            return ind
        m = re.match(r"^GO\s*TO\s+(\d+)$", st)
        if m:
            self.w(ind, "goto L%s;" % m.group(1))
            return ind
        # IF
        if re.match(r"^IF\s*\(", st):
            j = match_paren(st, st.index("("))
            cond, rest = st[st.index("(") + 1:j], st[j + 1:].strip()
            c = self.ex(parse_expr(cond), sc)
            if rest == "THEN":
                self.w(ind, "if (%s) {" % c)
                stack.append(("if", cname))
                return ind + 1
            self.w(ind, "if (%s) {" % c)
            ind2 = self.emit_stmt(rest, sc, ind + 1, stack)
            assert ind2 == ind + 1, "block statement after a logical IF"
            self.w(ind, "}")
            return ind
        m = re.match(r"^ELSE\s*IF\s*\(", st)
        if m:
            j = match_paren(st, st.index("("))
            c = self.ex(parse_expr(st[st.index("(") + 1:j]), sc)
            self.w(ind - 1, "} else if (%s) {" % c)
            return ind
        if re.match(r"^ELSE(\s+[A-Z_]\w*)?$", st):
            self.w(ind - 1, "} else {")
            return ind
        if re.match(r"^END\s*IF(\s+[A-Z_]\w*)?$", st):
            k = stack.pop()
            assert k[0] == "if", "ENDIF closes %r" % (k,)
            self.w(ind - 1, "}")
            return ind - 1
        # DO
        m = re.match(r"^DO\s+([A-Z_]\w*)\s*=\s*(.*)$", st)
        if m and not re.match(r"^DO\s+WHILE\b", st):
            var = m.group(1)
            parts = split_top(m.group(2))
            vs, q = self.lookup(sc, var)
            v = q + var
            self.uid += 1
            u = self.uid
            a = self.ex(parse_expr(parts[0]), sc)
            b = self.ex(parse_expr(parts[1]), sc)
            self.w(ind, "{")
            if len(parts) == 3:
                c = self.ex(parse_expr(parts[2]), sc)
                self.w(ind + 1, "const int _e%d = %s, _s%d = %s;" % (u, b, u, c))
                self.w(ind + 1, "for (%s = %s; _s%d > 0 ? %s <= _e%d : %s >= _e%d; %s += _s%d) {" % (v, a, u, v, u, v, u, v, u))
            else:
                self.w(ind + 1, "const int _e%d = %s;" % (u, b))
                self.w(ind + 1, "for (%s = %s; %s <= _e%d; ++%s) {" % (v, a, v, u, v))
            stack.append(("do", cname, u))
            return ind + 2
        m = re.match(r"^DO\s+WHILE\s*\((.*)\)$", st)
        if m:
            self.uid += 1
            self.w(ind, "{")
            self.w(ind + 1, "while (%s) {" % self.ex(parse_expr(m.group(1)), sc))
            stack.append(("do", cname, self.uid))
            return ind + 2
        if st == "DO":
            self.uid += 1
            self.w(ind, "{")
            self.w(ind + 1, "for (;;) {")
            stack.append(("do", cname, self.uid))
            return ind + 2
        if re.match(r"^END\s*DO(\s+[A-Z_]\w*)?$", st):
            k = stack.pop()
            assert k[0] == "do", "ENDDO closes %r" % (k,)
            self.w(ind - 1, "_c%d: ;" % k[2])
            self.w(ind - 1, "}")
            self.w(ind - 2, "}")
            self.w(ind - 2, "_x%d: ;" % k[2])
            return ind - 2
        m = re.match(r"^(EXIT|CYCLE)(?:\s+([A-Z_]\w*))?$", st)
        if m:
            target = None
            for k in reversed(stack):
                if k[0] == "do" and (m.group(2) is None or k[1] == m.group(2)):
                    target = k
                    break
            if target is None:
                raise SyntaxError("%s outside a loop" % st)
            self.w(ind, "goto %s%d;" % ("_x" if m.group(1) == "EXIT" else "_c", target[2]))
            return ind
        # SELECT CASE
        m = re.match(r"^SELECT\s*CASE\s*\((.*)\)$", st)
        if m:
            self.w(ind, "switch (%s) {" % self.ex(parse_expr(m.group(1)), sc))
            stack.append(("select", cname, False))
            return ind + 1
        m = re.match(r"^CASE\s*\((.*)\)$", st)
        if m:
            k = stack[-1]
            if k[2]:
                self.w(ind, "break;")
            stack[-1] = ("select", k[1], True)
            for v in split_top(m.group(1)):
                self.w(ind - 1, "case %s:" % self.ex(parse_expr(v), sc))
            return ind
        if re.match(r"^CASE\s+DEFAULT$", st):
            k = stack[-1]
            if k[2]:
                self.w(ind, "break;")
            stack[-1] = ("select", k[1], True)
            self.w(ind - 1, "default:")
            return ind
        if re.match(r"^END\s*SELECT", st):
            stack.pop()
            self.w(ind, "break;")
            self.w(ind - 1, "}")
            return ind - 1
        # WHERE (mask) array = expr   (single-statement form) and the block form
        m = re.match(r"^WHERE\s*\(", st)
        if m:
            j = match_paren(st, st.index("("))
            mask, rest = st[st.index("(") + 1:j], st[j + 1:].strip()
            if rest:
                self.emit_where(mask, [rest], sc, ind)
                return ind
            stack.append(["where", mask, [], [], False])
            return ind
        if stack and stack[-1][0] == "where":
            k = stack[-1]
            if re.match(r"^END\s*WHERE", st):
                stack.pop()
                self.emit_where(k[1], k[2], sc, ind, k[3])
                return ind
            if re.match(r"^ELSE\s*WHERE$", st):
                k[4] = True
                return ind
            (k[3] if k[4] else k[2]).append(st)
            return ind
        # CALL
        m = re.match(r"^CALL\s+([A-Z_]\w*)\s*(\(.*\))?$", st)
        if m:
            name = m.group(1)
            if name in DROP_CALLS:
                self.w(ind, "/* call %s dropped */;" % name.lower())
                return ind
            if name in FATAL_CALLS:
                msg = '""'
                if m.group(2):
                    a = Parser(tokenize(m.group(2)))
                    a.next()
                    args = a.arglist()
                    strs = [x for x in args if isinstance(x, Str)]
                    if strs:
                        msg = self.ex(strs[-1], sc)
                self.w(ind, "F_FATAL(%s);" % msg)
                return ind
            if name in self.prog.noop:
                self.w(ind, "/* call %s dropped: the harness fills what it reads */;" % name.lower())
                return ind
            callee, cq = self.find_sub(sc, name)
            if callee is None and name in self.prog.skipped:
                self.w(ind, 'F_FATAL("call to %s, which was left out of the translation");' % name)
                return ind
            if callee is None:
                raise NameError("call to unknown subroutine %s from %s" % (name, sc.name))
            actuals = []
            if m.group(2):
                a = Parser(tokenize(m.group(2)))
                a.next()
                actuals = a.arglist()
            self.w(ind, "%sS_%s(%s);" % (cq, name, ", ".join(self.call_args(callee, actuals, sc))))
            return ind
        # assignment
        eq = find_assign_eq(st)
        if eq is not None:
            self.emit_assign(st[:eq].strip(), st[eq + 1:].strip(), sc, ind)
            return ind
        raise SyntaxError("statement not understood")

    def emit_where(self, mask, stmts, sc, ind, else_stmts=()):
        me = parse_expr(mask)
        sh = self.shape_of(me, sc)
        self.uid += 1
        ks = ["_k%d_%d" % (self.uid, d) for d in range(len(sh))]
        for d in reversed(range(len(sh))):
            self.w(ind, "for (int %s = 0; %s < %s; ++%s)" % (ks[d], ks[d], sh[d][1], ks[d]))
            ind += 1
        self.w(ind, "if (%s) {" % self.ex(me, sc, ks))
        for branch, lst in ((0, stmts), (1, else_stmts)):
            if branch and else_stmts:
                self.w(ind, "} else {")
            for s_ in lst:
                eq = find_assign_eq(s_)
                lhs, rhs = parse_expr(s_[:eq].strip()), parse_expr(s_[eq + 1:].strip())
                rsh = self.shape_of(rhs, sc)
                self.w(ind + 1, "%s = %s;" % (self.ex(lhs, sc, ks), self.ex(rhs, sc, ks if rsh else None)))
        self.w(ind, "}")

    def emit_data(self, sc, ind):
        """DATA statements: executed once (static storage)"""
        if not sc.data_stmts:
            return
        self.uid += 1
        self.w(ind, "static const int _data%d = [%s]{" % (self.uid, "" if sc.kind == "module" else "&"))
        for tgt, val in sc.data_stmts:
            if tgt[0] == "lin":
                s_, q = self.lookup(sc, tgt[1])
                self.w(ind + 1, "%s%s__s[%d] = %s;" % (q, tgt[1], tgt[2], self.cast(parse_expr(val), s_.ty, sc)))
            else:
                e = parse_expr(tgt[1])
                s_, q = self.lookup(sc, e.name)
                self.w(ind + 1, "%s = %s;" % (self.ex(e, sc), self.cast(parse_expr(val), s_.ty, sc)))
        self.w(ind + 1, "return 0; }();")
        self.w(ind, "(void)_data%d;" % self.uid if sc.kind != "module" else "")

    def file_id(self, sc):
        path = getattr(sc, "path", None)
        while path is None and sc.parent is not None:
            sc = sc.parent
            path = getattr(sc, "path", None)
        if path not in self.files:
            self.files.append(path)
        return self.files.index(path)

    def emit_registry(self, order):
        names = sorted(set(n for _, n in self.markers))
        self.w(0, "static const int ref_marker_table[] = {%s};" % ", ".join("%d,%d" % (m, names.index(n)) for m, n in self.markers))
        self.w(0, 'extern "C" const int* ref_markers(int* n) { *n = %d; return ref_marker_table; }' % len(self.markers))
        self.w(0, 'extern "C" const char* ref_marker_names() { return "%s"; }' % ";".join(names))
        self.w(0, 'extern "C" const char* ref_files() { return "%s"; }' % ";".join(str(x) for x in self.files))
        """name -> storage table of the module variables, and an untyped entry point per module procedure"""
        self.w(0, "struct RefVar { const char* name; void* p; unsigned long bytes; char type; };")
        self.w(0, "static const RefVar ref_var_table[] = {")
        for m in order:
            mod = self.prog.modules[m]
            for n in mod.order:
                s = mod.syms[n]
                if s.ty in (None, "char") or s.param is not None:
                    continue
                tl = {"int": "i", "real": "r", "double": "d", "logical": "l"}[s.ty]
                if s.dims:
                    self.w(1, '{"%s.%s", (void*)M_%s::%s__s, sizeof(M_%s::%s__s), \'%s\'},' % (m, n, m, n, m, n, tl))
                else:
                    self.w(1, '{"%s.%s", (void*)&M_%s::%s, sizeof(M_%s::%s), \'%s\'},' % (m, n, m, n, m, n, tl))
        self.w(1, "{nullptr, nullptr, 0, 0}};")
        self.w(0, 'extern "C" const RefVar* ref_vars() { return ref_var_table; }')
        seen = set()
        for m in order:
            mod = self.prog.modules[m]
            for sub in mod.subs.values():
                if sub.skip:
                    continue
                ename = sub.name if sub.name not in seen else "%s__%s" % (m, sub.name)
                seen.add(ename)
                args, sig = [], []
                for i, a in enumerate(sub.args):
                    sy = sub.syms[a]
                    t = CTYPE[sy.ty]
                    tl = {"int": "i", "real": "r", "double": "d", "logical": "l", "char": "c"}[sy.ty]
                    if sy.ty == "char":
                        args.append("(const char*)a[%d]" % i)
                    elif sy.dims:
                        args.append("(%s*)a[%d]" % (t, i))
                        tl = tl.upper() + str(sy.rank)
                    elif sy.optional:
                        args.append("a[%d] ? *(%s*)a[%d] : F_ABSENT<%s>()" % (i, t, i, t))
                    else:
                        args.append("*(%s*)a[%d]" % (t, i))
                    sig.append("%s:%s" % (a, tl))
                self.w(0, 'extern "C" int ref_call_%s(void** a, char* msg, int nmsg) {' % ename)
                self.w(1, "try { M_%s::S_%s(%s); return 0; }" % (m, sub.name, ", ".join(args)))
                self.w(1, "catch (const std::exception& e) { if (msg && nmsg > 0) { std::strncpy(msg, e.what(), nmsg - 1); msg[nmsg - 1] = 0; } return 1; }")
                self.w(0, "}")
                self.w(0, 'extern "C" const char* ref_sig_%s() { return "%s"; }' % (ename, ",".join(sig)))

    # ---- declarations ----------------------------------------------------------------------------------------
    def view_type(self, s):
        return "FV%d<%s>" % (s.rank, CTYPE[s.ty])

    def view_args(self, s, sc):
        parts = []
        for d in range(s.rank):
            lo, n = self.dim_lo_ext(s, d, sc)
            parts.append(lo)
            if d < s.rank - 1:
                parts.append(n)
        return ", ".join(parts)

    def total_size(self, s, sc):
        return "*".join("(%s)" % self.dim_lo_ext(s, d, sc)[1] for d in range(s.rank))

    def emit_local(self, s, sc, ind, static=False):
        t = CTYPE[s.ty]
        pre = "static " if static or s.save else ""
        if s.ty == "char":
            return
        if s.param is not None:
            if s.dims:
                items = s.param.items if isinstance(s.param, ArrCons) else None
                if items is None:
                    raise SyntaxError("array PARAMETER %s without a constructor" % s.name)
                self.w(ind, "static %s %s__s[] = {%s};" % (t, s.name, ", ".join(self.cast(i, s.ty, sc) for i in items)))
                self.w(ind, "static const %s %s{%s__s, %s};" % (self.view_type(s), s.name, s.name, self.view_args(s, sc)))
            else:
                self.w(ind, "%sconst %s %s = %s;" % ("static " if static else "", t, s.name, self.cast(s.param, s.ty, sc)))
            return
        if s.dims:
            if s.init is not None:
                items = s.init.items
                self.w(ind, "static %s %s__s[%s] = {%s};" % (t, s.name, self.total_size(s, sc),
                                                             ", ".join(self.cast(i, s.ty, sc) for i in items)))
            elif pre:
                self.w(ind, "static %s %s__s[%s];" % (t, s.name, self.total_size(s, sc)))
            else:
                self.w(ind, "%s %s__s[%s]; F_ZERO(%s__s, %s);" % (t, s.name, self.total_size(s, sc), s.name, self.total_size(s, sc)))
            self.w(ind, "%s%s %s{%s__s, %s};" % (pre, self.view_type(s), s.name, s.name, self.view_args(s, sc)))
            return
        if s.init is not None:
            self.w(ind, "static %s %s = %s;" % (t, s.name, self.cast(s.init, s.ty, sc)))
        else:
            self.w(ind, "%s%s %s = %s;" % (pre, t, s.name, "false" if s.ty == "logical" else "0"))

    def cast(self, e, ty, sc):
        return "(%s)(%s)" % (CTYPE[ty], self.ex(e, sc))

    def signature(self, sub):
        parts = []
        for a in sub.args:
            s = sub.syms[a]
            if s.ty is None:
                raise SyntaxError("dummy %s of %s has no type" % (a, sub.name))
            if s.ty == "char":
                parts.append("const char* %s" % a)
            elif s.dims:
                parts.append("%s* %s__p" % (CTYPE[s.ty], a))
            elif s.optional:
                parts.append("%s& %s__r" % (CTYPE[s.ty], a))
            else:
                parts.append("%s& %s" % (CTYPE[s.ty], a))
        return "(" + ", ".join(parts) + ")"

    def emit_sub(self, sub, ind, as_lambda=False):
        split_decls(sub)
        if as_lambda:
            self.w(ind, "auto S_%s = [&]%s -> void {" % (sub.name, self.signature(sub)))
        else:
            self.w(ind, "void S_%s%s {" % (sub.name, self.signature(sub)))
        ind += 1
        # dummies first (array views may need the scalar dummies), then locals in declaration order
        for a in sub.args:
            s = sub.syms[a]
            if s.dims:
                self.w(ind, "const %s %s{%s__p, %s};" % (self.view_type(s), a, a, self.view_args(s, sub)))
            elif s.optional and s.ty != "char":
                self.w(ind, "%s* %s__p = F_IS_ABSENT(%s__r) ? nullptr : &%s__r; %s& %s = %s__r;" % (CTYPE[s.ty], a, a, a, CTYPE[s.ty], a, a))
        for n in sub.order:
            s = sub.syms[n]
            if s.dummy:
                continue
            if s.ty is None:
                raise SyntaxError("%s in %s has no type" % (n, sub.name))
            self.emit_local(s, sub, ind)
        self.emit_data(sub, ind)
        for fn, (fargs, ftext) in sub.stmt_funcs.items():
            sig = ", ".join("%s %s" % (CTYPE[sub.syms[a].ty], a) for a in fargs)
            self.w(ind, "auto SF_%s = [&](%s) -> %s { return %s; };" % (fn, sig, CTYPE[sub.syms[fn].ty], self.ex(parse_expr(ftext), sub)))
        for isub in sub.subs.values():
            if not isub.skip:
                self.emit_sub(isub, ind, as_lambda=True)
        self.emit_body(sub, ind)
        ind -= 1
        self.w(ind, "};" if as_lambda else "}")
        self.w(ind, "")

    def emit_module(self, mod):
        split_decls(mod)
        self.w(0, "namespace M_%s {" % mod.name)
        for n in mod.order:
            s = mod.syms[n]
            if s.ty is None:
                continue
            self.emit_local(s, mod, 0, static=True)
        self.emit_data(mod, 0)
        subs = [s for s in mod.subs.values() if not s.skip]
        # prototypes need the dummies' types: parse the specification parts first
        for s in subs:
            split_decls(s)
        for s in subs:
            self.w(0, "void S_%s%s;" % (s.name, self.signature(s)))
        self.w(0, "}  // namespace M_%s" % mod.name)
        self.w(0, "")

    def emit_module_subs(self, mod):
        self.w(0, "namespace M_%s {" % mod.name)
        for s in mod.subs.values():
            if not s.skip:
                self.emit_sub(s, 0)
        self.w(0, "}  // namespace M_%s" % mod.name)
        self.w(0, "")


def split_decls(scope):
    """specification part -> symbol table, once per scope (and its contained procedures)"""
    if getattr(scope, "_split", False):
        return
    scope._split = True
    _split_decls_impl(scope)
    for isub in scope.subs.values():
        if not isub.skip:
            split_decls(isub)


def top_name(sc):
    while sc.parent is not None and sc.parent.kind != "module":
        sc = sc.parent
    return sc.name


def find_assign_eq(st):
    """index of the assignment '=' at parenthesis depth 0 (not ==, /=, <=, >=, =>), or None"""
    depth, q = 0, None
    for i, ch in enumerate(st):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if st[i + 1:i + 2] == "=" or st[i - 1:i] in ("=", "/", "<", ">") or st[i + 1:i + 2] == ">":
                continue
            return i
    return None


PRELUDE = r"""// GENERATED by oracle/ref/f90cxx.py from the reference's Fortran sources — do not edit, do not commit.
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include "ref_prelude.h"
"""


def module_order(prog):
    """modules in dependency order"""
    done, order = set(), []

    def visit(m):
        if m in done or m not in prog.modules:
            return
        done.add(m)
        mod = prog.modules[m]
        split_decls(mod)
        for um, _ in mod.uses:
            visit(um)
        for s in mod.subs.values():
            if not s.skip:
                split_decls(s)
                for um, _ in s.uses:
                    visit(um)
        order.append(m)

    for m in prog.modules:
        visit(m)
    return order


def main(argv):
    out = argv[1]
    prog = Program()
    for spec in argv[2:]:
        parts = spec.split(":")
        skip, only = (), None
        for p in parts[1:]:
            if p.startswith("skip="):
                skip = tuple(p[5:].upper().split(","))
            elif p.startswith("only="):
                only = tuple(p[5:].upper().split(","))
            elif p.startswith("noop="):
                prog.noop.update(p[5:].upper().split(","))
        parse_file(prog, parts[0], skip, only)
    em = Emitter(prog)
    em.out.append(PRELUDE)
    order = module_order(prog)
    for m in order:
        em.emit_module(prog.modules[m])
    for m in order:
        em.emit_module_subs(prog.modules[m])
    em.emit_registry(order)
    open(out, "w").write("\n".join(em.out) + "\n")
    for wmsg in sorted(set(em.warnings)):
        print("warning:", wmsg, file=sys.stderr)
    print("wrote %s: %d modules, %d subroutines, %d lines" % (out, len(order), len(prog.all_subs), len(em.out)))


if __name__ == "__main__":
    main(sys.argv)

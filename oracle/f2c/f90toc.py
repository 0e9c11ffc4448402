#!/usr/bin/env python
"""f90toc.py -- a small Fortran-90 -> C transpiler for the subset of the language
the reference's hot-path routines are written in (TEST INFRASTRUCTURE).

Why.  The reference is Fortran and this image has no Fortran compiler, so the
reference cannot be built here.  This tool reads the reference's OWN source
files where they lie under /root/reference/src (nothing is copied into the
repository), translates them statement by statement into C, and
oracle/f2c/Makefile compiles the result into oracle/_ref/libflexref.so.  The
tests then run the reference's advance / initialize / interpol_* / hanna* /
cbl / conccalc / ... code on the same inputs as the hand-written oracle and
compare bit for bit (tests/test_ref_transpiled.py).

What it handles: free-form source, modules with module variables, subroutines
and functions, explicit-shape arrays with arbitrary lower bounds, parameters,
SAVEd / initialised locals, do loops, block and one-line if, goto / labelled
statements, exit / cycle, call, return, stop, and expressions with the Fortran
operators and the intrinsics these files use.  Typing is left to the C compiler:
every variable is declared with its Fortran type (real -> float, real(dp) ->
double, integer -> int), real literals get the matching suffix, and generic
intrinsics / the ** operator dispatch on the operand types with C11 _Generic
(oracle/f2c/f2c_rt.h).  With -O2 -ffp-contract=off the arithmetic is then the
IEEE arithmetic gfortran -O2 produces for the same statements; transcendental
functions go through oracle/fpo_math.h (correctly rounded), the same definition
the oracle and the device's strict mode use.

par_mod's integer parameters (nxmax, nymax, nzmax, maxpart, maxspec, maxnests,
...) become run-time variables so that a test can shrink the compiled-in
extents; every module array is allocated by ref_alloc() from those values.
Anything outside the subset raises an error instead of being guessed.
"""
import re
import sys

DOTOPS = ("and", "or", "not", "eq", "ne", "lt", "le", "gt", "ge", "eqv", "neqv", "true", "false")
CTYPE = {"real": "float", "double": "double", "integer": "int", "logical": "int", "int8": "signed char",
         "int16": "short", "int64": "long long"}
# allocatable module arrays: extents as allocated by the reference
# (src/com_mod.f90:782-852 com_mod_allocate_part/_nests, src/outgrid_init.f90:192-201,
#  src/outgrid_init_nest.f90:40-60)
ALLOC_DIMS = {
    "xtra1": "maxpart", "ytra1": "maxpart", "ztra1": "maxpart", "itra1": "maxpart", "npoint": "maxpart",
    "nclass": "maxpart", "idt": "maxpart", "itramem": "maxpart", "itrasplit": "maxpart",
    "uap": "maxpart", "ucp": "maxpart", "uzp": "maxpart", "us": "maxpart", "vs": "maxpart", "ws": "maxpart",
    "cbt": "maxpart", "xmass1": "maxpart,maxspec", "xscav_frac1": "maxpart,maxspec",
    "gridunc": "0:numxgrid-1,0:numygrid-1,numzgrid,maxspec,maxpointspec_act,nclassunc,maxageclass",
    "griduncn": "0:numxgridn-1,0:numygridn-1,numzgrid,maxspec,maxpointspec_act,nclassunc,maxageclass",
    "drygridunc": "0:numxgrid-1,0:numygrid-1,maxspec,maxpointspec_act,nclassunc,maxageclass",
    "drygriduncn": "0:numxgridn-1,0:numygridn-1,maxspec,maxpointspec_act,nclassunc,maxageclass",
    "wetgridunc": "0:numxgrid-1,0:numygrid-1,maxspec,maxpointspec_act,nclassunc,maxageclass",
    "wetgriduncn": "0:numxgridn-1,0:numygridn-1,maxspec,maxpointspec_act,nclassunc,maxageclass",
    "outheight": "numzgrid", "outheighthalf": "numzgrid",
    "flux": "6,0:numxgrid-1,0:numygrid-1,numzgrid,nspec,maxpointspec_act,nageclass",
    "init_cond": "0:numxgrid-1,0:numygrid-1,numzgrid,maxspec,maxpointspec_act",
    "npart_av": "maxpart", "part_av_cartx": "maxpart", "part_av_carty": "maxpart", "part_av_cartz": "maxpart",
    "part_av_z": "maxpart", "part_av_topo": "maxpart", "part_av_pv": "maxpart", "part_av_qv": "maxpart",
    "part_av_tt": "maxpart", "part_av_rho": "maxpart", "part_av_tro": "maxpart", "part_av_hmix": "maxpart",
    "part_av_uu": "maxpart", "part_av_vv": "maxpart", "part_av_energy": "maxpart",
    "xmass": "numpoint,maxspec", "npart": "numpoint",
    "uun": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests", "vvn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests",
    "wwn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests", "ttn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests",
    "rhon": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests", "drhodzn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests",
    "ustarn": "0:nxmaxn-1,0:nymaxn-1,1,numwfmem,maxnests", "wstarn": "0:nxmaxn-1,0:nymaxn-1,1,numwfmem,maxnests",
    "hmixn": "0:nxmaxn-1,0:nymaxn-1,1,numwfmem,maxnests", "olin": "0:nxmaxn-1,0:nymaxn-1,1,numwfmem,maxnests",
    "tropopausen": "0:nxmaxn-1,0:nymaxn-1,1,numwfmem,maxnests",
    "vdepn": "0:nxmaxn-1,0:nymaxn-1,maxspec,numwfmem,maxnests",
    "cloudsn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests", "cloudshn": "0:nxmaxn-1,0:nymaxn-1,numwfmem,maxnests",
    "ctwcn": "0:nxmaxn-1,0:nymaxn-1,numwfmem,maxnests",
    "tthn": "0:nxmaxn-1,0:nymaxn-1,nuvzmax,numwfmem,maxnests", "qvhn": "0:nxmaxn-1,0:nymaxn-1,nuvzmax,numwfmem,maxnests",
    "qvn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests", "pvn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests",
    "clwcn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests", "ciwcn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests",
    "clwn": "0:nxmaxn-1,0:nymaxn-1,nzmax,numwfmem,maxnests",
    "clwchn": "0:nxmaxn-1,0:nymaxn-1,nuvzmax,numwfmem,maxnests", "ciwchn": "0:nxmaxn-1,0:nymaxn-1,nuvzmax,numwfmem,maxnests",
    "zpoint1": "numpoint", "zpoint2": "numpoint", "xpoint1": "numpoint", "xpoint2": "numpoint",
    "ypoint1": "numpoint", "ypoint2": "numpoint", "ireleasestart": "numpoint", "ireleaseend": "numpoint",
    "kindz": "numpoint", "rho_rel": "numpoint", "xmasssave": "numpoint",
    "area": "0:numxgrid-1,0:numygrid-1", "volume": "0:numxgrid-1,0:numygrid-1,numzgrid",
    "areaeast": "0:numxgrid-1,0:numygrid-1,numzgrid", "areanorth": "0:numxgrid-1,0:numygrid-1,numzgrid",
    "factor3d": "0:numxgrid-1,0:numygrid-1,numzgrid", "grid": "0:numxgrid-1,0:numygrid-1,numzgrid",
    "gridsigma": "0:numxgrid-1,0:numygrid-1,numzgrid", "wetgrid": "0:numxgrid-1,0:numygrid-1",
    "drygrid": "0:numxgrid-1,0:numygrid-1", "wetgridsigma": "0:numxgrid-1,0:numygrid-1",
    "drygridsigma": "0:numxgrid-1,0:numygrid-1",
    "sparse_dump_r": "numxgrid*numygrid*numzgrid", "sparse_dump_i": "numxgrid*numygrid*numzgrid",
    "densityoutgrid": "0:numxgrid-1,0:numygrid-1,numzgrid", "densitydrygrid": "0:numxgrid-1,0:numygrid-1,numzgrid",
}


class F2CError(Exception):
    pass


# calls inside extracted line ranges that belong to other subsystems and are guarded by
# switches the tests leave off (ipout=3, iflux=1, linit_cond>=1, DRYBKDEP): reaching one aborts
UNREACHED_CALLS = {"get_vdep_prob", "partpos_average", "calcfluxes", "initial_cond_calc"}


# ----------------------------------------------------------------- lexer ----
TOK = re.compile(r"""
    (?P<ws>\s+)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<dot>\.(?:and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv|true|false)\.)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_[a-z0-9_]+)?)
  | (?P<id>[a-z_][a-z0-9_]*)
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|\(/|/\)|::|[-+*/(),=<>:%;])
""", re.X)


def lex(line):
    toks, i, n = [], 0, len(line)
    while i < n:
        # a number directly followed by a dotted operator: "1.eq.2"
        m = re.match(r"(\d+)(?=\.(?:and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv)\.)", line[i:])
        if m:
            toks.append(("num", m.group(1)))
            i += m.end()
            continue
        m = TOK.match(line, i)
        if not m:
            raise F2CError(f"cannot lex: {line[i:i+30]!r} in {line!r}")
        i = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        toks.append((k, m.group(k)))
    return toks


def logical_lines(text):
    """[(label, lowercase statement text)] with comments stripped and continuations joined."""
    out, cur = [], ""
    for raw in text.splitlines():
        # strip comment (not inside a string)
        s, q, j = "", None, 0
        while j < len(raw):
            ch = raw[j]
            if q:
                s += ch
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                s += ch
            elif ch == "!":
                break
            else:
                s += ch
            j += 1
        s = s.rstrip()
        if not s.strip():
            continue
        if s.lstrip().startswith("#"):
            continue
        st = s.strip()
        if cur:
            if st.startswith("&"):
                st = st[1:]
            cur += " " + st
        else:
            cur = st
        if cur.endswith("&"):
            cur = cur[:-1]
            continue
        out.append(cur)
        cur = ""
    res = []
    for s in out:
        # lower-case outside strings
        parts = re.split(r"('(?:[^']|'')*'|\"(?:[^\"]|\"\")*\")", s)
        s = "".join(p if k % 2 else p.lower() for k, p in enumerate(parts))
        for piece in split_semicolons(s):
            m = re.match(r"^(\d+)\s+(.*)$", piece)
            if m:
                res.append((int(m.group(1)), m.group(2).strip()))
            else:
                res.append((None, piece.strip()))
    return res


def split_semicolons(s):
    out, depth, q, cur = [], 0, None, ""
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        if ch == ";":
            out.append(cur)
            cur = ""
            continue
        cur += ch
    out.append(cur)
    return [x for x in out if x.strip()]


# ---------------------------------------------------------------- symbols ----
class Sym:
    def __init__(self, name, ftype, dims=None, param=None, init=None, save=False, arg=False, module=None):
        self.name, self.ftype, self.dims, self.param, self.init = name, ftype, dims, param, init
        self.save, self.arg, self.module = save, arg, module
        self.alloc = False

    @property
    def ctype(self):
        return CTYPE[self.ftype]


INTRINSICS = {"abs", "sqrt", "exp", "log", "log10", "sin", "cos", "tan", "atan", "atan2", "asin", "acos", "max", "min",
              "amax1", "amin1", "mod", "modulo", "int", "nint", "real", "dble", "float", "sign", "floor", "tiny",
              "erf", "sngl", "ifix", "aint", "anint", "iabs", "dabs", "dsqrt", "dexp", "dlog", "dsin", "dcos", "datan2",
              "datan", "ceiling", "huge", "epsilon", "kind", "selected_real_kind", "selected_int_kind", "maxval", "minval",
              "isnan", "alog", "alog10"}


class Unit:
    """one subroutine / function / module body"""

    def __init__(self, kind, name, args, lines, module=None, result=None, rtype=None):
        self.kind, self.name, self.args, self.lines, self.module = kind, name, args, lines, module
        self.result, self.rtype = result or name, rtype
        self.syms = {}
        self.uses = []
        self.body = []


class Program:
    def __init__(self):
        self.modules = {}   # name -> Unit (declarations only)
        self.units = {}     # name -> Unit
        self.globals = {}   # name -> Sym (all module variables, flat)
        self.runtime_params = set()

    # ------------------------------------------------------------ parsing --
    def add_source(self, text, fname="?"):
        L = logical_lines(text)
        i, n = 0, len(L)
        cur_mod = None
        while i < n:
            lab, s = L[i]
            m = re.match(r"^module\s+([a-z_0-9]+)$", s)
            if m and not s.startswith("module procedure"):
                cur_mod = Unit("module", m.group(1), [], [])
                self.modules[cur_mod.name] = cur_mod
                i += 1
                # declarations until contains / end module
                while i < n and not re.match(r"^(contains|end\s*module)", L[i][1]):
                    cur_mod.lines.append(L[i])
                    i += 1
                if re.match(r"^end\s*module", L[i][1]):
                    cur_mod = None
                i += 1
                continue
            if re.match(r"^end\s*module", s):
                cur_mod = None
                i += 1
                continue
            m = re.match(r"^(?:(real|integer|logical|double\s*precision)(?:\s*\([^)]*\))?\s+)?(subroutine|function)\s+([a-z_0-9]+)\s*(?:\(([^)]*)\))?(?:\s*result\s*\(\s*([a-z_0-9]+)\s*\))?$", s)
            if m:
                rtype, kind, name, args, result = m.groups()
                args = [a.strip() for a in (args or "").split(",") if a.strip()]
                j = i + 1
                body = []
                while j < n and not re.match(rf"^end\s*({kind})?(\s+{name})?$", L[j][1]):
                    body.append(L[j])
                    j += 1
                if j >= n:
                    raise F2CError(f"{fname}: no end for {kind} {name}")
                u = Unit(kind, name, args, body, module=cur_mod.name if cur_mod else None, result=result, rtype=rtype)
                self.units[name] = u
                i = j + 1
                continue
            if s.startswith("contains") or s.startswith("implicit") or s.startswith("private") or s.startswith("public"):
                i += 1
                continue
            raise F2CError(f"{fname}: unexpected top-level statement: {s}")

    def add_extract(self, text, parent, name, first, last, args):
        """Synthetic subroutine `name(args)` = the declarations of subroutine `parent`
        + its executable source lines first..last (1-based, inclusive)."""
        if parent not in self.units:
            raise F2CError(f"extract: {parent} not parsed")
        decl = []
        for lab, s in self.units[parent].lines:
            if lab is None and (s.startswith("use ") or re.match(r"^(real|integer|logical|double\s*precision|character)\b", s)
                                or s.startswith("implicit")):
                decl.append((lab, s))
        body = logical_lines("\n".join(text.splitlines()[first - 1:last]))
        u = Unit("subroutine", name, list(args), decl + body)
        self.units[name] = u

    # --------------------------------------------------------- declarations --
    DECL = re.compile(r"^(real|integer|logical|double\s*precision|character)\s*(\([^)]*\)|\*\s*\d+)?\s*((?:,\s*[a-z]+(?:\s*\([^()]*(?:\([^()]*\)[^()]*)*\))?\s*)*)(::)?\s*(.*)$")

    def parse_decl(self, s, unit, module=None):
        m0 = re.match(r"^(real|integer|logical|double\s*precision|character)\b\s*", s)
        if not m0:
            return False
        base, rest0 = m0.group(1), s[m0.end():]
        kind = None
        if rest0.startswith("("):
            inner, rest0 = take_paren(rest0)
            kind = "(" + inner + ")"
        elif rest0.startswith("*"):
            mk = re.match(r"^\*\s*\d+\s*", rest0)
            kind, rest0 = mk.group(0), rest0[mk.end():]
        rest0 = rest0.strip()
        if "::" in rest0:
            attrs, rest = rest0.split("::", 1)
        else:
            if rest0.startswith(","):
                return False
            attrs, rest = "", rest0
        if base != "character" and not rest.strip():
            return False
        if base == "character":
            # strings are only used for messages and file names: remember the names so that
            # assignments to them can be dropped
            if unit is not None:
                names = getattr(unit, "char_vars", set())
                for ent in split_top(rest, ","):
                    mm = re.match(r"^\s*([a-z_][a-z0-9_]*)", ent)
                    if mm:
                        names.add(mm.group(1))
                unit.char_vars = names
            return True
        ftype = "double" if base.startswith("double") else base
        if kind:
            k = kind.replace(" ", "")
            if ftype == "real" and ("dp" in k or "8" in k):
                ftype = "double"
            elif ftype == "real" and ("sp" in k or "4" in k or "dep_prec" in k):
                ftype = "real"
            elif ftype == "integer" and re.search(r"(kind=)?1\)", k):
                ftype = "int8"
            elif ftype == "integer" and (re.search(r"(kind=)?2\)", k) or re.fullmatch(r"\*2", k)):
                ftype = "int16"
            elif ftype == "integer" and ("selected_int_kind(16)" in k or "8)" in k):
                ftype = "int64"
        attrs = attrs or ""
        is_param = "parameter" in attrs
        is_save = "save" in attrs
        is_alloc = "allocatable" in attrs
        dim_attr = None
        dm = re.search(r"dimension\s*\((.*)\)", attrs)
        if dm:
            dim_attr = dm.group(1)
        for ent in split_top(rest, ","):
            ent = ent.strip()
            if not ent:
                continue
            em = re.match(r"^([a-z_][a-z0-9_]*)\s*(\((.*?)\))?\s*(?:=\s*(.*))?$", ent)
            if not em:
                raise F2CError(f"cannot parse entity {ent!r} in {s!r}")
            name, _, dims, init = em.groups()
            # the greedy/lazy split above can cut "(...)" wrongly when init has parens: redo carefully
            name, dims, init = split_entity(ent)
            dims = dims if dims is not None else dim_attr
            sym = Sym(name, ftype, module=module)
            if dims is not None:
                if ":" in dims and re.fullmatch(r"[:,\s]*", dims):
                    sym.alloc = True
                    if name not in ALLOC_DIMS:
                        sym.dims = None  # unknown allocatable: only an error if referenced
                        sym.alloc_unknown = True
                    else:
                        sym.dims = parse_dims(ALLOC_DIMS[name])
                else:
                    sym.dims = parse_dims(dims)
            if is_alloc and sym.dims is None and name in ALLOC_DIMS:
                sym.dims = parse_dims(ALLOC_DIMS[name])
                sym.alloc = True
            if is_param:
                sym.param = init
            elif init is not None:
                sym.init = init
                sym.save = True
            sym.save = sym.save or is_save
            if name in unit.args:
                sym.arg = True
            unit.syms[name] = sym
        return True

    def collect_module_globals(self):
        for mod in self.modules.values():
            for lab, s in mod.lines:
                if s.startswith(("use ", "implicit", "save", "private", "public", "namelist", "type", "end type", "interface", "end interface")):
                    continue
                try:
                    ok = self.parse_decl(s, mod, module=mod.name)
                except F2CError:
                    ok = True  # declarations this subset cannot express are only an error if referenced
                if not ok and not re.match(r"^(external|intrinsic|data|common|equivalence)", s):
                    pass
            for name, sym in mod.syms.items():
                self.globals.setdefault(name, sym)


def split_top(s, sep):
    out, depth, cur, q = [], 0, "", None
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
            out.append(cur)
            cur = ""
        else:
            cur += ch
        i += 1
    out.append(cur)
    return out


def split_entity(ent):
    """name[(dims)][=init] -> (name, dims, init)"""
    m = re.match(r"^([a-z_][a-z0-9_]*)\s*", ent)
    name = m.group(1)
    rest = ent[m.end():]
    dims = init = None
    if rest.startswith("("):
        depth = 0
        for k, ch in enumerate(rest):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    dims = rest[1:k]
                    rest = rest[k + 1:].strip()
                    break
    rest = rest.strip()
    if rest.startswith("="):
        init = rest[1:].strip()
    elif rest:
        raise F2CError(f"cannot parse entity tail {rest!r} in {ent!r}")
    return name, dims, init


def parse_dims(d):
    dims = []
    for part in split_top(d, ","):
        part = part.strip()
        if part == "*":
            dims.append(("1", None))
        elif ":" in part:
            lo, hi = split_top(part, ":")
            dims.append((lo.strip() or "1", hi.strip() or None))
        else:
            dims.append(("1", part))
    return dims


# ------------------------------------------------------------- expressions ----
class ExprParser:
    def __init__(self, toks, ctx):
        self.t, self.i, self.ctx = toks, 0, ctx

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, v):
        k, s = self.next()
        if s != v:
            raise F2CError(f"expected {v!r}, got {s!r} in {self.t}")

    # precedence climbing: .eqv. < .or. < .and. < .not. < relational < +- < */ < unary < **
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        a = self.p_or()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            b = self.p_or()
            a = f"(({a}) != 0) {'==' if op == '.eqv.' else '!='} (({b}) != 0)"
            a = f"({a})"
        return a

    def p_or(self):
        a = self.p_and()
        while self.peek()[1] == ".or.":
            self.next()
            b = self.p_and()
            a = f"({a} || {b})"
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek()[1] == ".and.":
            self.next()
            b = self.p_not()
            a = f"({a} && {b})"
        return a

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.next()
            return f"(!{self.p_not()})"
        return self.p_rel()

    REL = {".eq.": "==", "==": "==", ".ne.": "!=", "/=": "!=", ".lt.": "<", "<": "<", ".le.": "<=", "<=": "<=",
           ".gt.": ">", ">": ">", ".ge.": ">=", ">=": ">="}

    def p_rel(self):
        a = self.p_add()
        if self.peek()[1] in self.REL:
            op = self.REL[self.next()[1]]
            b = self.p_add()
            return f"({a} {op} {b})"
        return a

    def p_add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = self.p_mul()
            a = f"({op}{a})"
        else:
            a = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            b = self.p_mul()
            a = f"({a} {op} {b})"
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            b = self.p_pow()
            a = f"({a} {op} {b})"
        return a

    def p_pow(self):
        a = self.p_prim()
        if self.peek()[1] == "**":
            self.next()
            # right associative; the exponent may carry a unary sign
            if self.peek()[1] in ("+", "-"):
                sg = self.next()[1]
                b = f"({sg}{self.p_pow()})"
            else:
                b = self.p_pow()
            return f"F_POW({a}, {b})"
        return a

    def p_prim(self):
        k, s = self.next()
        if k == "num":
            return number(s)
        if k == "dot":
            if s == ".true.":
                return "1"
            if s == ".false.":
                return "0"
            raise F2CError(f"unexpected {s}")
        if k == "str":
            return '"' + s[1:-1].replace('"', '\\"') + '"'
        if s == "(":
            e = self.expr()
            self.expect(")")
            return f"({e})"
        if s in ("+", "-"):
            return f"({s}{self.p_prim()})"
        if k == "id":
            args = None
            if self.peek()[1] == "(":
                self.next()
                args = []
                if self.peek()[1] != ")":
                    while True:
                        # keyword argument (kind=dp)
                        if self.peek()[0] == "id" and self.i + 1 < len(self.t) and self.t[self.i + 1][1] == "=":
                            kw = self.next()[1]
                            self.next()
                            args.append(("kw", kw, self.expr()))
                        else:
                            args.append(self.expr())
                        if self.peek()[1] == ",":
                            self.next()
                            continue
                        break
                self.expect(")")
            return self.ctx.ref(s, args)
        raise F2CError(f"unexpected token {s!r} in {self.t}")


def number(s):
    m = re.match(r"^((?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?)(?:_([a-z0-9_]+))?$", s)
    body, kind = m.groups()
    is_real = ("." in body) or ("e" in body) or ("d" in body)
    if not is_real:
        return body.lstrip("0") or "0" if not kind else body
    dbl = ("d" in body) or (kind in ("dp", "8"))
    body = body.replace("d", "e")
    if body.endswith("."):
        body += "0"
    if body.startswith("."):
        body = "0" + body
    body = re.sub(r"\.e", ".0e", body)
    return body if dbl else body + "f"


# ---------------------------------------------------------------- codegen ----
class Ctx:
    def __init__(self, prog, unit):
        self.prog, self.unit = prog, unit

    def lookup(self, name):
        if name in self.unit.syms:
            return self.unit.syms[name]
        g = self.prog.globals.get(name)
        if g is not None:
            self.prog.used_globals.add(name)
        return g

    def cexpr(self, text):
        p = ExprParser(lex(text), self)
        e = p.expr()
        if p.peek()[0] != "eof":
            raise F2CError(f"trailing tokens in expression {text!r}: {p.t[p.i:]}")
        return e

    def index(self, sym, args):
        if len(args) != len(sym.dims):
            raise F2CError(f"{sym.name}: rank {len(sym.dims)} but {len(args)} subscripts")
        # column-major offset
        off = None
        stride = None
        for (lo, hi), a in zip(sym.dims, args):
            lo_c = self.cexpr(lo)
            term = f"(({a}) - ({lo_c}))"
            if off is None:
                off = term
            else:
                off = f"{off} + {stride} * {term}"
            if hi is None:
                ext = None
            else:
                ext = f"(({self.cexpr(hi)}) - ({lo_c}) + 1)"
            if stride is None:
                stride = f"(long){ext}" if ext else None
            else:
                stride = f"{stride} * {ext}" if ext else None
        return off

    def var_c(self, sym):
        n = "f_" + sym.name
        if sym.arg and sym.dims is None:
            return f"(*{n})"
        return n

    def ref(self, name, args):
        sym = self.lookup(name)
        if sym is not None and getattr(sym, "alloc_unknown", False):
            raise F2CError(f"allocatable {name} has no entry in ALLOC_DIMS")
        if sym is not None and sym.dims is not None:
            if args is None:
                return "f_" + name  # whole array (as actual argument)
            return f"f_{name}[{self.index(sym, args)}]"
        if sym is not None and args is None:
            if sym.param is not None and sym.module is None:
                return "f_" + name
            return self.var_c(sym)
        if args is None:
            if name == self.unit.result and self.unit.kind == "function":
                return "f_" + name + "_result"
            raise F2CError(f"{self.unit.name}: undeclared variable {name}")
        if sym is not None and sym.dims is None and name != self.unit.name and name not in self.prog.units \
                and name not in INTRINSICS:
            raise F2CError(f"{self.unit.name}: scalar {name} referenced with arguments")
        return self.call_fn(name, args)

    def call_fn(self, name, args):
        pos = [a for a in args if not isinstance(a, tuple)]
        kw = {a[1]: a[2] for a in args if isinstance(a, tuple)}
        if name in INTRINSICS and name not in self.prog.units:
            return self.intrinsic(name, pos, kw)
        if name in self.prog.units:
            u = self.prog.units[name]
            self.prog.called.add(name)
            return f"f_{name}({', '.join(self.actual(a) for a in pos)})"
        raise F2CError(f"{self.unit.name}: unknown function {name}")

    def actual(self, cexp):
        """actual argument by reference"""
        m = re.fullmatch(r"\(\*(f_[a-z0-9_]+)\)", cexp)
        if m:
            return m.group(1)  # dummy scalar passed on
        if re.fullmatch(r"f_[a-z0-9_]+", cexp):
            nm = cexp[2:]
            sym = self.lookup(nm)
            if sym is not None and sym.dims is not None:
                return cexp  # array -> pointer
            if sym is not None and sym.param is not None:
                return f"&(__typeof__({cexp})){{{cexp}}}"
            return "&" + cexp
        if re.fullmatch(r"f_[a-z0-9_]+\[.*\]", cexp) and balanced_index(cexp):
            return "&" + cexp
        return f"&(__typeof__({cexp})){{{cexp}}}"

    def intrinsic(self, name, a, kw):
        n = len(a)
        if name in ("max", "amax1"):
            e = a[0]
            for b in a[1:]:
                e = f"F_MAX({e}, {b})"
            return e
        if name in ("min", "amin1"):
            e = a[0]
            for b in a[1:]:
                e = f"F_MIN({e}, {b})"
            return e
        if name in ("abs", "iabs", "dabs"):
            return f"F_ABS({a[0]})"
        one = {"sqrt": "F_SQRT", "dsqrt": "F_SQRT", "exp": "F_EXP", "dexp": "F_EXP", "log": "F_LOG", "dlog": "F_LOG",
               "alog": "F_LOG", "alog10": "F_LOG10", "log10": "F_LOG10", "sin": "F_SIN", "dsin": "F_SIN", "cos": "F_COS", "dcos": "F_COS", "tan": "F_TAN",
               "atan": "F_ATAN", "datan": "F_ATAN", "asin": "F_ASIN", "acos": "F_ACOS", "erf": "F_ERF",
               "nint": "F_NINT", "floor": "F_FLOOR", "ceiling": "F_CEILING", "aint": "F_AINT", "anint": "F_ANINT"}
        if name in one:
            return f"{one[name]}({a[0]})"
        if name in ("atan2", "datan2"):
            return f"F_ATAN2({a[0]}, {a[1]})"
        if name == "mod":
            return f"F_MOD({a[0]}, {a[1]})"
        if name == "modulo":
            return f"F_MODULO({a[0]}, {a[1]})"
        if name == "sign":
            return f"F_SIGN({a[0]}, {a[1]})"
        if name in ("int", "ifix"):
            return f"((int)({a[0]}))"
        if name in ("float", "sngl"):
            return f"((float)({a[0]}))"
        if name == "dble":
            return f"((double)({a[0]}))"
        if name == "real":
            k = kw.get("kind") or (a[1] if n > 1 else None)
            if k is not None and ("dp" in k or k.strip("()") == "8"):
                return f"((double)({a[0]}))"
            return f"((float)({a[0]}))"
        if name == "isnan":
            return f"(isnan({a[0]}) != 0)"
        if name == "tiny":
            return f"F_TINY({a[0]})"
        if name == "huge":
            return f"F_HUGE({a[0]})"
        raise F2CError(f"intrinsic {name} not supported")


def balanced_index(c):
    k = c.index("[")
    depth = 0
    for j in range(k, len(c)):
        if c[j] == "[":
            depth += 1
        elif c[j] == "]":
            depth -= 1
            if depth == 0:
                return j == len(c) - 1
    return False


class Gen:
    def __init__(self, prog):
        self.prog = prog
        prog.used_globals = set()
        prog.called = set()

    def unit_c(self, u):
        ctx = Ctx(self.prog, u)
        P = self.prog
        out, decl_done = [], False
        body = []
        # pass 1: declarations
        stmts = []
        for lab, s in u.lines:
            if s.startswith("use ") or s.startswith("implicit") or s.startswith("external") or s.startswith("intrinsic"):
                continue
            m = re.match(r"^save\s+(.*)$", s)
            if m:
                for nm in m.group(1).split(","):
                    u.pending_save = getattr(u, "pending_save", []) + [nm.strip()]
                continue
            if s == "save":
                u.save_all = True
                continue
            m = re.match(r"^parameter\s*\((.*)\)$", s)
            if m and lab is None:
                for ent in split_top(m.group(1), ","):
                    nm, val = ent.split("=", 1)
                    nm = nm.strip()
                    if nm not in u.syms:
                        raise F2CError(f"{u.name}: parameter {nm} not declared")
                    u.syms[nm].param = val.strip()
                continue
            if lab is None and P.parse_decl(s, u) and "::" in s or (lab is None and re.match(r"^(real|integer|logical|double\s*precision|character)\b(?!\s*function)", s) and P.parse_decl(s, u)):
                continue
            stmts.append((lab, s))
        for nm in getattr(u, "pending_save", []):
            if nm in u.syms:
                u.syms[nm].save = True
        if u.kind == "function":
            rs = u.syms.get(u.result)
            if rs is None:
                if not u.rtype:
                    raise F2CError(f"function {u.name}: no result type")
                rt = "double" if u.rtype.startswith("double") else u.rtype
                rs = Sym(u.result, rt)
            u.rsym = rs
            u.syms.pop(u.result, None)
        # signature
        params = []
        for a in u.args:
            sym = u.syms.get(a)
            if sym is None:
                raise F2CError(f"{u.name}: dummy argument {a} not declared")
            sym.arg = True
            params.append(f"{sym.ctype} *f_{a}")
        rtype = u.rsym.ctype if u.kind == "function" else "void"
        sig = f"{rtype} f_{u.name}({', '.join(params) if params else 'void'})"
        u.sig = sig
        out.append(sig + " {")
        # locals
        for nm, sym in u.syms.items():
            if sym.arg:
                continue
            if sym.dims is None and sym.param is None and (nm in P.units and P.units[nm].kind == "function"
                                                            or nm in INTRINSICS):
                continue  # type declaration of an external / intrinsic function
            ct = sym.ctype
            if sym.param is not None:
                if sym.dims is not None:
                    vals = array_ctor(sym.param, ctx)
                    lo = ctx.cexpr(sym.dims[0][0])
                    out.append(f"  static const {ct} f_{nm}_v[] = {{{', '.join(vals)}}};")
                    out.append(f"  const {ct} *f_{nm} = f_{nm}_v;")
                else:
                    out.append(f"  const {ct} f_{nm} = {ctx.cexpr(sym.param)};")
                continue
            static = "static " if (sym.save or getattr(u, "save_all", False)) else ""
            if sym.dims is not None:
                size = " * ".join(f"(({ctx.cexpr(hi)}) - ({ctx.cexpr(lo)}) + 1)" for lo, hi in sym.dims)
                if static:
                    n_const = try_const(size, u, ctx)
                    if n_const is None:
                        # extent from a par_mod parameter (run-time value here): allocated at the first call
                        if sym.init is not None:
                            raise F2CError(f"{u.name}: initialised SAVEd array {nm} with non-constant size")
                        out.append(f"  static {ct} *f_{nm} = NULL; if (!f_{nm}) f_{nm} = ({ct} *)calloc((size_t)({size}), sizeof({ct}));")
                        continue
                    init = ""
                    if sym.init is not None:
                        vals = array_ctor(sym.init, ctx, n_const)
                        init = " = {" + ", ".join(vals) + "}"
                    out.append(f"  static {ct} f_{nm}[{n_const}]{init};")
                else:
                    n_const = try_const(size, u, ctx)
                    if n_const is not None and n_const <= 16384:
                        out.append(f"  {ct} f_{nm}[{size}]; memset(f_{nm}, 0, sizeof f_{nm});")
                    else:
                        # automatic array of run-time (or large) extent: kept on the heap between calls,
                        # zeroed at every entry (-finit-local-zero, src/makefile_meteoswiss:89)
                        out.append(f"  static {ct} *f_{nm} = NULL; static size_t f_{nm}_n = 0; "
                                   f"if (f_{nm}_n != (size_t)({size})) {{ free(f_{nm}); f_{nm}_n = (size_t)({size}); "
                                   f"f_{nm} = ({ct} *)malloc(f_{nm}_n * sizeof({ct})); }} "
                                   f"memset(f_{nm}, 0, f_{nm}_n * sizeof({ct}));")
            else:
                init = f" = {ctx.cexpr(sym.init)}" if sym.init is not None else " = 0"
                out.append(f"  {static}{ct} f_{nm}{init};")
        if u.kind == "function":
            out.append(f"  {u.rsym.ctype} f_{u.result}_result = 0;")
            # inside the function its own name is the result variable
        # pass 2: executable statements
        self.indent = 1
        self.do_stack = []
        self.where_stack = []
        for lab, s in stmts:
            for line in self.stmt(ctx, u, lab, s):
                out.append("  " * self.indent_for(line) + line)
        if u.kind == "function":
            out.append(f"  return f_{u.result}_result;")
        out.append("}")
        return "\n".join(out)

    def indent_for(self, line):
        return 1

    def stmt(self, ctx, u, lab, s):
        res = []
        if lab is not None:
            res.append(f"L{lab}: ;")
        if s == "continue":
            return res
        # one-line if
        m = re.match(r"^if\s*\(", s)
        if m:
            cond, rest = take_paren(s[s.index("("):])
            rest = rest.strip()
            c = ctx.cexpr(cond)
            if rest == "then":
                return res + [f"if ({c}) {{"]
            if not rest:
                raise F2CError(f"arithmetic if? {s}")
            inner = self.stmt(ctx, u, None, rest)
            return res + [f"if ({c}) {{"] + inner + ["}"]
        m = re.match(r"^else\s*if\s*\(", s)
        if m:
            cond, rest = take_paren(s[s.index("("):])
            if rest.strip() != "then":
                raise F2CError(f"bad else if: {s}")
            return res + [f"}} else if ({ctx.cexpr(cond)}) {{"]
        if s == "else":
            return res + ["} else {"]
        if re.match(r"^end\s*if$", s):
            return res + ["}"]
        # named do construct: `name: do ...`, left through `exit name` (a goto past its end)
        m = re.match(r"^([a-z_][a-z0-9_]*)\s*:\s*(do\b.*)$", s)
        do_name = None
        if m:
            do_name, s = m.group(1), m.group(2)
        if re.match(r"^do\b", s):
            self.do_count = getattr(self, "do_count", 0) + 1
            self.do_stack.append((do_name, self.do_count))
        m = re.match(r"^do\s+([a-z_][a-z0-9_]*)\s*=\s*(.*)$", s)
        if m:
            var, rng = m.groups()
            parts = split_top(rng, ",")
            v = ctx.ref(var, None)
            a, b = ctx.cexpr(parts[0]), ctx.cexpr(parts[1])
            if len(parts) == 3:
                st = ctx.cexpr(parts[2])
                return res + [f"{{ const int _b = {b}, _s = {st}; for ({v} = {a}; _s > 0 ? {v} <= _b : {v} >= _b; {v} += _s) {{"]
            return res + [f"{{ const int _b = {b}; for ({v} = {a}; {v} <= _b; {v}++) {{"]
        if s == "do":
            return res + ["{ for (;;) {"]
        m = re.match(r"^end\s*do(?:\s+([a-z_][a-z0-9_]*))?$", s)
        if m:
            nm, cnt = self.do_stack.pop() if self.do_stack else (None, 0)
            if nm is not None:
                return res + ["} }", f"Lx_{nm}_{cnt}: ;"]
            return res + ["} }"]
        m = re.match(r"^exit\s+([a-z_][a-z0-9_]*)$", s)
        if m:
            for nm, cnt in reversed(self.do_stack):
                if nm == m.group(1):
                    return res + [f"goto Lx_{nm}_{cnt};"]
            raise F2CError(f"{u.name}: exit {m.group(1)}: no such construct")
        m = re.match(r"^go\s*to\s*(\d+)$", s)
        if m:
            return res + [f"goto L{m.group(1)};"]
        if s == "exit":
            return res + ["break;"]
        if s == "cycle":
            return res + ["continue;"]
        if s == "return":
            if u.kind == "function":
                return res + [f"return f_{u.result}_result;"]
            return res + ["return;"]
        if s.startswith("stop"):
            return res + ['f2c_stop();']
        if re.match(r"^(write|print|open|close|format|read|flush)\b", s):
            return res + ["/* i/o statement skipped */;"]
        m = re.match(r"^call\s+([a-z_][a-z0-9_]*)\s*(?:\((.*)\))?$", s)
        if m:
            name, args = m.groups()
            if name == "mean":  # generic interface of mean_mod; dep_prec = sp selects mean_sp (src/par_mod.f90:33)
                name = "mean_sp"
            if name not in self.prog.units:
                if name in ("flush", "mpif_mtime", "caldate"):
                    return res + [f"/* call {name} skipped */;"]
                if name in UNREACHED_CALLS:
                    return res + [f"f2c_stop(); /* call {name}: outside the path, must not be reached */"]
                raise F2CError(f"{u.name}: call to unknown subroutine {name}")
            self.prog.called.add(name)
            al = [ctx.actual(ctx.cexpr(a)) for a in split_top(args, ",")] if args and args.strip() else []
            return res + [f"f_{name}({', '.join(al)});"]
        # WHERE construct: the assignments of its body become masked element loops
        m = re.match(r"^where\s*\(", s)
        if m:
            mask, rest = take_paren(s[s.index("("):])
            rest = rest.strip()
            if rest:
                self.where_stack.append(mask)
                try:
                    inner = self.stmt(ctx, u, None, rest)
                finally:
                    self.where_stack.pop()
                return res + inner
            self.where_stack.append(mask)
            return res
        if re.match(r"^else\s*where$", s):
            if not self.where_stack:
                raise F2CError(f"{u.name}: elsewhere outside where")
            self.where_stack[-1] = ".not.(" + self.where_stack[-1] + ")"
            return res
        if re.match(r"^end\s*where$", s):
            self.where_stack.pop()
            return res
        # assignment
        lhs, rhs = split_assign(s)
        if lhs is None:
            raise F2CError(f"{u.name}: cannot translate statement: {s}")
        el = self.elementalize(ctx, u, lhs, rhs, self.where_stack[-1] if self.where_stack else None)
        if el is not None:
            return res + el
        lm = re.match(r"^([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", lhs.strip())
        if not lm:
            raise F2CError(f"{u.name}: bad assignment target {lhs}")
        name = lm.group(1)
        sym = ctx.lookup(name)
        if u.kind == "function" and name == u.result and lm.group(2) is None:
            return res + [f"f_{name}_result = {ctx.cexpr(rhs)};"]
        if sym is None and name in getattr(u, "char_vars", set()):
            return res + ["/* character assignment skipped */;"]
        if sym is None:
            raise F2CError(f"{u.name}: assignment to undeclared {name}")
        if sym.param is not None:
            raise F2CError(f"{u.name}: assignment to parameter {name}")
        if sym.dims is not None and lm.group(2) is None:
            # whole-array assignment of a scalar
            size = " * ".join(f"(({ctx.cexpr(hi)}) - ({ctx.cexpr(lo)}) + 1)" for lo, hi in sym.dims)
            return res + [f"{{ long _n = {size}; for (long _k = 0; _k < _n; _k++) f_{name}[_k] = {ctx.cexpr(rhs)}; }}"]
        if sym.dims is not None and ":" in (lm.group(3) or "") and len(sym.dims) == 1 \
                and not re.fullmatch(r"[:,\s]*", lm.group(3)):
            sec = lm.group(3).strip()
            lo, hi = [x.strip() for x in split_top(sec, ":")]
            pat = "(" + sec + ")"
            if rhs.replace(" ", "").count(":") != rhs.replace(" ", "").count(pat.replace(" ", "")):
                raise F2CError(f"{u.name}: array section assignment not supported: {s}")
            body = f"{name}(k_sec_) = " + rhs.replace(pat, "(k_sec_)")
            u.syms.setdefault("k_sec_", Sym("k_sec_", "integer"))
            inner = self.stmt(ctx, u, None, body)
            return res + [f"{{ int f_k_sec_; for (f_k_sec_ = {ctx.cexpr(lo)}; f_k_sec_ <= {ctx.cexpr(hi)}; f_k_sec_++) {{"] + inner + ["} }"]
        if sym.dims is not None and ":" in (lm.group(3) or ""):
            if re.fullmatch(r"[:,\s]*", lm.group(3)):
                size = " * ".join(f"(({ctx.cexpr(hi)}) - ({ctx.cexpr(lo)}) + 1)" for lo, hi in sym.dims)
                return res + [f"{{ long _n = {size}; for (long _k = 0; _k < _n; _k++) f_{name}[_k] = {ctx.cexpr(rhs)}; }}"]
            raise F2CError(f"{u.name}: array section assignment not supported: {s}")
        return res + [f"{ctx.cexpr(lhs)} = {ctx.cexpr(rhs)};"]


    # ---- array syntax -> element loops -----------------------------------------------------------
    def elementalize(self, ctx, u, lhs, rhs, mask):
        """`A(sections) = expr` (or whole-array `A = expr` with an array-valued expr, or any assignment
        under a WHERE mask) as nested element loops; None when the statement is scalar."""
        lm = re.match(r"^([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", lhs.strip())
        if not lm:
            return None
        name = lm.group(1)
        sym = ctx.lookup(name)
        if sym is None or sym.dims is None:
            if mask is not None:
                raise F2CError(f"{u.name}: scalar assignment inside where: {lhs} = {rhs}")
            return None
        secs = []   # (lo, hi) Fortran text of every sectioned dimension of the target, in order

        def rewrite_ref(sym_r, args_text, target=False):
            """subscript list of one array reference with its sections replaced by loop indices;
            returns (new text or None when the reference has no section, number of sections)"""
            if args_text is None:
                subs = [":"] * len(sym_r.dims)
            else:
                subs = [x.strip() for x in split_top(args_text, ",")]
            if len(subs) != len(sym_r.dims):
                raise F2CError(f"{u.name}: {sym_r.name}: rank {len(sym_r.dims)} but {len(subs)} subscripts")
            out, k = [], 0
            for (dlo, dhi), sub in zip(sym_r.dims, subs):
                parts = split_top(sub, ":")
                if len(parts) == 1:
                    out.append(rewrite_expr(sub))
                    continue
                if len(parts) > 2:
                    raise F2CError(f"{u.name}: strided section {sub}")
                lo = parts[0].strip() or dlo
                hi = parts[1].strip() or dhi
                out.append(f"(({lo})+isec{k}_)")
                if target:
                    secs.append((lo, hi))
                k += 1
            return ",".join(out), k

        def rewrite_expr(text):
            """array references with sections / whole arrays in an expression -> element references"""
            toks = lex(text)
            out, i = [], 0
            while i < len(toks):
                kind, val = toks[i]
                if kind == "id":
                    sy = ctx.lookup(val)
                    is_arr = sy is not None and sy.dims is not None
                    if i + 1 < len(toks) and toks[i + 1][1] == "(":
                        depth, j = 0, i + 1
                        while True:
                            if toks[j][1] == "(":
                                depth += 1
                            elif toks[j][1] == ")":
                                depth -= 1
                                if depth == 0:
                                    break
                            j += 1
                        inner = " ".join(t[1] for t in toks[i + 2:j])
                        if is_arr:
                            new, k = rewrite_ref(sy, inner)
                            if k and k != nsec[0]:
                                raise F2CError(f"{u.name}: non-conforming section of {val} in {lhs} = {rhs}")
                            out.append(f"{val}({new})")
                        else:
                            args = [rewrite_expr(a) for a in split_top(inner, ",")] if inner.strip() else []
                            out.append(f"{val}({','.join(args)})")
                        i = j + 1
                        continue
                    if is_arr and sy.param is None:
                        if len(sy.dims) != nsec[0]:
                            raise F2CError(f"{u.name}: whole array {val} (rank {len(sy.dims)}) in {lhs} = {rhs}")
                        new, _ = rewrite_ref(sy, None)
                        out.append(f"{val}({new})")
                        i += 1
                        continue
                out.append(val)
                i += 1
            return " ".join(out)

        nsec = [0]
        new_lhs_args, n = rewrite_ref(sym, lm.group(3), target=True)
        if n == 0:
            if mask is not None:
                raise F2CError(f"{u.name}: element assignment inside where: {lhs} = {rhs}")
            return None
        nsec[0] = n
        # a scalar right-hand side without any array keeps the fast whole-array path of the caller
        new_rhs = rewrite_expr(rhs)
        cond = rewrite_expr(mask) if mask is not None else None
        if cond is not None and re.search(rf"\b{name}\b", mask):
            raise F2CError(f"{u.name}: where mask depends on its target {name}")
        lines = []
        for k in range(n - 1, -1, -1):
            u.syms.setdefault(f"isec{k}_", Sym(f"isec{k}_", "integer"))
            lo, hi = secs[k]
            lines.append(f"{{ int f_isec{k}_; const int _n{k} = ({ctx.cexpr(hi)}) - ({ctx.cexpr(lo)}); "
                         f"for (f_isec{k}_ = 0; f_isec{k}_ <= _n{k}; f_isec{k}_++) {{")
        body = f"{ctx.cexpr(name + '(' + new_lhs_args + ')')} = {ctx.cexpr(new_rhs)};"
        if cond is not None:
            body = f"if ({ctx.cexpr(cond)}) {body}"
        lines.append(body)
        lines += ["} }"] * n
        return lines


def take_paren(s):
    """s starts with '(' -> (inside, rest)"""
    depth = 0
    q = None
    for k, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return s[1:k], s[k + 1:]
    raise F2CError(f"unbalanced parentheses: {s}")


def split_assign(s):
    depth, q = 0, None
    for k, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if s[k + 1:k + 2] == "=" or s[k - 1:k] in ("=", "/", "<", ">"):
                continue
            return s[:k], s[k + 1:]
    return None, None


def array_ctor(text, ctx, n=None):
    t = text.strip()
    if not (t.startswith("(/") and t.endswith("/)")):
        raise F2CError(f"array initialiser {text!r} not supported")
    inner = t[2:-2].strip()
    m = re.match(r"^\(\s*(.*?)\s*,\s*([a-z_][a-z0-9_]*)\s*=\s*(.*?)\s*,\s*(.*?)\s*\)$", inner)
    if m and m.group(2) not in m.group(1):
        # implied do with a constant element
        if n is None:
            raise F2CError("implied-do constructor needs a known size")
        return [ctx.cexpr(m.group(1))] * n
    return [ctx.cexpr(x) for x in split_top(inner, ",")]


def try_const(size_expr, u=None, ctx=None):
    """integer value of a C size expression made of literals and local integer parameters"""
    e = re.sub(r"\(long\)", "", size_expr)
    for _ in range(8):
        names = set(re.findall(r"f_([a-z0-9_]+)", e))
        if not names:
            break
        for nm in names:
            sym = u.syms.get(nm) if u else None
            if sym is None or sym.param is None:
                return None
            e = re.sub(rf"\bf_{nm}\b", "(" + ctx.cexpr(sym.param) + ")", e)
    try:
        v = eval(e.replace("/", "//"), {"__builtins__": {}}, {})
        return int(v)
    except Exception:
        return None


# ----------------------------------------------------------------- driver ----
def emit(prog, wanted, runtime_int_params):
    gen = Gen(prog)
    # closure of called units
    todo, done, bodies = list(wanted), [], {}
    while todo:
        name = todo.pop()
        if name in bodies:
            continue
        if name not in prog.units:
            raise F2CError(f"unit {name} not found")
        before = set(prog.called)
        bodies[name] = gen.unit_c(prog.units[name])
        done.append(name)
        for c in prog.called - set(bodies):
            todo.append(c)
    # globals actually referenced (closure over dimension / parameter expressions)
    ctx = Ctx(prog, Unit("module", "_globals", [], []))
    pending = set(prog.used_globals)
    ginfo = {}
    while pending:
        nm = pending.pop()
        if nm in ginfo:
            continue
        sym = prog.globals[nm]
        before = set(prog.used_globals)
        info = {"sym": sym}
        if sym.param is not None and not (sym.ftype == "integer" and sym.module == "par_mod" and nm in runtime_int_params):
            info["value"] = ctx.cexpr(sym.param) if not sym.param.strip().startswith("selected_") else "0"
        elif sym.param is not None:
            info["runtime"] = ctx.cexpr(sym.param)
        if sym.dims is not None:
            info["dims"] = [(ctx.cexpr(lo), ctx.cexpr(hi) if hi is not None else None) for lo, hi in sym.dims]
        elif sym.init is not None:
            info["init"] = ctx.cexpr(sym.init)
        ginfo[nm] = info
        pending |= prog.used_globals - set(ginfo)
    return done, bodies, ginfo


def main(argv):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", required=True, help="reference src directory")
    ap.add_argument("--files", nargs="+", required=True)
    ap.add_argument("--units", nargs="+", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--extract", nargs="*", default=[],
                    help="name:file:parent:first-last:arg,arg  (line range of a subroutine as its own unit)")
    args = ap.parse_args(argv)
    import os
    prog = Program()
    for f in args.files:
        # (a path with a directory part is taken as it is: the stub modules under oracle/f2c/stubs)
        path = f if os.sep in f else os.path.join(args.src, f)
        prog.add_source(open(path, errors="replace").read(), os.path.basename(f))
    for spec in args.extract:
        name, f, parent, rng, al = spec.split(":")
        first, last = [int(x) for x in rng.split("-")]
        prog.add_extract(open(os.path.join(args.src, f), errors="replace").read(), parent, name, first, last,
                         [a for a in al.split(",") if a])
        args.units.append(name)
    prog.collect_module_globals()
    runtime = {"nxmax", "nymax", "nuvzmax", "nwzmax", "nzmax", "maxnests", "nxmaxn", "nymaxn", "maxpart", "maxspec",
               "maxageclass", "nclassunc", "maxreceptor", "maxrand", "numwfmem", "nconvlevmax", "na", "maxpoint"}
    done, bodies, ginfo = emit(prog, args.units, runtime)
    with open(args.out, "w") as o:
        o.write("/* GENERATED by oracle/f2c/f90toc.py from the reference's Fortran sources -- do not commit */\n")
        o.write('#include <stdlib.h>\n#include <string.h>\n#include <stdio.h>\n#include "f2c_rt.h"\n\n')
        # globals: scalars first (so that dimension expressions can see them)
        order = sorted(ginfo, key=lambda n: (ginfo[n]["sym"].dims is not None, n))
        for nm in order:
            g = ginfo[nm]
            sym = g["sym"]
            ct = sym.ctype
            if "value" in g:
                if sym.ftype in ("integer", "logical"):
                    o.write(f"enum {{ f_{nm} = {g['value']} }};\n") if is_int_const(g["value"]) else o.write(f"static const {ct} f_{nm} = {g['value']};\n")
                else:
                    o.write(f"#define f_{nm} (({ct})({g['value']}))\n")
            elif "runtime" in g:
                o.write(f"{ct} f_{nm}; /* par_mod parameter, run-time here; reference value {g['runtime']} */\n")
            elif sym.dims is None:
                o.write(f"{ct} f_{nm}{' = ' + g['init'] if 'init' in g else ''};\n")
            else:
                o.write(f"{ct} *f_{nm};\n")
        o.write("\n/* prototypes */\n")
        for nm in done:
            o.write(prog.units[nm].sig + ";\n")
        o.write("\n")
        for nm in reversed(done):
            o.write(bodies[nm] + "\n\n")
        # registry
        o.write("/* ---- registry for the test harness ---- */\n")
        o.write("void ref_defaults(void) {\n")
        for nm in order:
            g = ginfo[nm]
            if "runtime" in g:
                o.write(f"  f_{nm} = {g['runtime']};\n")
        o.write("}\n")
        o.write("void ref_alloc(void) {\n")
        for nm in order:
            g = ginfo[nm]
            if "dims" in g:
                size = " * ".join(f"((long)({hi}) - ({lo}) + 1)" for lo, hi in g["dims"])
                o.write(f"  free(f_{nm}); {{ long n_ = {size}; if (n_ < 1) n_ = 1; f_{nm} = calloc((size_t)n_, sizeof *f_{nm}); }}\n")
        o.write("}\n")
        o.write("void *ref_ptr(const char *name) {\n")
        for nm in order:
            g = ginfo[nm]
            if "value" in g:
                continue
            if "dims" in g:
                o.write(f'  if (!strcmp(name, "{nm}")) return f_{nm};\n')
            else:
                o.write(f'  if (!strcmp(name, "{nm}")) return &f_{nm};\n')
        o.write("  return 0;\n}\n")
        o.write("long ref_extent(const char *name, int d) {\n")
        for nm in order:
            g = ginfo[nm]
            if "dims" in g:
                for k, (lo, hi) in enumerate(g["dims"]):
                    o.write(f'  if (!strcmp(name, "{nm}") && d == {k}) return (long)({hi}) - ({lo}) + 1;\n')
        o.write("  return -1;\n}\n")
        o.write("const char *ref_type(const char *name) {\n")
        for nm in order:
            g = ginfo[nm]
            if "value" in g:
                continue
            o.write(f'  if (!strcmp(name, "{nm}")) return "{g["sym"].ctype}";\n')
        o.write("  return 0;\n}\n")
    print(f"f90toc: {len(done)} units, {len(ginfo)} module variables -> {args.out}")


def is_int_const(v):
    return re.fullmatch(r"[-+()\d\s*/]+", v) is not None


if __name__ == "__main__":
    main(sys.argv[1:])

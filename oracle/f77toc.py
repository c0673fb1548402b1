#!/usr/bin/env python3
"""f77toc.py -- transliterate the reference's own Fortran-77 kernels to C so they can be compiled here.

TEST INFRASTRUCTURE (oracle pin).  The image has no Fortran compiler, so the reference's numerical
kernels cannot be built as shipped.  This script reads the Fortran sources WHERE THEY LIE under the
reference tree (never copied into the repo), translates the handful of subroutines on the Vlasov RHS
path statement by statement into C (same loops, same expression trees, every binary operation
parenthesised so the evaluation order is the Fortran parse order, `x**n` expanded the way gfortran
expands integer powers), and writes the result ONLY into oracle/_ref/ (git-ignored).  The Makefile then
compiles it with `gcc -O2 -ffp-contract=off` -- the analogue of the reference's `gfortran -O2
-fdefault-real-8` build (configure.in:219-232, 273-280) -- into oracle/_ref/libloki_ref.so, which
tests/test_oracle_pin.py uses to pin the hand-written restatement (oracle/loki_oracle.c) bit for bit.

Supported subset (all that the listed routines use): fixed-form source, comment/continuation lines,
implicit none, integer / integer*8 / real / double precision declarations with explicit-bound or
assumed-size arrays, do / end do, block and one-line if, call, assignment, return, the intrinsics
max/min/abs, external real functions, relational/logical dot-operators, integer-constant powers.

usage: f77toc.py <reference_dir> <out_dir>
"""
import os
import re
import sys

ROUTINES = {
    "KineticSpeciesF.f": ["xpby4d", "setphasespacevel4d", "setphasespacevelmaxwell4d", "weno43fit4d",
                          "weno65fit4d", "setaccelerationbcs4d", "setadvectionbcs4d", "setaccelerationbcs4djb", "setadvectionbcs4djb", "computeadvectionderivatives4d",
                          "computeaccelerationderivatives4d", "computecurrents", "computekeedot", "computeke",
                          "computekemaxwell", "appendkrook", "weno43avg4d", "weno65avg4d", "computeflux4d",
                          "computeadvectionfluxes4d", "computeaccelerationfluxes4d", "accumfluxdiv4d", "computekeflux",
                          "computekevelspaceflux"],
    "PoissonF.f": ["neutralizecharge4d", "computeefieldfrompotential"],
    "MaxwellF.f": ["maxwellevalrhs", "sgmetricfunction", "maxwellevalvzrhs", "xpby2d", "zeroghost2d",
                   "maxwelladdantennasource", "maxwellsetembcs", "maxwellsetvzbcs"],
    "PitchAngleCollisionOperatorF.f": ["evaluatecollisionality", "conservativepitchangle_4th",
                                       "conservativepitchangle_6th", "nonconservativepitchangle_4th",
                                       "appendpitchanglecollision", "computepitchanglespeciesmoments",
                                       "computepitchanglespeciesreducedfields", "computepitchanglespecieskec",
                                       "computepitchanglespeciesvthermal"],
    "TZSourceF.f": ["settrigtzsource", "computetrigtzsourceerror"],
    "ElectronTZSourceF.f": ["setelectrontrigtzsource", "computeelectrontrigtzsourceerror"],
    "TwoSpecies_ElectronTZSourceF.f": ["settwoelectrontrigtzsource", "computetwoelectrontrigtzsourceerror"],
    "TwoSpecies_IonTZSourceF.f": ["settwoiontrigtzsource", "computetwoiontrigtzsourceerror"],
}
ALL_WANTED = {r for rs in ROUTINES.values() for r in rs}
INTRINSICS = {"max": "fmax", "min": "fmin", "abs": "fabs", "sqrt": "sqrt", "sin": "sin", "cos": "cos", "exp": "exp",
              "atan": "atan"}
EXTERNAL_REAL_FUNCS = {"initialconditionatpoint"}


# ----------------------------------------------------------------------------- source reading
def logical_lines(path):
    """fixed-form F77 -> list of lower-cased statements with continuations joined, comments dropped"""
    out = []
    for raw in open(path, errors="replace"):
        line = raw.rstrip("\n")
        if not line.strip():
            continue
        if line[0] in "cC*!":
            continue
        # strip inline comments (the routines we translate hold no character literals except include)
        if "!" in line:
            line = line[: line.index("!")]
        if not line.strip():
            continue
        line = line.expandtabs(8)
        is_cont = len(line) > 5 and line[:5].strip() == "" and line[5] not in " 0"
        body = line[6:72] if len(line) > 6 else ""
        if is_cont and out:
            # blanks are insignificant in fixed form: literals may be split across the line break
            out[-1] += body.strip()
        else:
            out.append(body.strip())
    # a trailing `;` (an empty second statement, TwoSpecies_*TZSourceF.f:60-61) is not part of the statement
    return [s.lower().rstrip().rstrip(";").rstrip() for s in out]


def split_routines(stmts):
    routines, cur, name = {}, None, None
    for s in stmts:
        m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)\s*$", s)
        if m:
            name, cur = m.group(1), [s]
            continue
        if cur is not None:
            cur.append(s)
            if s == "end":
                routines[name] = cur
                cur = None
    return routines


def split_top(s, sep=","):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


# ----------------------------------------------------------------------------- expression parser
TOK = re.compile(r"\s*(\.[a-z]+\.|\*\*|[0-9]+\.(?![a-z]+\.)[0-9]*(?:[ed][+-]?[0-9]+)?|\.[0-9]+(?:[ed][+-]?[0-9]+)?|[0-9]+(?:[ed][+-]?[0-9]+)?|[a-z_]\w*|[-+*/(),=])")


def tokenize(s):
    toks, pos = [], 0
    s = s.strip()
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            raise SyntaxError("cannot tokenize %r at %r" % (s, s[pos:]))
        toks.append(m.group(1))
        pos = m.end()
    return toks


class Ctx:
    def __init__(self):
        self.dummies = []      # argument names in order
        self.types = {}        # name -> 'int' | 'int64_t' | 'double'
        self.dims = {}         # array name -> list of (lo_expr_tokens|None, hi_expr_tokens|'*')
        self.funcs = set()     # external functions declared as typed scalars but called


class Parser:
    """recursive descent over Fortran expression tokens, emitting fully parenthesised C"""

    def __init__(self, toks, ctx):
        self.t, self.i, self.c = toks, 0, ctx

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self, expect=None):
        tok = self.peek()
        if expect is not None and tok != expect:
            raise SyntaxError("expected %r got %r in %r" % (expect, tok, self.t))
        self.i += 1
        return tok

    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.peek() == ".or.":
            self.take()
            a = "(%s || %s)" % (a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek() == ".and.":
            self.take()
            a = "(%s && %s)" % (a, self.p_not())
        return a

    def p_not(self):
        if self.peek() == ".not.":
            self.take()
            return "(!%s)" % self.p_not()
        return self.p_rel()

    REL = {".eq.": "==", ".ne.": "!=", ".gt.": ">", ".ge.": ">=", ".lt.": "<", ".le.": "<="}

    def p_rel(self):
        a = self.p_add()
        if self.peek() in self.REL:
            op = self.REL[self.take()]
            a = "(%s %s %s)" % (a, op, self.p_add())
        return a

    def p_add(self):
        if self.peek() in ("-", "+"):
            sign = self.take()
            a = self.p_mul()
            a = "(-%s)" % a if sign == "-" else a
        else:
            a = self.p_mul()
        while self.peek() in ("+", "-"):
            op = self.take()
            a = "(%s %s %s)" % (a, op, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek() in ("*", "/"):
            op = self.take()
            a = "(%s %s %s)" % (a, op, self.p_pow())
        return a

    def p_pow(self):
        base = self.p_primary()
        if self.peek() == "**":
            self.take()
            # exponent: unary-signed power expression (right associative)
            if self.peek() in ("-", "+"):
                raise SyntaxError("signed exponent unsupported")
            ex = self.p_pow()
            m = re.fullmatch(r"\(?([0-9]+)\)?", ex)
            if m and 2 <= int(m.group(1)) <= 8:
                return "f77_powi%d(%s)" % (int(m.group(1)), base)
            return "pow(%s, %s)" % (base, ex)
        return base

    def p_primary(self):
        tok = self.take()
        if tok == "(":
            a = self.expr()
            self.take(")")
            return a   # already parenthesised by the binary rules; keep grouping explicit
        if tok in (".true.", ".false."):
            return "1" if tok == ".true." else "0"
        if re.match(r"[0-9.]", tok):
            return self.number(tok)
        if re.match(r"[a-z_]", tok):
            if self.peek() == "(":
                self.take("(")
                args = []
                if self.peek() != ")":
                    args.append(self.arg())
                    while self.peek() == ",":
                        self.take()
                        args.append(self.arg())
                self.take(")")
                return self.call_or_index(tok, args)
            return self.var(tok)
        raise SyntaxError("unexpected token %r in %r" % (tok, self.t))

    def arg(self):
        """returns (c_expr, raw_tokens) so that call sites can decide how to pass by reference"""
        start = self.i
        e = self.expr()
        return (e, self.t[start:self.i])

    @staticmethod
    def number(tok):
        tok = tok.replace("d", "e")
        if re.fullmatch(r"[0-9]+", tok):
            return tok
        return tok

    def var(self, name):
        c = self.c
        if name in c.dims:
            return name            # whole array -> pointer
        if name in c.dummies:
            return "(*%s)" % name
        return name

    def call_or_index(self, name, args):
        c = self.c
        if name in c.dims:
            return "%s[%s]" % (name, index_expr(c, name, [a[0] for a in args]))
        if name in INTRINSICS:
            return "%s(%s)" % (INTRINSICS[name], ", ".join(a[0] for a in args))
        if name in EXTERNAL_REAL_FUNCS or name in c.funcs:
            return "%s_(%s)" % (name, ", ".join(by_ref(c, a) for a in args))
        raise SyntaxError("unknown function or array %r" % name)


def by_ref(c, arg):
    """Fortran passes everything by reference"""
    e, raw = arg
    if len(raw) == 1 and re.match(r"[a-z_]", raw[0]):
        n = raw[0]
        if n in c.dims or n in c.dummies:
            return n
        return "&%s" % n
    if raw and raw[0] in c.dims and raw[1] == "(":
        return "&%s" % e
    raise SyntaxError("cannot pass expression by reference: %r" % (raw,))


def index_expr(c, name, idx):
    dims = c.dims[name]
    if len(idx) != len(dims):
        raise SyntaxError("rank mismatch for %s" % name)
    # column major: off = (i1-lo1) + ext1*((i2-lo2) + ext2*(...))
    expr = None
    for k in reversed(range(len(dims))):
        lo, hi = dims[k]
        term = "(%s - %s)" % (idx[k], lo)
        if expr is None:
            expr = term
        else:
            expr = "(%s + (int64_t)(%s - %s + 1) * %s)" % (term, hi, lo, expr)
    return expr


def cexpr(s, ctx):
    p = Parser(tokenize(s), ctx)
    e = p.expr()
    if p.peek() is not None:
        raise SyntaxError("trailing tokens in %r" % s)
    return e


# ----------------------------------------------------------------------------- statements
DECL = re.compile(r"^(integer\*8|integer|real|double precision)\s+(.*)$")
CTYPE = {"integer*8": "int64_t", "integer": "int", "real": "double", "double precision": "double"}


def translate(name, stmts):
    ctx = Ctx()
    m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)\s*$", stmts[0])
    ctx.dummies = [a.strip() for a in split_top(m.group(2))]
    body = []
    decl_raw = []
    for s in stmts[1:]:
        if s in ("implicit none",) or s.startswith("include"):
            continue
        d = DECL.match(s)
        if d:
            decl_raw.append((CTYPE[d.group(1)], d.group(2)))
            continue
        body.append(s)
    # first pass: names and types (bounds may reference later-declared scalars, so parse dims after)
    pending_dims = {}
    for ctype, rest in decl_raw:
        for item in split_top(rest):
            mm = re.match(r"^(\w+)\s*(?:\((.*)\))?$", item)
            n = mm.group(1)
            ctx.types[n] = ctype
            if mm.group(2) is not None:
                pending_dims[n] = mm.group(2)
                ctx.dims[n] = None
    for n in list(ctx.types):
        if n in EXTERNAL_REAL_FUNCS:
            ctx.funcs.add(n)
    for n, spec in pending_dims.items():
        dims = []
        for dspec in split_top(spec):
            if ":" in dspec:
                lo, hi = dspec.split(":")
            else:
                lo, hi = "1", dspec
            lo_c = cexpr(lo, ctx)
            hi_c = "0" if hi.strip() == "*" else cexpr(hi, ctx)
            dims.append((lo_c, hi_c))
        ctx.dims[n] = dims
    out = []
    params = []
    for a in ctx.dummies:
        t = ctx.types.get(a)
        if t is None:
            raise SyntaxError("%s: untyped dummy %s" % (name, a))
        params.append("%s* restrict %s" % (t, a) if a in ctx.dims else "%s* %s" % (t, a))
    out.append("void %s_(%s) {" % (name, ", ".join(params)))
    for n, t in ctx.types.items():
        if n in ctx.dummies or n in ctx.funcs:
            continue
        if n in ctx.dims:
            # local array with constant bounds (e.g. eCoeffs(1:6,1:6)): a flat, zero-initialised C array,
            # indexed column-major by index_expr like the dummy arrays
            size = " * ".join("(%s - %s + 1)" % (hi, lo) for lo, hi in ctx.dims[n])
            try:
                count = int(eval(size, {"__builtins__": {}}))
            except Exception:
                raise SyntaxError("%s: local array %s needs constant bounds" % (name, n))
            out.append("  %s %s[%d] = {0};" % (t, n, count))
            continue
        out.append("  %s %s = 0;" % (t, n))
    ind = 1

    def emit(line):
        out.append("  " * ind + line)

    def stmt(s):
        nonlocal ind
        if s in ("return",):
            emit("return;")
            return
        mm = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s)
        if mm and not re.match(r"^do\s*\w+\s*=\s*[^,]*$", s):
            var, rng = mm.group(1), split_top(mm.group(2))
            lo, hi = cexpr(rng[0], ctx), cexpr(rng[1], ctx)
            v = Parser([var], ctx).var(var)
            if len(rng) == 3:
                raise SyntaxError("do-loop stride unsupported")
            # Fortran evaluates the bounds once, before the loop
            emit("{ const int lo__ = %s, hi__ = %s; for (%s = lo__; %s <= hi__; ++%s) {" % (lo, hi, v, v, v))
            ind += 1
            return
        if s in ("end do", "enddo"):
            ind -= 1
            emit("} }")
            return
        mm = re.match(r"^if\s*\((.*)\)\s*then$", s)
        if mm:
            emit("if (%s) {" % cexpr(mm.group(1), ctx))
            ind += 1
            return
        mm = re.match(r"^else\s*if\s*\((.*)\)\s*then$", s)
        if mm:
            ind -= 1
            emit("} else if (%s) {" % cexpr(mm.group(1), ctx))
            ind += 1
            return
        if s == "else":
            ind -= 1
            emit("} else {")
            ind += 1
            return
        if s in ("end if", "endif"):
            ind -= 1
            emit("}")
            return
        if s.startswith("if"):
            # one-line if: find the balanced condition
            p0 = s.index("(")
            depth, p = 0, p0
            while True:
                if s[p] == "(":
                    depth += 1
                elif s[p] == ")":
                    depth -= 1
                    if depth == 0:
                        break
                p += 1
            cond, rest = s[p0 + 1:p], s[p + 1:].strip()
            emit("if (%s) {" % cexpr(cond, ctx))
            ind += 1
            stmt(rest)
            ind -= 1
            emit("}")
            return
        mm = re.match(r"^call\s+(\w+)\s*\((.*)\)$", s)
        if mm:
            if mm.group(1) not in ALL_WANTED:
                # a routine outside the path (the reference only reaches these under `if (.false.)`: limiter and
                # artificial-viscosity experiments): not translated, and loud if it were ever reached
                emit('f77_untranslated("%s");' % mm.group(1))
                return
            args, temps = [], []
            for a in split_top(mm.group(2)):
                toks = tokenize(a)
                p = Parser(toks, ctx)
                e = p.expr()
                try:
                    args.append(by_ref(ctx, (e, toks)))
                except SyntaxError:
                    # an expression or a literal actual argument: Fortran passes the address of a temporary
                    # (integer arithmetic on the box bounds and literal direction numbers in the routines we translate)
                    if any(re.match(r"^\d*\.\d*|\d+[de]", t) for t in toks):
                        raise
                    t = "_arg%d" % len(temps)
                    temps.append("int %s = %s;" % (t, e))
                    args.append("&" + t)
            if temps:
                emit("{ " + " ".join(temps))
                emit("  %s_(%s); }" % (mm.group(1), ", ".join(args)))
            else:
                emit("%s_(%s);" % (mm.group(1), ", ".join(args)))
            return
        # assignment: split at the top-level '='
        depth = 0
        for k, ch in enumerate(s):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                lhs, rhs = s[:k].strip(), s[k + 1:].strip()
                if lhs in ctx.dims:
                    # whole-array assignment of a scalar (rN = 0.0): every element of the declared extent
                    size = " * ".join("(int64_t)(%s - %s + 1)" % (hi, lo) for lo, hi in ctx.dims[lhs])
                    emit("{ const int64_t n__ = %s; for (int64_t k__ = 0; k__ < n__; ++k__) %s[k__] = %s; }"
                         % (size, lhs, cexpr(rhs, ctx)))
                    return
                emit("%s = %s;" % (cexpr(lhs, ctx), cexpr(rhs, ctx)))
                return
        raise SyntaxError("%s: cannot translate statement %r" % (name, s))

    for s in body:
        if s == "end":
            break
        stmt(s)
    out.append("}")
    return "\n".join(out), params


PRELUDE = """/* GENERATED by oracle/f77toc.py from the reference Fortran -- do not commit, do not edit. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
static inline void f77_untranslated(const char* name) { fprintf(stderr, "f77toc: %s is not translated\\n", name); abort(); }
static inline double f77_powi2(double x) { return x * x; }
static inline double f77_powi3(double x) { return (x * x) * x; }
static inline double f77_powi4(double x) { double t = x * x; return t * t; }
static inline double f77_powi5(double x) { double t = x * x; return (t * x) * t; }
static inline double f77_powi6(double x) { double t = (x * x) * x; return t * t; }
static inline double f77_powi7(double x) { double t = x * x; double u = t * x; return (u * u) * x; }
static inline double f77_powi8(double x) { double t = x * x; t = t * t; return t * t; }
double initialconditionatpoint_(int64_t* ic, int* i1, int* i2, int* i3, int* i4);
"""


def main():
    ref, outdir = sys.argv[1], sys.argv[2]
    os.makedirs(outdir, exist_ok=True)
    protos = []
    for fname, wanted in ROUTINES.items():
        routines = split_routines(logical_lines(os.path.join(ref, fname)))
        chunks = []
        for r in wanted:
            if r not in routines:
                raise SystemExit("routine %s not found in %s" % (r, fname))
            code, params = translate(r, routines[r])
            protos.append("void %s_(%s);" % (r, ", ".join(params)))
            chunks.append(code)
        with open(os.path.join(outdir, os.path.splitext(fname)[0] + "_f77.c"), "w") as fh:
            fh.write(PRELUDE)
            fh.write('#include "loki_ref_protos.h"\n\n')
            fh.write("\n\n".join(chunks))
            fh.write("\n")
    with open(os.path.join(outdir, "loki_ref_protos.h"), "w") as fh:
        fh.write("/* GENERATED by oracle/f77toc.py */\n#include <stdint.h>\n" + "\n".join(protos) + "\n")
    print("f77toc: wrote %d routines into %s" % (len(protos), outdir))


if __name__ == "__main__":
    main()

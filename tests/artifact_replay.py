"""Replays horde-ad's PRINTED gradient artifacts -- the text `printArtifactPretty` emits, e.g. the complete
MnistCnnShaped2 gradient program at test/simplified/TestMnistPP.hs:728-735 -- against any carrier backend
(the CPU oracle, the device carrier, or the device carrier while a lazy program is being recorded).

The printed artifact IS the op program horde-ad's interpreter sends to a backend (AstInterpret.hs:69-366: one class
method per AST node), so parsing and evaluating it is the closest thing to running the reference's own output
through the C ABI without GHC.  Test infrastructure: a recursive-descent parser for the pretty
printer's expression language (let / lambda / type applications / infix + * !$ `sindex0` <=.) and an evaluator that
maps every AST name to the carrier method of the same name.

`sscatter @shm t f` does not print its destination shape; it is recovered from the forward `sgather` that carries
the same index function (the adjoint pairs of DeltaEval.hs:538-562), which the evaluator records as it goes.
"""
from __future__ import annotations

import re

import numpy as np

from horde_ad_b200 import ix as ixm

TOKEN = re.compile(r"\s*(?:(\d+\.\d+(?:e-?\d+)?|\d+)|([A-Za-z_][A-Za-z0-9_']*)|(->|<=\.|!\$|[()\[\],;=\\@+*`-]))")


def tokenize(s: str):
    out, pos = [], 0
    s = s.strip()
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {s[pos:pos + 40]!r}")
        num, ident, sym = m.groups()
        if num is not None:
            out.append(("num", float(num) if "." in num or "e" in num else int(num)))
        elif ident is not None:
            out.append(("id", ident))
        else:
            out.append(("sym", sym))
        pos = m.end()
    return out


class Parser:
    """expr   := 'let' binds 'in' expr | infix
       infix  := precedence climbing over  <=. (4)  + (6)  * (7)  !$ `f` (9)
       app    := atom atom*            (juxtaposition; '@' type arguments are atoms)
       atom   := num | id | '(' expr ')' | '(' '\\' '[' ids ']' '->' '[' exprs ']' ')' | '[' exprs ']' | '@' atom"""

    PREC = {"<=.": 4, "+": 6, "*": 7, "!$": 9, "`": 9}

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None)

    def eat(self, kind=None, val=None):
        tok = self.peek()
        if (kind and tok[0] != kind) or (val is not None and tok[1] != val):
            raise SyntaxError(f"expected {kind} {val}, got {tok} at token {self.i}")
        self.i += 1
        return tok

    def top(self):
        # \dret u1 -> body   |   \u1 -> body
        self.eat("sym", "\\")
        params = []
        while self.peek() != ("sym", "->"):
            params.append(self.eat("id")[1])
        self.eat("sym", "->")
        body = self.expr()
        if self.peek()[0] != "eof":
            raise SyntaxError(f"trailing tokens at {self.i}: {self.t[self.i:self.i + 5]}")
        return params, body

    def expr(self):
        if self.peek() == ("id", "let"):
            self.eat()
            binds = []
            while True:
                name = self.eat("id")[1]
                self.eat("sym", "=")
                binds.append((name, self.expr()))
                if self.peek() == ("sym", ";"):
                    self.eat()
                    continue
                break
            self.eat("id", "in")
            return ("let", binds, self.expr())
        return self.infix(0)

    def infix(self, minp):
        lhs = self.app()
        while True:
            tok = self.peek()
            if tok[0] != "sym" or tok[1] not in self.PREC or self.PREC[tok[1]] < minp:
                return lhs
            op = tok[1]
            self.eat()
            if op == "`":
                op = self.eat("id")[1]
                self.eat("sym", "`")
                p = 9
            else:
                p = self.PREC[op]
            rhs = self.infix(p + 1)
            lhs = ("bin", op, lhs, rhs)

    def starts_atom(self):
        k, v = self.peek()
        if k in ("num",):
            return True
        if k == "id":
            return v not in ("in", "let")
        return k == "sym" and v in ("(", "[", "@")

    def app(self):
        f = self.atom()
        args = []
        while self.starts_atom():
            args.append(self.atom())
        return ("app", f, args) if args else f

    def atom(self):
        k, v = self.peek()
        if k == "num":
            self.eat()
            return ("num", v)
        if k == "id":
            self.eat()
            return ("var", v)
        if (k, v) == ("sym", "@"):
            self.eat()
            return ("targ", self.atom())
        if (k, v) == ("sym", "["):
            self.eat()
            items = []
            while self.peek() != ("sym", "]"):
                if self.peek() == ("sym", "-"):      # negative literal inside a list
                    self.eat()
                    items.append(("num", -self.eat("num")[1]))
                else:
                    items.append(self.expr())
                if self.peek() == ("sym", ","):
                    self.eat()
            self.eat("sym", "]")
            return ("list", items)
        if (k, v) == ("sym", "("):
            self.eat()
            if self.peek() == ("sym", "\\"):
                self.eat()
                vars_ = self.atom()
                self.eat("sym", "->")
                body = self.atom()
                self.eat("sym", ")")
                return ("lam", [x[1] for x in vars_[1]], body[1], self._text_of(body))
            e = self.expr()
            self.eat("sym", ")")
            return e
        raise SyntaxError(f"unexpected {self.peek()} at token {self.i}")

    def _text_of(self, node):
        return repr(node)


def parse(text: str):
    return Parser(tokenize(text)).top()


def _canon_lambda(vars_, bodies):
    """Index-function text with its variables renamed v0.. (so that a scatter finds the gather it is the adjoint of)."""
    ren = {v: f"v{k}" for k, v in enumerate(vars_)}

    def go(n):
        if n[0] == "var":
            return ("var", ren.get(n[1], n[1]))
        if n[0] in ("num",):
            return n
        if n[0] == "targ":
            return ("targ", go(n[1]))
        if n[0] == "list":
            return ("list", tuple(go(x) for x in n[1]))
        if n[0] == "app":
            return ("app", go(n[1]), tuple(go(x) for x in n[2]))
        if n[0] == "bin":
            return ("bin", n[1], go(n[2]), go(n[3]))
        raise ValueError(n[0])
    return repr(tuple(go(b) for b in bodies))


class Evaluator:
    def __init__(self, T, dtype=np.float64):
        self.T = T
        self.dtype = dtype
        self.gather_shp = {}     # canonical index function -> destination/source leading shape
        self.n_calls = 0

    # ------------------------------------------------------------ helpers
    def _ints(self, node, env):
        v = self.ev(node, env, False)
        return [int(x) for x in v] if isinstance(v, list) else int(v)

    def _tensor(self, v):
        if isinstance(v, (int, float)):
            return self.T.sscalar(float(v), self.dtype)
        return v

    def _index_fn(self, lam, env):
        _, vars_, bodies, _ = lam

        def f(ixs):
            e2 = dict(env)
            for name, s in zip(vars_, ixs):
                e2[name] = s
            return [self.ev(b, e2, True) for b in bodies]
        return f, len(bodies), _canon_lambda(vars_, bodies)

    # ---------------------------------------------------------------- eval
    def ev(self, n, env, ixmode):
        k = n[0]
        if k == "num":
            return n[1]
        if k == "var":
            if n[1] not in env:
                return self.apply(n[1], [], [], env, ixmode)
            return env[n[1]]
        if k == "list":
            return [self.ev(x, env, ixmode) for x in n[1]]
        if k == "let":
            e2 = dict(env)
            for name, e in n[1]:
                e2[name] = self.ev(e, e2, ixmode)
            return self.ev(n[2], e2, ixmode)
        if k == "lam":
            return n
        if k == "bin":
            return self.binop(n[1], n[2], n[3], env, ixmode)
        if k == "app":
            f = n[1]
            if f[0] != "var":
                raise ValueError(f"application of a non-name: {f}")
            targs = [a[1] for a in n[2] if a[0] == "targ"]
            args = [a for a in n[2] if a[0] != "targ"]
            return self.apply(f[1], targs, args, env, ixmode)
        raise ValueError(k)

    def binop(self, op, a, b, env, ixmode):
        T = self.T
        if op in ("!$", "sindex0"):
            t = self.ev(a, env, ixmode)
            ixs = self.ev(b, env, ixmode)
            if ixmode:
                if op == "sindex0":
                    return T.ix_tbl(t, ixs)
                return ("indexed", t, ixs)            # consumed by kargMax
            return T.sindex(t, [int(i) for i in ixs])
        x, y = self.ev(a, env, ixmode), self.ev(b, env, ixmode)
        if ixmode:
            if op == "+":
                return x + y
            if op == "*":
                return x * y
            if op == "<=.":
                return ixm.Ix.lift(x) <= y
            raise ValueError(op)
        self.n_calls += 1
        if op == "+":
            return T.binary("add", self._tensor(x), self._tensor(y))
        if op == "*":
            return T.binary("mul", self._tensor(x), self._tensor(y))
        raise ValueError(op)

    def apply(self, name, targs, args, env, ixmode):
        T = self.T
        A = lambda i: self.ev(args[i], env, ixmode)
        if ixmode:
            if name == "kargMax":
                v = A(0)
                assert v[0] == "indexed"
                return T.ix_argmax(v[1], v[2])
            if name == "ifH":
                return ixm.ifH(A(0), A(1), A(2))
            if name == "negate":
                return -A(0)
            if name in ("splainPart", "sfromPlain"):
                return A(0)
            if name == "sconcrete":
                return self.ev(args[0], env, False)
            if name == "sfromListLinear":
                return self.apply(name, targs, args, env, False)
            raise ValueError(f"index expression: {name}")
        self.n_calls += 1
        if name in ("sfromPlain", "splainPart", "sconcrete", "sfromPrimal", "sprimalPart"):
            return A(0)
        if name == "tproject1":
            return A(0)[0]
        if name == "tproject2":
            return A(0)[1]
        if name == "tpair":
            return (A(0), A(1))
        if name == "str":
            return T.stranspose([1, 0], A(0))
        if name == "stranspose":
            return T.stranspose(self._ints(targs[0], env), A(0))
        if name == "sreplicate":
            if not targs:        # the Nested constant `sreplicate [11] 2` inside an sconcrete
                sh = [int(x) for x in self.ev(args[0], env, False)]
                v = self.ev(args[1], env, False)
                return T.sconcrete(np.full(sh, v, dtype=self.dtype if isinstance(v, float) else np.int64))
            return T.sreplicate(self._ints(targs[0], env), self._tensor(A(0)))
        if name in ("sreplicateN", "sreplicate0N"):
            return T.sreplicateN(self._ints(targs[0], env), self._tensor(A(0)))
        if name == "sreshape":
            return T.sreshape(self._ints(targs[0], env), A(0))
        if name == "ssum":
            t = A(0)
            assert T.shape(t)[0] == self._ints(targs[0], env)
            return T.ssum(t)
        if name == "ssumN":
            return T.ssumN(A(0), len(self._ints(targs[0], env)))
        if name == "sdot1In":
            return T.sdot1In(A(0), A(1))
        if name == "smatmul2":
            return T.smatmul2(A(0), A(1))
        if name == "negate":
            return T.unary("negate", A(0))
        if name == "SNat":
            return self._ints(targs[0], env)
        if name == "siota":
            return T.siota(int(A(0)), np.int64)
        if name == "sfromListLinear":
            sh = [int(x) for x in self.ev(args[0], env, False)]
            vals = self.ev(args[1], env, False)
            isf = any(isinstance(v, float) for v in vals)
            return T.sconcrete(np.array(vals, dtype=self.dtype if isf else np.int64).reshape(sh))
        if name == "soneHot":
            # the printed form omits the one-hot axes' extents; the artifacts at hand only index unit axes (`[0]`)
            ixs = [int(i) for i in self.ev(args[1], env, False)]
            if any(ixs):
                raise ValueError("soneHot at a non-zero index: the extent of the new axis is not printed")
            return T.soneHot([1] * len(ixs), A(0), ixs)
        if name == "sgather":
            shm = self._ints(targs[0], env)
            src = A(0)
            f, sp, key = self._index_fn(args[1], env)
            self.gather_shp[key] = list(T.shape(src)[:sp])
            return T.sgather(shm, src, f, sp)
        if name == "sscatter":
            shm = self._ints(targs[0], env)
            src = A(0)
            f, sp, key = self._index_fn(args[1], env)
            if key not in self.gather_shp:
                raise KeyError("sscatter without a forward sgather of the same index function: destination shape unknown")
            return T.sscatter(self.gather_shp[key], src, f, len(shm))
        raise ValueError(f"artifact primitive not handled: {name}")


def run_artifact(T, text: str, inputs: dict, dtype=np.float64):
    """Evaluate a printed artifact `\\x y -> body`; `inputs` maps the lambda's parameter names to values (tensors of the
    backend, nested pairs as Python tuples).  Returns the value (nested tuples for tpair)."""
    params, body = parse(text)
    ev = Evaluator(T, dtype)
    env = {p: inputs[p] for p in params}
    out = ev.ev(body, env, False)
    return out, ev

"""Symbolic index scalars: the device carrier's `IntOf target`.

horde-ad hands gather/scatter their index functions as closures
`IxSOf target shm -> IxSOf target shp` (Ops.hs:689-705).  A carrier whose
index scalar is *symbolic* can reify such a closure by applying it to fresh
variables -- exactly what the reference's own AstTensor instance does
(OpsAst.hs:474-485, AstFreshId.hs:153-166).  `Ix` is that symbolic scalar
(grammar: SURVEY.md appendix B = Ast.hs:487-528,575-579,654-665) and
`compile_program` lowers the resulting expression DAG to the register
bytecode of include/hordeb200.h (`hb_insn`).

The same expression type is the input language of the CPU oracle's index
evaluator, which interprets it independently with numpy.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

from . import ffi

# node kinds --------------------------------------------------------------
# ty: 'i' (Int), 'f' (real), 'b' (Bool)


class Ix:
    __slots__ = ("op", "args", "ty", "val", "tensor")

    def __init__(self, op, args=(), ty="i", val=None, tensor=None):
        self.op = op
        self.args = tuple(args)
        self.ty = ty
        self.val = val
        self.tensor = tensor

    # -- construction helpers
    @staticmethod
    def var(k: int) -> "Ix":
        return Ix("var", (), "i", k)

    @staticmethod
    def lift(x) -> "Ix":
        if isinstance(x, Ix):
            return x
        if isinstance(x, bool):
            return Ix("const", (), "b", int(x))
        if isinstance(x, int):
            return Ix("const", (), "i", int(x))
        if isinstance(x, float):
            return Ix("const", (), "f", float(x))
        try:  # numpy scalars
            import numpy as np
            if isinstance(x, np.integer):
                return Ix("const", (), "i", int(x))
            if isinstance(x, np.floating):
                return Ix("const", (), "f", float(x))
        except ImportError:  # pragma: no cover
            pass
        raise TypeError(f"cannot lift {type(x)} to an index expression")

    def _bin(self, op, other, rty=None):
        o = Ix.lift(other)
        a, b = self, o
        if a.ty != b.ty:
            if {a.ty, b.ty} == {"i", "f"}:  # numeric literal promotion
                a = a if a.ty == "f" else Ix("i2f", (a,), "f")
                b = b if b.ty == "f" else Ix("i2f", (b,), "f")
            else:
                raise TypeError(f"index expression type mismatch: {a.ty} vs {b.ty}")
        return Ix(op, (a, b), rty or a.ty)

    # Num / IntegralH
    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return Ix.lift(o)._bin("add", self)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return Ix.lift(o)._bin("sub", self)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return Ix.lift(o)._bin("mul", self)
    def __neg__(self): return Ix("neg", (self,), self.ty)
    def __abs__(self): return Ix("abs", (self,), self.ty)
    def signum(self): return Ix("signum", (self,), self.ty)
    def quotH(self, o): return self._bin("quot", o)
    def remH(self, o): return self._bin("rem", o)
    def __truediv__(self, o): return self._bin("div", o)
    # EqH / OrdH  (<=. etc.)
    def __le__(self, o): return self._bin("leq", o, "b")
    def __lt__(self, o): return self._bin("lt", o, "b")
    def __ge__(self, o): return Ix.lift(o)._bin("leq", self, "b")
    def __gt__(self, o): return Ix.lift(o)._bin("lt", self, "b")
    def eq(self, o): return self._bin("eq", o, "b")
    def neq(self, o): return Ix("not", (self._bin("eq", o, "b"),), "b")
    # Boolean
    def __and__(self, o): return Ix("and", (self, Ix.lift(o)), "b")
    def __or__(self, o): return Ix("or", (self, Ix.lift(o)), "b")
    def __invert__(self): return Ix("not", (self,), "b")

    def __repr__(self):
        if self.op == "var":
            return f"i{self.val}"
        if self.op == "const":
            return repr(self.val)
        if self.op in ("tbl", "argmax", "argmin"):
            return f"{self.op}(<tensor>, {list(self.args)})"
        return f"{self.op}({', '.join(map(repr, self.args))})"


def ifH(c, a, b) -> Ix:
    c, a, b = Ix.lift(c), Ix.lift(a), Ix.lift(b)
    if a.ty != b.ty:
        raise TypeError("ifH: branches differ in type")
    return Ix("cond", (c, a, b), a.ty)


def minH(a, b): return Ix.lift(a)._bin("min", b)
def maxH(a, b): return Ix.lift(a)._bin("max", b)
def fromIntegral(a) -> Ix: return Ix("i2f", (Ix.lift(a),), "f")
def floorH(a) -> Ix: return Ix("floor", (Ix.lift(a),), "i")
def fun1(name: str, a) -> Ix: return Ix("fun1", (Ix.lift(a),), "f", name)


def tbl(tensor, ix: Sequence, is_float: bool) -> Ix:
    """`tensor sindex0 [ix...]` inside an index program (AstIndexK)."""
    return Ix("tbl", tuple(Ix.lift(i) for i in ix), "f" if is_float else "i", None, tensor)


def argmax(tensor, ix: Sequence) -> Ix:
    """`kargMax (tensor !$ [ix...])` inside an index program (AstArgMaxK)."""
    return Ix("argmax", tuple(Ix.lift(i) for i in ix), "i", None, tensor)


def argmin(tensor, ix: Sequence) -> Ix:
    return Ix("argmin", tuple(Ix.lift(i) for i in ix), "i", None, tensor)


def map_input(k: int, is_float: bool) -> Ix:
    """k-th input of a fused elementwise map."""
    return Ix("in", (), "f" if is_float else "i", k)


def reify(f: Callable, n_vars: int) -> list[Ix]:
    """Apply an index closure to fresh variables (OpsAst.hs:474-485)."""
    out = f([Ix.var(k) for k in range(n_vars)])
    return [Ix.lift(o) for o in out]


# ---------------------------------------------------------------- compiler
_IBIN = {"add": "IADD", "sub": "ISUB", "mul": "IMUL", "quot": "IQUOT", "rem": "IREM",
         "min": "IMIN", "max": "IMAX", "leq": "ILEQ", "lt": "ILT", "eq": "IEQ"}
_FBIN = {"add": "FADD", "sub": "FSUB", "mul": "FMUL", "div": "FDIV", "pow": "FPOW",
         "logbase": "FLOGBASE", "atan2": "FATAN2", "min": "FMIN", "max": "FMAX",
         "leq": "FLEQ", "lt": "FLT", "eq": "FEQ"}
_IUN = {"neg": "INEG", "abs": "IABS", "signum": "ISIGNUM"}


class Program:
    """Compiled program: instruction array + the tensor operands it reads."""

    def __init__(self, insns, tensors, n_vars, n_out):
        self.insns = insns
        self.tensors = tensors
        self.n_vars = n_vars
        self.n_out = n_out


def compile_program(outs: Sequence[Ix], n_vars: int) -> Program:
    """Lower expression DAGs to hb_insn bytecode with register allocation.

    Common subexpressions (shared Python objects, = AstLet-shared indices) are
    evaluated once.  Registers are recycled after the last use.
    """
    order: list[Ix] = []
    uses: dict[int, int] = {}
    seen: dict[int, Ix] = {}

    def visit(e: Ix):
        stack = [(e, False)]
        while stack:
            node, done = stack.pop()
            if done:
                order.append(node)
                continue
            if id(node) in seen:
                continue
            seen[id(node)] = node
            stack.append((node, True))
            for a in node.args:
                uses[id(a)] = uses.get(id(a), 0) + 1
                if id(a) not in seen:
                    stack.append((a, False))

    for o in outs:
        uses[id(o)] = uses.get(id(o), 0) + 1
        visit(o)

    insns: list[ffi.Insn] = []
    tensors: list = []
    reg_of: dict[int, int] = {}
    free = list(range(32 - 1, -1, -1))
    remaining = dict(uses)

    def tensor_slot(t):
        for i, u in enumerate(tensors):
            if u is t:
                return i
        tensors.append(t)
        return len(tensors) - 1

    def emit(op, dst=0, a=0, b=0, c=0, n=0, imm_i=None, imm_f=None, regs=None):
        ins = ffi.Insn()
        ins.op = ffi.VMOP[op]
        ins.dst, ins.a, ins.b, ins.c, ins.n = dst, a, b, c, n
        if imm_i is not None:
            ins.imm.i = imm_i
        elif imm_f is not None:
            ins.imm.f = imm_f
        elif regs is not None:
            for k, r in enumerate(regs):
                ins.imm.regs[k] = r
        insns.append(ins)

    def release(e: Ix):
        remaining[id(e)] -= 1
        if remaining[id(e)] == 0:
            free.append(reg_of[id(e)])

    for node in order:
        srcs = [reg_of[id(a)] for a in node.args]
        # operands may be released before the destination is allocated
        # (dst may alias a dying source: every op reads before it writes)
        for a in node.args:
            release(a)
        if not free:
            raise ffi.HordeError("index program needs more than 32 live registers")
        dst = free.pop()
        reg_of[id(node)] = dst
        op = node.op
        if op == "var":
            emit("IVAR", dst, a=node.val)
        elif op == "const":
            if node.ty == "f":
                emit("FCONST", dst, imm_f=node.val)
            else:
                emit("ICONST", dst, imm_i=node.val)
        elif op == "in":
            emit("IN", dst, a=node.val)
        elif op in ("neg", "abs", "signum"):
            if node.ty == "f":
                emit("FUN1", dst, a=srcs[0], b=ffi.UNOP[{"neg": "negate"}.get(op, op)])
            else:
                emit(_IUN[op], dst, a=srcs[0])
        elif op == "fun1":
            emit("FUN1", dst, a=srcs[0], b=ffi.UNOP[node.val])
        elif op in ("and", "or"):
            emit("BAND" if op == "and" else "BOR", dst, a=srcs[0], b=srcs[1])
        elif op == "not":
            emit("BNOT", dst, a=srcs[0])
        elif op == "cond":
            emit("SEL", dst, a=srcs[0], b=srcs[1], c=srcs[2])
        elif op == "i2f":
            emit("I2F", dst, a=srcs[0])
        elif op == "floor":
            emit("FLOOR", dst, a=srcs[0])
        elif op in ("tbl", "argmax", "argmin"):
            if len(srcs) > 8:
                raise ffi.HordeError("index program: tensor operand of rank > 8")
            emit({"tbl": "TBL", "argmax": "ARGMAX", "argmin": "ARGMIN"}[op], dst,
                 a=tensor_slot(node.tensor), n=len(srcs), regs=srcs)
        else:
            aty = node.args[0].ty
            table = _FBIN if aty == "f" else _IBIN
            if op not in table:
                raise ffi.HordeError(f"index program: op {op} on type {aty}")
            emit(table[op], dst, a=srcs[0], b=srcs[1])
    for k, o in enumerate(outs):
        emit("OUT", a=reg_of[id(o)], b=k)
    if len(insns) > 96:
        raise ffi.HordeError(f"index program too long ({len(insns)} instructions)")
    return Program(insns, tensors, n_vars, len(outs))

"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Vectorised numpy evaluator of symbolic index expressions: the oracle's
counterpart of interpreting the Int/Bool sub-language of AstTensor
(Ast.hs:487-528,575-579,654-665; AstInterpret.hs:237-283) at `Concrete`.
Written independently of the device VM (horde-ad_b200/csrc/vm.cuh): it walks
the expression DAG, not the bytecode.

Scalar semantics restated here:
  quotH a 0 = a, remH a 0 = a; quot/rem truncate toward zero  (Types.hs:394-419)
  sindex0 out of bounds -> 0                                   (OpsConcrete.hs:1498-1506)
  kargMax over the last axis, first maximum                    (OpsConcrete.hs:1060-1070)
"""
from __future__ import annotations

import numpy as np

_UN = {"exp": np.exp, "log": np.log, "sqrt": np.sqrt, "sin": np.sin, "cos": np.cos, "tan": np.tan,
       "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan, "sinh": np.sinh, "cosh": np.cosh,
       "tanh": np.tanh, "asinh": np.arcsinh, "acosh": np.arccosh, "atanh": np.arctanh}


def _trunc_quot(a, b):
    q = (np.abs(a) // np.maximum(np.abs(b), 1)) * np.sign(a) * np.sign(b)
    return q.astype(np.int64)


def evaluate(outs, env, M, backend, inputs=None, f32=False):
    """Evaluate expression DAGs at M positions.  env: var index -> int64[M]."""
    memo = {}
    rty = np.float32 if f32 else np.float64

    def rnd(x):
        return x.astype(rty).astype(np.float64) if f32 else x

    def ev(e):
        k = id(e)
        if k in memo:
            return memo[k]
        op = e.op
        a = [ev(x) for x in e.args]
        with np.errstate(all="ignore"):
            if op == "var":
                r = env[e.val]
            elif op == "const":
                if e.ty == "f":
                    r = rnd(np.full(M, e.val, dtype=np.float64))
                elif e.ty == "b":
                    r = np.full(M, bool(e.val))
                else:
                    r = np.full(M, e.val, dtype=np.int64)
            elif op == "in":
                x = inputs[e.val]
                r = x.astype(np.float64) if e.ty == "f" else x.astype(np.int64)
            elif op == "add":
                r = a[0] + a[1]
            elif op == "sub":
                r = a[0] - a[1]
            elif op == "mul":
                r = a[0] * a[1]
            elif op == "div":
                r = a[0] / a[1]
            elif op == "pow":
                r = np.power(a[0], a[1])
            elif op == "logbase":
                r = np.log(a[1]) / np.log(a[0])
            elif op == "atan2":
                r = np.arctan2(a[0], a[1])
            elif op == "neg":
                r = -a[0]
            elif op == "abs":
                r = np.abs(a[0])
            elif op == "signum":
                r = np.where(a[0] > 0, 1, np.where(a[0] < 0, -1, a[0])).astype(a[0].dtype)
            elif op == "quot":
                r = np.where(a[1] == 0, a[0], _trunc_quot(a[0], a[1]))
            elif op == "rem":
                r = np.where(a[1] == 0, a[0], a[0] - _trunc_quot(a[0], a[1]) * a[1])
            elif op == "min":
                r = np.where(a[1] < a[0], a[1], a[0])
            elif op == "max":
                r = np.where(a[0] < a[1], a[1], a[0])
            elif op == "leq":
                r = a[0] <= a[1]
            elif op == "lt":
                r = a[0] < a[1]
            elif op == "eq":
                r = a[0] == a[1]
            elif op == "and":
                r = a[0] & a[1]
            elif op == "or":
                r = a[0] | a[1]
            elif op == "not":
                r = ~a[0]
            elif op == "cond":
                r = np.where(a[0], a[1], a[2])
            elif op == "i2f":
                r = a[0].astype(np.float64)
            elif op == "floor":
                r = np.floor(a[0]).astype(np.int64)
            elif op == "fun1":
                if f32:
                    r = _UN[e.val](a[0].astype(np.float32)).astype(np.float64) if e.val in _UN else None
                else:
                    r = _UN[e.val](a[0]) if e.val in _UN else None
                if r is None:
                    if e.val == "recip":
                        r = 1.0 / a[0]
                    elif e.val == "negate":
                        r = -a[0]
                    else:
                        raise ValueError(e.val)
            elif op in ("tbl", "argmax", "argmin"):
                t = np.asarray(backend.to_numpy(e.tensor)) if not isinstance(e.tensor, np.ndarray) else e.tensor
                n = len(a)
                ok = np.ones(M, dtype=bool)
                safe = []
                for c, ext in zip(a, t.shape[:n]):
                    c = c.astype(np.int64)
                    ok &= (c >= 0) & (c < ext)
                    safe.append(np.where((c >= 0) & (c < ext), c, 0))
                if op == "tbl":
                    vals = t[tuple(safe)] if n else np.broadcast_to(t, (M,))
                    vals = vals.astype(np.float64) if e.ty == "f" else vals.astype(np.int64)
                    r = np.where(ok, vals, 0)
                else:
                    rows = t[tuple(safe)] if n else np.broadcast_to(t, (M,) + t.shape)
                    am = (np.argmax if op == "argmax" else np.argmin)(rows, axis=-1) if rows.shape[-1] else np.zeros(M, np.int64)
                    r = np.where(ok, am, 0).astype(np.int64)
            else:
                raise ValueError(f"unknown index op {op}")
        if e.ty == "f" and op not in ("var",):
            r = rnd(np.asarray(r, dtype=np.float64))
        r = np.broadcast_to(r, (M,)) if np.ndim(r) == 0 else r
        memo[k] = r
        return r

    return [ev(o) for o in outs]

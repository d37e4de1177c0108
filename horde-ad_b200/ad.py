"""Host-side mirror of horde-ad's dual-number AD over an arbitrary carrier:
`ADVal target` (CarriersADVal.hs:47, OpsADVal.hs:164-548) and the delta
evaluator `gradientFromDelta` / `evalRev` (DeltaEval.hs:64-77,274-730).

In the real drop-in this layer is horde-ad itself, unchanged Haskell running
over the new `Device` carrier; it is restated here only so that the reference's
gradient golden vectors and the `cgrad` call sequence can be
replayed through the C ABI without GHC.  It is target-generic: the same code
drives the CPU oracle in tests and the device backend in the product path, and
it touches tensors only through backend methods -- just like DeltaEval uses
only BaseTensor methods of `target`.

Every differentiable op records a node (a fresh id from a global counter, as
DeltaFreshId.hs:56-68) with a pullback implementing the transposition rule of
the corresponding Delta constructor; `grad` walks the nodes in DESCENDING id
order accumulating cotangents with taddTarget (evalRevFromnMap,
DeltaEval.hs:720-730; accumulation :139-148,317-332).
"""
from __future__ import annotations

import heapq
import os
import itertools
from typing import Callable, Sequence

import numpy as np

_ids = itertools.count(1)


class Node:
    __slots__ = ("id", "parents", "pullback", "acc")

    def __init__(self, parents, pullback, acc=None):
        self.id = next(_ids)
        self.parents = parents      # list[Node | None]
        self.pullback = pullback    # c -> list of cotangents (None = zero)
        # (slot, fn): fn(c, acc) is the pullback with the cotangent of parents[slot] ACCUMULATED into `acc` by the
        # pass that produces it (the sweep offers it only an accumulator it created itself)
        self.acc = acc


class Fresh:
    """A cotangent contribution in an array nobody else holds: the sweep may later update it in place."""
    __slots__ = ("x",)

    def __init__(self, x):
        self.x = x


class ADVal:
    """A dual number: primal tensor + delta node (None = constant)."""
    __slots__ = ("p", "d")

    def __init__(self, p, d=None):
        self.p = p
        self.d = d

    @property
    def shape(self):
        return tuple(self.p.shape)


class LazyAppend(ADVal):
    """`sappend a b` not yet evaluated.  A `sslice` that reads back exactly one operand returns that operand -- the
    rule `astSlice` of an `AstAppend` that the reference's simplifier applies to the symbolic artifact
    (AstSimplify.hs), here applied by the dual-number layer: the recurrent state `sappend vec1 vec2` of
    MnistRnnShaped2.hs:73-99 is only ever read back as its two halves, so neither the append nor (in the reverse
    sweep) the re-assembly of its cotangent is ever launched.  Any other use forces the append."""
    __slots__ = ("T", "a", "b", "_forced")

    def __init__(self, T, a, b):
        self.T, self.a, self.b, self._forced = T, a, b, None

    def _force(self):
        if self._forced is None:
            self._forced = self.T._sappend_now(self.a, self.b)
        return self._forced

    @property
    def p(self):
        return self._force().p

    @property
    def d(self):
        return self._force().d

    @property
    def shape(self):
        sa, sb = self.a.shape, self.b.shape
        return (sa[0] + sb[0],) + tuple(sa[1:])


_NO_LAZY_APPEND = bool(os.environ.get("HB_AD_NO_LAZY_APPEND"))   # measurement switch


def _inv_perm(perm):
    inv = [0] * len(perm)
    for i, p in enumerate(perm):
        inv[p] = i
    return inv


class ADBackend:
    """`target = ADVal base`: same method surface as a carrier backend."""

    def __init__(self, base):
        self.B = base
        self.name = "adval(" + base.name + ")"

    # ----------------------------------------------------------- plumbing
    def lift(self, x) -> ADVal:
        return x if isinstance(x, ADVal) else ADVal(x)

    def primal(self, x):
        return x.p if isinstance(x, ADVal) else x

    def _node(self, p, parents: Sequence[ADVal], pullback, acc=None) -> ADVal:
        ds = [a.d for a in parents]
        if all(d is None for d in ds):
            return ADVal(p)
        return ADVal(p, Node(ds, pullback, acc))

    def shape(self, t): return tuple(self.primal(t).shape)
    def dtype(self, t): return self.B.dtype(self.primal(t))
    def sconcrete(self, a): return ADVal(self.B.sconcrete(a))
    def srepl(self, shape, value, dtype=np.float64): return ADVal(self.B.srepl(shape, value, dtype))
    def sscalar(self, value, dtype=np.float64): return ADVal(self.B.sscalar(value, dtype))
    def to_numpy(self, t): return self.B.to_numpy(self.primal(t))
    def kfromS(self, t): return self.B.kfromS(self.primal(t))
    def siota(self, n, dtype=np.int64): return ADVal(self.B.siota(n, dtype))
    def sfromPrimal(self, t): return ADVal(self.primal(t))
    def sprimalPart(self, t): return self.primal(t)

    # index-expression references read primal (plain) parts
    def ix_tbl(self, t, ixs): return self.B.ix_tbl(self.primal(t), ixs)
    def ix_argmax(self, t, ixs): return self.B.ix_argmax(self.primal(t), ixs)
    def ix_argmin(self, t, ixs): return self.B.ix_argmin(self.primal(t), ixs)

    # --------------------------------------------------------- arithmetic
    def binary(self, op, a, b):
        B = self.B
        a, b = self.lift(a), self.lift(b)
        p = B.binary(op, a.p, b.p)
        if op == "add":
            return self._node(p, [a, b], lambda c: [c, c])
        if op == "sub":
            return self._node(p, [a, b], lambda c: [c, B.unary("negate", c)])
        if op == "mul":  # DeltaScale (DeltaEval.hs:392-396)
            return self._node(p, [a, b], lambda c: [B.binary("mul", b.p, c), B.binary("mul", a.p, c)])
        if op == "divide":
            def pb(c):
                ca = B.binary("divide", c, b.p)
                cb = B.unary("negate", B.binary("divide", B.binary("mul", c, p), b.p))
                return [ca, cb]
            return self._node(p, [a, b], pb)
        if op == "power":
            def pb(c):
                one = B.srepl(B.shape(p), 1, B.dtype(p))
                ca = B.binary("mul", c, B.binary("mul", b.p, B.binary("power", a.p, B.binary("sub", b.p, one))))
                cb = B.binary("mul", c, B.binary("mul", p, B.unary("log", a.p)))
                return [ca, cb]
            return self._node(p, [a, b], pb)
        if op == "atan2":  # atan2H y x: d/dy = x/(x^2+y^2), d/dx = -y/(x^2+y^2)
            def pb(c):
                r2 = B.binary("add", B.binary("mul", a.p, a.p), B.binary("mul", b.p, b.p))
                ca = B.binary("mul", c, B.binary("divide", b.p, r2))
                cb = B.binary("mul", c, B.unary("negate", B.binary("divide", a.p, r2)))
                return [ca, cb]
            return self._node(p, [a, b], pb)
        if op in ("max", "min", "quot", "rem", "logbase"):
            if a.d is None and b.d is None:
                return ADVal(p)
            raise NotImplementedError(f"derivative of {op}")
        raise ValueError(op)

    def unary(self, op, a):
        B = self.B
        a = self.lift(a)
        p = B.unary(op, a.p)
        if a.d is None:
            return ADVal(p)
        sh, dt = B.shape(p), B.dtype(p)
        one = lambda: B.srepl(sh, 1, dt)

        def scale(k):  # DeltaScale k d
            return self._node(p, [a], lambda c: [B.binary("mul", k(), c)])
        if op == "negate":
            return self._node(p, [a], lambda c: [B.unary("negate", c)])
        if op == "abs":
            return scale(lambda: B.unary("signum", a.p))
        if op == "signum":
            return ADVal(p)
        if op == "recip":
            return scale(lambda: B.unary("negate", B.binary("mul", p, p)))
        if op == "exp":
            return scale(lambda: p)
        if op == "log":
            return scale(lambda: B.unary("recip", a.p))
        if op == "sqrt":
            return scale(lambda: B.unary("recip", B.binary("add", p, p)))
        if op == "sin":
            return scale(lambda: B.unary("cos", a.p))
        if op == "cos":
            return scale(lambda: B.unary("negate", B.unary("sin", a.p)))
        if op == "tan":
            return scale(lambda: B.binary("add", one(), B.binary("mul", p, p)))
        if op == "tanh":
            return scale(lambda: B.binary("sub", one(), B.binary("mul", p, p)))
        if op == "sinh":
            return scale(lambda: B.unary("cosh", a.p))
        if op == "cosh":
            return scale(lambda: B.unary("sinh", a.p))
        if op == "asin":
            return scale(lambda: B.unary("recip", B.unary("sqrt", B.binary("sub", one(), B.binary("mul", a.p, a.p)))))
        if op == "acos":
            return scale(lambda: B.unary("negate", B.unary("recip", B.unary("sqrt", B.binary("sub", one(), B.binary("mul", a.p, a.p))))))
        if op == "atan":
            return scale(lambda: B.unary("recip", B.binary("add", one(), B.binary("mul", a.p, a.p))))
        if op == "asinh":
            return scale(lambda: B.unary("recip", B.unary("sqrt", B.binary("add", one(), B.binary("mul", a.p, a.p)))))
        if op == "acosh":
            return scale(lambda: B.unary("recip", B.unary("sqrt", B.binary("sub", B.binary("mul", a.p, a.p), one()))))
        if op == "atanh":
            return scale(lambda: B.unary("recip", B.binary("sub", one(), B.binary("mul", a.p, a.p))))
        raise ValueError(op)

    def scast(self, t, dtype):
        B = self.B
        t = self.lift(t)
        src = B.dtype(t.p)
        if not (np.issubdtype(np.dtype(dtype), np.floating) and np.issubdtype(src, np.floating)):
            return ADVal(B.scast(t.p, dtype))
        return self._node(B.scast(t.p, dtype), [t], lambda c: [B.scast(c, src)])

    def sfromIntegral(self, t, dtype): return ADVal(self.B.sfromIntegral(self.primal(t), dtype))
    def sfloor(self, t, dtype=np.int64): return ADVal(self.B.sfloor(self.primal(t), dtype))
    def sargMax(self, t): return ADVal(self.B.sargMax(self.primal(t)))
    def sargMin(self, t): return ADVal(self.B.sargMin(self.primal(t)))

    # ---------------------------------------------------------- structure
    def stranspose(self, perm, t):
        B = self.B
        t = self.lift(t)
        perm = list(perm)
        return self._node(B.stranspose(perm, t.p), [t], lambda c: [B.stranspose(_inv_perm(perm), c)])

    def sreshape(self, sh, t):
        B = self.B
        t = self.lift(t)
        old = B.shape(t.p)
        return self._node(B.sreshape(sh, t.p), [t], lambda c: [B.sreshape(old, c)])

    # DeltaReplicateS -> tssumN (DeltaEval.hs:534-555)
    def sreplicateN(self, shm, t):
        B = self.B
        t = self.lift(t)
        n = len(shm)
        return self._node(B.sreplicateN(shm, t.p), [t], lambda c: [B.ssumN(c, n)])

    def sreplicate(self, k, t): return self.sreplicateN([k], t)
    def sreplicate0N(self, shm, t0): return self.sreplicateN(shm, t0)

    # DeltaSumS -> tsreplicateN
    def ssumN(self, t, n_axes):
        B = self.B
        t = self.lift(t)
        shm = list(B.shape(t.p)[:n_axes])
        return self._node(B.ssumN(t.p, n_axes), [t], lambda c: [B.sreplicateN(shm, c)])

    def ssum(self, t): return self.ssumN(t, 1)
    def ssum0(self, t): return self.ssumN(t, len(self.shape(t)))

    def sslice(self, i, n, t):
        B = self.B
        t = self.lift(t)
        if isinstance(t, LazyAppend) and t._forced is None:
            m = t.a.shape[0]
            if i == 0 and n == m:
                return t.a
            if i == m and n == t.b.shape[0]:
                return t.b
        sh = B.shape(t.p)
        total, rest, dt = sh[0], list(sh[1:]), B.dtype(t.p)

        def pb(c):
            # the zero-padded cotangent (DeltaSlice, DeltaEval.hs:538-562) is handed over unevaluated: two slices
            # that tile the parent (state = sappend vec1 vec2 read back as two halves) become ONE sappend in the
            # sweep instead of two zero fills, four copies and an add
            if n == total:
                return [c]
            return [DeferredPad(c, i, n, total, rest, dt)]
        return self._node(B.sslice(i, n, t.p), [t], pb)

    def sappend(self, a, b):
        a, b = self.lift(a), self.lift(b)
        if _NO_LAZY_APPEND:
            return self._sappend_now(a, b)
        return LazyAppend(self, a, b)

    def _sappend_now(self, a, b):
        B = self.B
        m, n = B.shape(a.p)[0], B.shape(b.p)[0]
        return self._node(B.sappend(a.p, b.p), [a, b], lambda c: [B.sslice(0, m, c), B.sslice(m, n, c)])

    def sreverse(self, t):
        B = self.B
        t = self.lift(t)
        return self._node(B.sreverse(t.p), [t], lambda c: [B.sreverse(c)])

    def scontiguous(self, t):
        """Identity on values (a materialised copy of a view)."""
        B = self.B
        t = self.lift(t)
        return self._node(B.scontiguous(t.p), [t], lambda c: [c])

    def sfromVector(self, ts):
        B = self.B
        ts = [self.lift(t) for t in ts]
        return self._node(B.sfromVector([t.p for t in ts]), ts, lambda c: B.sunravelToList(c))

    def sunravelToList(self, t):
        return [self.sindex(t, [i]) for i in range(self.shape(t)[0])]

    # DeltaIndexS -> tsoneHot
    def sindex(self, t, ixs):
        B = self.B
        t = self.lift(t)
        ixs = [int(i) for i in ixs]
        shp = list(B.shape(t.p)[:len(ixs)])
        return self._node(B.sindex(t.p, ixs), [t], lambda c: [B.soneHot(shp, c, ixs)])

    # DeltaGatherS -> tsscatter ; DeltaScatterS -> tsgather (DeltaEval.hs:538-562)
    def sgather(self, shm, t, f, shp_rank, prog=None):
        B = self.B
        t = self.lift(t)
        shp = list(B.shape(t.p)[:shp_rank])
        shm = list(shm)
        return self._node(B.sgather(shm, t.p, f, shp_rank, prog), [t],
                          lambda c: [B.sscatter(shp, c, f, len(shm), prog)])

    def sscatter(self, shp, t, f, shm_rank, prog=None):
        B = self.B
        t = self.lift(t)
        shm = list(B.shape(t.p)[:shm_rank])
        shp = list(shp)
        return self._node(B.sscatter(shp, t.p, f, shm_rank, prog), [t],
                          lambda c: [B.sgather(shm, c, f, len(shp), prog)])

    # ---------------------------------------------------------- contraction
    # DeltaDot0S v d -> v * replicate c (DeltaEval.hs:496-501)
    def sdot0(self, a, b):
        B = self.B
        a, b = self.lift(a), self.lift(b)
        sh = list(B.shape(a.p))
        return self._node(B.sdot0(a.p, b.p), [a, b],
                          lambda c: [B.binary("mul", b.p, B.sreplicate0N(sh, c)),
                                     B.binary("mul", a.p, B.sreplicate0N(sh, c))])

    def _bcast_last(self, c, n):
        B = self.B
        r = len(B.shape(c))
        rep = B.sreplicate(n, c)                    # [n, sh...]
        return B.stranspose(list(range(1, r + 1)) + [0], rep)  # [sh..., n]

    def sdot1In(self, a, b):
        B = self.B
        a, b = self.lift(a), self.lift(b)
        n = B.shape(a.p)[-1]
        return self._node(B.sdot1In(a.p, b.p), [a, b],
                          lambda c: [B.binary("mul", b.p, self._bcast_last(c, n)),
                                     B.binary("mul", a.p, self._bcast_last(c, n))])

    def smatmul2(self, a, b):
        B = self.B
        a, b = self.lift(a), self.lift(b)
        # the two cotangents are handed to the sweep as DEFERRED products, so that a contribution to a cotangent
        # that already exists (a weight used at every time step of an unrolled loop) is accumulated by the GEMM
        # itself (`smatmul2_acc`) instead of being materialised and added in a second pass
        return self._node(B.smatmul2(a.p, b.p), [a, b],
                          lambda c: [DeferredMatmul(c, B.stranspose([1, 0], b.p)),
                                     DeferredMatmul(B.stranspose([1, 0], a.p), c)])

    def smatmul2_sum(self, a1, b1, a2, b2):
        """smatmul2 a1 b1 + smatmul2 a2 b2 as ONE node (the pre-activation `wX x + wS s` of rnnMnistLayerS,
        MnistRnnShaped2.hs:55-69): one paired launch forward; backward the two left cotangents are deferred
        products marked as partners (the sweep accumulates both weights' cotangents in one launch) and the two
        right cotangents, which share c, come from one paired launch."""
        B = self.B
        a1, b1, a2, b2 = self.lift(a1), self.lift(b1), self.lift(a2), self.lift(b2)
        tr = lambda t: B.stranspose([1, 0], t)

        def pb(c):
            da1 = DeferredMatmul(c, tr(b1.p)) if a1.d is not None else None
            da2 = DeferredMatmul(c, tr(b2.p)) if a2.d is not None else None
            if da1 is not None and da2 is not None:
                da1.partner, da2.partner = 2, 0          # slot of the partner in this node's parent list
            if b1.d is not None and b2.d is not None and B.shape(a1.p) == B.shape(a2.p) and not _use_group(B):
                db1, db2 = B.smatmul2_pair(tr(a1.p), c, tr(a2.p), c)
            else:       # deferred: the sweep launches them together with whatever else is waiting (smatmul2_group)
                db1 = DeferredMatmul(tr(a1.p), c) if b1.d is not None else None
                db2 = DeferredMatmul(tr(a2.p), c) if b2.d is not None else None
            return [da1, db1, da2, db2]
        return self._node(B.smatmul2_sum(a1.p, b1.p, a2.p, b2.p), [a1, b1, a2, b2], pb)

    def smatmul2_sum_tanh(self, a1, b1, a2, b2, bias):
        """tanh (a1 b1 + a2 b2 + bias): `smatmul2_sum` and `tanh_affine` as ONE node (one launch forward: bias and
        tanh ride in the epilogue of the paired product); the pullback is the composition of theirs."""
        B = self.B
        a1, b1, a2, b2, bias = self.lift(a1), self.lift(b1), self.lift(a2), self.lift(b2), self.lift(bias)
        return self._sum_tanh_node(B.smatmul2_sum_tanh(a1.p, b1.p, a2.p, b2.p, bias.p), a1, b1, a2, b2, bias)

    def smatmul2_sum_tanh_2(self, first, second):
        """(smatmul2_sum_tanh first, smatmul2_sum_tanh second) for two layer bodies that do not depend on each other
        (layer 1 at time t and layer 2 at time t - 1 of rnnMnistTwoS: the wavefront schedule of the unrolled loop):
        ONE launch forward (smatmul2_group: four products, two results, two tanh epilogues); two ordinary nodes, each
        with the pullback of smatmul2_sum_tanh."""
        B = self.B
        f = [self.lift(t) for t in first]
        g = [self.lift(t) for t in second]
        y1, y2 = B.smatmul2_group([(None, f[4].p), (None, g[4].p)],
                                  [(f[0].p, f[1].p, 0), (f[2].p, f[3].p, 0), (g[0].p, g[1].p, 1), (g[2].p, g[3].p, 1)])
        return self._sum_tanh_node(y1, *f), self._sum_tanh_node(y2, *g)

    def _sum_tanh_node(self, y, a1, b1, a2, b2, bias):
        B = self.B
        tr = lambda t: B.stranspose([1, 0], t)

        def pb(c0, acc=None):
            if acc is not None:
                c, dbias = B.tanh_affine_bwd_acc(y, c0, acc)
            else:
                c, dbias = B.tanh_affine_bwd(y, c0, want_bias=bias.d is not None)
                dbias = Fresh(dbias) if dbias is not None else None
            da1 = DeferredMatmul(c, tr(b1.p)) if a1.d is not None else None
            da2 = DeferredMatmul(c, tr(b2.p)) if a2.d is not None else None
            if da1 is not None and da2 is not None:
                da1.partner, da2.partner = 2, 0
            if b1.d is not None and b2.d is not None and B.shape(a1.p) == B.shape(a2.p) and not _use_group(B):
                db1, db2 = B.smatmul2_pair(tr(a1.p), c, tr(a2.p), c)
            else:       # deferred: the sweep launches them together with whatever else is waiting (smatmul2_group)
                db1 = DeferredMatmul(tr(a1.p), c) if b1.d is not None else None
                db2 = DeferredMatmul(tr(a2.p), c) if b2.d is not None else None
            return [da1, db1, da2, db2, dbias]
        use_acc = hasattr(B, "tanh_affine_bwd_acc") and not _NO_BIAS_ACC
        return self._node(y, [a1, b1, a2, b2, bias], pb, acc=(4, pb) if use_acc else None)

    def smatvecmul(self, m, v):
        B = self.B
        m, v = self.lift(m), self.lift(v)
        rows = B.shape(m.p)[0]
        cols = B.shape(m.p)[1]
        return self._node(B.smatvecmul(m.p, v.p), [m, v],
                          lambda c: [B.binary("mul", B.stranspose([1, 0], B.sreplicate(cols, c)), B.sreplicate(rows, v.p)),
                                     B.smatvecmul(B.stranspose([1, 0], m.p), c)])

    # --------------------------------------- fused blocks (custom derivatives)
    def conv2dPreserving(self, K, A):
        B = self.B
        K, A = self.lift(K), self.lift(A)
        KH, KW = B.shape(K.p)[2], B.shape(K.p)[3]
        return self._node(B.conv2dPreserving(K.p, A.p), [K, A],
                          lambda c: [B.conv2d_dkrn(A.p, c, KH, KW) if K.d is not None else None,
                                     B.conv2d_dinp(K.p, c) if A.d is not None else None])

    def conv_relu_pool(self, K, bias, A, ksize):
        """convMnistLayerS as one node (MnistCnnShaped2.hs:42-62): conv, then the
        fused bias + relu + max-pool pass; the adjoint needs only the pooled
        cotangent and one code byte per pooled element."""
        B = self.B
        K, bias, A = self.lift(K), self.lift(bias), self.lift(A)
        y, code = B.conv_relu_pool_fwd(K.p, bias.p, A.p, ksize)

        def pb(c):
            dK, dbias, dA = B.conv_relu_pool_bwd(K.p, A.p, c, code, ksize, want_bias=bias.d is not None,
                                                 want_dA=A.d is not None)
            return [dK, dbias, dA]
        return self._node(y, [K, bias, A], pb)

    def tanh_affine(self, u, v, bias):
        """The body of rnnMnistLayerS (MnistRnnShaped2.hs:55-69) as one node: tanh ((u + v) + bias
        stretched over the batch); dz = (1 - y * y) * c is the cotangent of u and of v."""
        B = self.B
        u, bias = self.lift(u), self.lift(bias)
        if v is None:        # u already holds the sum (smatmul2_sum)
            y = B.tanh_affine_fwd(u.p, None, bias.p)

            def pb1(c):
                dz, db = B.tanh_affine_bwd(y, c, want_bias=bias.d is not None)
                return [dz, db]
            return self._node(y, [u, bias], pb1)
        v = self.lift(v)
        y = B.tanh_affine_fwd(u.p, v.p, bias.p)

        def pb(c):
            dz, db = B.tanh_affine_bwd(y, c, want_bias=bias.d is not None)
            return [dz, dz, db]
        return self._node(y, [u, v, bias], pb)

    def custom_unary(self, t, primal, vjp):
        """tD with a hand-written dual part (e.g. logistic,
        CommonRankedOps.hs:55-63): primal(B, x) and vjp(B, x, y, c)."""
        B = self.B
        t = self.lift(t)
        y = primal(B, t.p)
        return self._node(y, [t], lambda c: [vjp(B, t.p, y, c)])

    def relu(self, t):
        B = self.B
        t = self.lift(t)
        sh, dt = B.shape(t.p), B.dtype(t.p)

        def pb(c):  # oneIfGtZero * c, with oneIfGtZero = relu(x) > 0 mask
            mask = B.unary("signum", B.relu(t.p))
            return [B.binary("mul", mask, c)]
        return self._node(B.relu(t.p), [t], pb)

    def maxpool2d(self, x, ksize, stride):
        B = self.B
        x = self.lift(x)
        H, W = B.shape(x.p)[2], B.shape(x.p)[3]
        y, am = B.maxpool2d(x.p, ksize, stride)
        return self._node(y, [x], lambda c: [B.maxpool2d_bwd(c, am, ksize, stride, H, W)])

    # lossSoftMaxCrossEntropyS: tD primal (dual via (softMax - expected) * du)
    def softmax_xent(self, logits, expected, scale=1.0, want_grad=True):
        B = self.B
        logits = self.lift(logits)
        loss, dl = B.softmax_xent(logits.p, self.primal(expected), scale, want_grad=logits.d is not None)
        sh = list(B.shape(logits.p))
        return self._node(loss, [logits], lambda c: [B.binary("mul", dl, B.sreplicate0N(sh, c))]), None

    def prod_fold(self, x):
        """sfold (*) 1 x with the DeltaMapAccumL transposition (DeltaEval.hs:351-389)."""
        B = self.B
        x = self.lift(x)
        prod, _ = B.prod_fold_grad(x.p, 1.0, want_grad=False)
        return self._node(prod, [x], lambda c: [B.prod_fold_grad(x.p, float(B.kfromS(c)))[1]])


_NO_MATMUL_ACC = bool(os.environ.get("HB_AD_NO_MATMUL_ACC"))   # measurement switch: product + separate add
_NO_MATMUL_LIST = bool(os.environ.get("HB_AD_NO_MATMUL_LIST"))  # measurement switch: accumulate weight cotangents per step
_NO_BIAS_ACC = bool(os.environ.get("HB_AD_NO_BIAS_ACC"))        # measurement switch: bias cotangents by a separate add
_NO_MATMUL_GROUP = bool(os.environ.get("HB_AD_NO_MATMUL_GROUP"))  # measurement switch: deferred products one by one


def _use_group(base):
    return hasattr(base, "smatmul2_group") and not _NO_MATMUL_GROUP and not _NO_MATMUL_ACC


class DeferredMatmul:
    """A cotangent contribution `smatmul2 a b` not yet evaluated (see ADBackend.smatmul2)."""
    __slots__ = ("a", "b", "partner")

    def __init__(self, a, b):
        self.a, self.b = a, b
        self.partner = None     # parent slot of a second deferred product that may share the launch

    def force(self, base):
        return base.smatmul2(self.a, self.b)


class DeferredPad:
    """A cotangent contribution `zeros ++ c ++ zeros` along the outermost axis, not yet evaluated."""
    __slots__ = ("c", "i", "n", "total", "rest", "dt")

    def __init__(self, c, i, n, total, rest, dt):
        self.c, self.i, self.n, self.total, self.rest, self.dt = c, i, n, total, rest, dt

    def force(self, base):
        parts = self.c
        if self.i > 0:
            parts = base.sappend(base.srepl([self.i] + self.rest, 0, self.dt), parts)
        if self.total - self.i - self.n > 0:
            parts = base.sappend(parts, base.srepl([self.total - self.i - self.n] + self.rest, 0, self.dt))
        return parts

    def tile_with(self, other, base):
        """self + other when the two windows are adjacent: one sappend (still a pad if they do not cover the
        whole axis); None when they overlap or leave a gap."""
        lo, hi = (self, other) if self.i <= other.i else (other, self)
        if lo.i + lo.n != hi.i:
            return None
        joined = base.sappend(lo.c, hi.c)
        if lo.i == 0 and hi.i + hi.n == self.total:
            return joined
        return DeferredPad(joined, lo.i, lo.n + hi.n, self.total, self.rest, self.dt)


def grad(base, f: Callable, inputs: Sequence, dt=None):
    """cgrad / cvjp (ADEngine.hs:544-555): returns (value, [gradients]).

    `f` receives the ADBackend and the inputs as ADVals.  dt=None means a
    cotangent of ones (crevOnParams Nothing, OpsADVal.hs:62-73).  Gradients of
    untouched inputs are explicit zeros (rebuildInputs, DeltaEval.hs:151-184).
    """
    T = ADBackend(base)
    in_nodes = [Node([], None) for _ in inputs]
    advals = [ADVal(x, n) for x, n in zip(inputs, in_nodes)]
    out = f(T, *advals)
    out = T.lift(out)
    if dt is None:
        dt = base.srepl(base.shape(out.p), 1, base.dtype(out.p))
    grads = [None] * len(inputs)
    if out.d is not None:
        input_ix = {n.id: i for i, n in enumerate(in_nodes)}
        cot = {out.d.id: dt}
        nodes = {out.d.id: out.d}
        # cotangents that this sweep itself created by an accumulation: only those may be updated in
        # place (a pullback may hand the SAME array to several parents, e.g. `add` returns [c, c])
        owned = set()
        # deferred products whose parent is an INPUT (a weight used at every step of an unrolled loop): kept as a list
        # and evaluated ONCE, when the input is reached at the end of the sweep, as one product over the concatenated
        # reduced axes (hb_matmul2_list) instead of one small accumulating product per step
        pending = {}
        use_list = hasattr(base, "smatmul2_list") and not _NO_MATMUL_LIST and not _NO_MATMUL_ACC
        # deferred products whose parent is NOT an input stay unevaluated until one of them is needed (its parent is
        # reached); then ALL that are waiting go out in groups of up to four products per launch (smatmul2_group: e.g.
        # the three input cotangents of one time step of a two-layer recurrence), each accumulated into the cotangent
        # it belongs to by the product's own epilogue
        lazy = {}
        use_group = _use_group(base)

        def flush():
            targets = list(lazy.items())
            lazy.clear()
            group, n_prod = [], 0

            def launch():
                results, products, post = [], [], []
                for r, (pid, lst) in enumerate(group):
                    old = cot.pop(pid, None)
                    if isinstance(old, DeferredPad):
                        old = old.force(base)
                        owned.add(pid)
                    if old is not None and pid in owned:
                        results.append((old, None))
                        post.append(None)
                    else:                      # nothing yet, or an array other parents may hold too
                        results.append((None, None))
                        post.append(old)
                    del old
                    products += [(d.a, d.b, r) for d in lst]
                if len(products) == 1:         # nothing to share a launch with: the plain (accumulating) product
                    (acc, _), (a, b, _) = results[0], products[0]
                    del results
                    outs = [base.smatmul2(a, b) if acc is None else base.smatmul2_acc(acc, a, b)]
                    del acc
                else:
                    outs = base.smatmul2_group(results, products)
                    del results
                for (pid, _), o, old in zip(group, outs, post):
                    cot[pid] = o if old is None else base.binary("add", old, o)
                    owned.add(pid)
            for pid, lst in targets:
                while lst:
                    take = lst[:4 - n_prod] if len(group) < 4 else []
                    if not take:
                        launch()
                        group, n_prod = [], 0
                        continue
                    group.append((pid, take))
                    n_prod += len(take)
                    lst = lst[len(take):]
                    if n_prod == 4 and lst:        # the rest of this target in the next launch, on top of this one
                        launch()
                        group, n_prod = [], 0
            if group:
                launch()
        heap = [-out.d.id]
        while heap:
            nid = -heapq.heappop(heap)
            node = nodes.pop(nid)
            if nid in lazy:
                flush()
            c = cot.pop(nid)
            if isinstance(c, DeferredPad):
                c = c.force(base)
            if nid in input_ix:
                if nid in pending:
                    lst = pending.pop(nid)
                    if len(lst) == 1:                  # a single contribution: the plain product
                        lst_val = base.smatmul2(*lst[0])
                        c = lst_val if c is None else base.binary("add", c, lst_val)
                    elif c is None or nid in owned:    # an accumulator this sweep created: may be updated in place
                        c = base.smatmul2_list(lst, c)
                    else:                              # an array other parents may hold too
                        c = base.binary("add", c, base.smatmul2_list(lst, None))
                grads[input_ix[nid]] = c
                continue
            acc_slot = None
            if node.acc is not None:
                slot, fn = node.acc
                parent = node.parents[slot]
                if (parent is not None and parent.id in owned and cot.get(parent.id) is not None
                        and not isinstance(cot[parent.id], DeferredPad)):
                    acc_slot = slot
            if acc_slot is not None:
                pid = node.parents[acc_slot].id
                held = cot.pop(pid)            # hand over the only reference: the update may then be in place
                cs = list(fn(c, held))
                del held
                cot[pid] = cs[acc_slot]
                cs[acc_slot] = None
            else:
                cs = list(node.pullback(c))
            if use_list:
                for i, pc in enumerate(cs):
                    parent = node.parents[i]
                    if isinstance(pc, DeferredMatmul) and parent is not None and parent.id in input_ix:
                        pending.setdefault(parent.id, []).append((pc.a, pc.b))
                        if parent.id not in cot:
                            cot[parent.id] = None
                            nodes[parent.id] = parent
                            heapq.heappush(heap, -parent.id)
                        cs[i] = None
            # two deferred products marked as partners whose cotangents both exist and are owned by the sweep:
            # both accumulations in one launch
            for i, pc in enumerate(cs):
                j = pc.partner if isinstance(pc, DeferredMatmul) else None
                if j is None or j <= i or not isinstance(cs[j], DeferredMatmul) or _NO_MATMUL_ACC:
                    continue
                pi, pj = node.parents[i], node.parents[j]
                if (pi is not None and pj is not None and pi.id != pj.id and pi.id in owned and pj.id in owned
                        and pi.id in cot and pj.id in cot):
                    cot[pi.id], cot[pj.id] = base.smatmul2_acc_pair(cot[pi.id], pc.a, pc.b, cot[pj.id], cs[j].a, cs[j].b)
                    cs[i] = cs[j] = None
            for parent, pc in zip(node.parents, cs):
                if parent is None or pc is None:
                    continue
                if use_group and isinstance(pc, DeferredMatmul):
                    lazy.setdefault(parent.id, []).append(pc)
                    if parent.id not in cot:
                        cot[parent.id] = None          # placeholder: reached through the heap like any other
                        nodes[parent.id] = parent
                        heapq.heappush(heap, -parent.id)
                    continue
                fresh = isinstance(pc, Fresh)
                if fresh:
                    pc = pc.x
                    if cot.get(parent.id) is None:     # first contribution: the sweep owns this array
                        owned.add(parent.id)
                deferred = isinstance(pc, DeferredMatmul)
                if parent.id in cot and cot[parent.id] is None:    # placeholder of an input with pending products
                    if deferred or isinstance(pc, DeferredPad):
                        pc = pc.force(base)
                        owned.add(parent.id)
                    cot[parent.id] = pc
                    continue
                if isinstance(pc, DeferredPad):
                    old = cot.get(parent.id)
                    if old is None:                       # stays unevaluated until a second window arrives
                        cot[parent.id] = pc
                        nodes[parent.id] = parent
                        heapq.heappush(heap, -parent.id)
                        continue
                    if isinstance(old, DeferredPad):
                        joined = old.tile_with(pc, base)
                        if joined is not None:
                            cot[parent.id] = joined
                            if not isinstance(joined, DeferredPad):
                                owned.add(parent.id)      # a fresh array nobody else holds
                            continue
                        cot[parent.id] = old.force(base)
                        owned.add(parent.id)
                    pc = pc.force(base)
                elif isinstance(cot.get(parent.id), DeferredPad):
                    cot[parent.id] = cot[parent.id].force(base)
                    owned.add(parent.id)
                if parent.id in cot:
                    if parent.id in owned:
                        if deferred and not _NO_MATMUL_ACC:
                            cot[parent.id] = base.smatmul2_acc(cot[parent.id], pc.a, pc.b)
                        elif deferred:
                            cot[parent.id] = base.add_accumulate(cot[parent.id], pc.force(base))
                        else:
                            cot[parent.id] = base.add_accumulate(cot[parent.id], pc)
                    else:
                        cot[parent.id] = base.binary("add", cot[parent.id], pc.force(base) if deferred else pc)
                        owned.add(parent.id)
                else:
                    if deferred:   # a fresh array that nobody else holds: the sweep may update it in place
                        pc = pc.force(base)
                        owned.add(parent.id)
                    cot[parent.id] = pc
                    nodes[parent.id] = parent
                    heapq.heappush(heap, -parent.id)
    for i, x in enumerate(inputs):
        if grads[i] is None:
            grads[i] = base.srepl(base.shape(x), 0, base.dtype(x))
    return out.p, grads


# =====================================================================
# Forward mode: cjvp (ADEngine.hs:544-555; evalFwd, DeltaEval.hs:778)
# =====================================================================
class Dual:
    """Primal + tangent (None = zero tangent / constant)."""
    __slots__ = ("p", "t")

    def __init__(self, p, t=None):
        self.p = p
        self.t = t

    @property
    def shape(self):
        return tuple(self.p.shape)


class FwdBackend:
    """`target = ADVal base` evaluated forwards: every op computes its primal
    and pushes the tangent through the same linear map the reverse rule
    transposes.  Same method surface as ADBackend for the ops the example
    networks use."""

    def __init__(self, base):
        self.B = base
        self.name = "fwd(" + base.name + ")"

    def lift(self, x):
        return x if isinstance(x, Dual) else Dual(x)

    def primal(self, x):
        return x.p if isinstance(x, Dual) else x

    def _lin(self, ts, f):
        """f applied to the tangents (all operands of a linear op); None stays None."""
        return None if all(t is None for t in ts) else f

    def _add(self, a, b):
        if a is None:
            return b
        if b is None:
            return a
        return self.B.binary("add", a, b)

    def _zeros_like(self, p):
        return self.B.srepl(list(self.B.shape(p)), 0, self.B.dtype(p))

    def shape(self, t): return tuple(self.primal(t).shape)
    def dtype(self, t): return self.B.dtype(self.primal(t))
    def sconcrete(self, a): return Dual(self.B.sconcrete(a))
    def srepl(self, shape, value, dtype=np.float64): return Dual(self.B.srepl(shape, value, dtype))
    def sscalar(self, value, dtype=np.float64): return Dual(self.B.sscalar(value, dtype))
    def kfromS(self, t): return self.B.kfromS(self.primal(t))
    def sprimalPart(self, t): return self.primal(t)
    def ix_tbl(self, t, ixs): return self.B.ix_tbl(self.primal(t), ixs)
    def ix_argmax(self, t, ixs): return self.B.ix_argmax(self.primal(t), ixs)
    def ix_argmin(self, t, ixs): return self.B.ix_argmin(self.primal(t), ixs)

    def _map(self, x, f):
        """A linear structural op: f on the primal and on the tangent."""
        x = self.lift(x)
        return Dual(f(x.p), None if x.t is None else f(x.t))

    def binary(self, op, a, b):
        B = self.B
        a, b = self.lift(a), self.lift(b)
        p = B.binary(op, a.p, b.p)
        if a.t is None and b.t is None:
            return Dual(p)
        if op == "add":
            return Dual(p, self._add(a.t, b.t))
        if op == "sub":
            return Dual(p, self._add(a.t, None if b.t is None else B.unary("negate", b.t)))
        if op == "mul":
            ta = None if a.t is None else B.binary("mul", a.t, b.p)
            tb = None if b.t is None else B.binary("mul", a.p, b.t)
            return Dual(p, self._add(ta, tb))
        if op == "divide":
            ta = None if a.t is None else B.binary("divide", a.t, b.p)
            tb = None if b.t is None else B.unary("negate", B.binary("divide", B.binary("mul", p, b.t), b.p))
            return Dual(p, self._add(ta, tb))
        if op == "atan2":  # atan2H y x: d/dy = x / (x^2 + y^2), d/dx = -y / (x^2 + y^2)  (same factors as the pullback)
            r2 = B.binary("add", B.binary("mul", a.p, a.p), B.binary("mul", b.p, b.p))
            ta = None if a.t is None else B.binary("mul", a.t, B.binary("divide", b.p, r2))
            tb = None if b.t is None else B.binary("mul", b.t, B.unary("negate", B.binary("divide", a.p, r2)))
            return Dual(p, self._add(ta, tb))
        if op == "power":
            one = B.srepl(B.shape(p), 1, B.dtype(p))
            ta = None if a.t is None else B.binary("mul", a.t, B.binary("mul", b.p, B.binary("power", a.p, B.binary("sub", b.p, one))))
            tb = None if b.t is None else B.binary("mul", b.t, B.binary("mul", p, B.unary("log", a.p)))
            return Dual(p, self._add(ta, tb))
        raise NotImplementedError(f"forward derivative of {op}")

    def unary(self, op, a):
        B = self.B
        a = self.lift(a)
        p = B.unary(op, a.p)
        if a.t is None or op == "signum":
            return Dual(p)
        one = lambda: B.srepl(list(B.shape(p)), 1, B.dtype(p))
        sq = lambda x: B.binary("mul", x, x)
        k = {
            "negate": lambda: B.unary("negate", one()), "abs": lambda: B.unary("signum", a.p),
            "recip": lambda: B.unary("negate", sq(p)), "exp": lambda: p, "log": lambda: B.unary("recip", a.p),
            "sqrt": lambda: B.unary("recip", B.binary("add", p, p)), "sin": lambda: B.unary("cos", a.p),
            "cos": lambda: B.unary("negate", B.unary("sin", a.p)),
            "tan": lambda: B.binary("add", one(), sq(p)), "tanh": lambda: B.binary("sub", one(), sq(p)),
            "sinh": lambda: B.unary("cosh", a.p), "cosh": lambda: B.unary("sinh", a.p),
            "atan": lambda: B.unary("recip", B.binary("add", one(), sq(a.p))),
        }[op]()
        return Dual(p, B.binary("mul", k, a.t))

    def custom_unary(self, t, primal, vjp):
        B = self.B
        t = self.lift(t)
        y = primal(B, t.p)
        # an elementwise custom dual is symmetric: the vjp IS the jvp
        return Dual(y, None if t.t is None else vjp(B, t.p, y, t.t))

    def scast(self, t, dtype):
        B = self.B
        t = self.lift(t)
        fl = np.issubdtype(np.dtype(dtype), np.floating) and np.issubdtype(B.dtype(t.p), np.floating)
        return Dual(B.scast(t.p, dtype), B.scast(t.t, dtype) if (fl and t.t is not None) else None)

    # structure (all linear)
    # integer-valued results carry no tangent (tprimalPart of a plain tensor)
    def sfromIntegral(self, t, dtype): return Dual(self.B.sfromIntegral(self.primal(t), dtype))
    def sfloor(self, t, dtype=np.int64): return Dual(self.B.sfloor(self.primal(t), dtype))
    def sargMax(self, t): return Dual(self.B.sargMax(self.primal(t)))
    def sargMin(self, t): return Dual(self.B.sargMin(self.primal(t)))
    def siota(self, n, dtype=np.int64): return Dual(self.B.siota(n, dtype))
    def to_numpy(self, t): return self.B.to_numpy(self.primal(t))

    def stranspose(self, perm, t): return self._map(t, lambda x: self.B.stranspose(list(perm), x))
    def sreshape(self, sh, t): return self._map(t, lambda x: self.B.sreshape(list(sh), x))
    def sreplicateN(self, shm, t): return self._map(t, lambda x: self.B.sreplicateN(list(shm), x))
    def sreplicate(self, k, t): return self.sreplicateN([k], t)
    def sreplicate0N(self, shm, t0): return self.sreplicateN(shm, t0)
    def ssumN(self, t, n): return self._map(t, lambda x: self.B.ssumN(x, n))
    def ssum(self, t): return self.ssumN(t, 1)
    def ssum0(self, t): return self.ssumN(t, len(self.shape(t)))
    def sslice(self, i, n, t): return self._map(t, lambda x: self.B.sslice(i, n, x))
    def sreverse(self, t): return self._map(t, lambda x: self.B.sreverse(x))
    def scontiguous(self, t): return self._map(t, lambda x: self.B.scontiguous(x))
    def sindex(self, t, ixs): return self._map(t, lambda x: self.B.sindex(x, [int(i) for i in ixs]))
    def sunravelToList(self, t): return [self.sindex(t, [i]) for i in range(self.shape(t)[0])]

    def sgather(self, shm, t, f, shp_rank, prog=None):
        return self._map(t, lambda x: self.B.sgather(list(shm), x, f, shp_rank, prog))

    def sscatter(self, shp, t, f, shm_rank, prog=None):
        return self._map(t, lambda x: self.B.sscatter(list(shp), x, f, shm_rank, prog))

    def _tangent_or_zero(self, d):
        return d.t if d.t is not None else self._zeros_like(d.p)

    def sappend(self, a, b):
        a, b = self.lift(a), self.lift(b)
        t = None if (a.t is None and b.t is None) else self.B.sappend(self._tangent_or_zero(a), self._tangent_or_zero(b))
        return Dual(self.B.sappend(a.p, b.p), t)

    def sfromVector(self, ts):
        ts = [self.lift(t) for t in ts]
        tt = None if all(t.t is None for t in ts) else self.B.sfromVector([self._tangent_or_zero(t) for t in ts])
        return Dual(self.B.sfromVector([t.p for t in ts]), tt)

    # contractions (bilinear)
    def _bilinear(self, f, a, b):
        a, b = self.lift(a), self.lift(b)
        ta = None if a.t is None else f(a.t, b.p)
        tb = None if b.t is None else f(a.p, b.t)
        return Dual(f(a.p, b.p), self._add(ta, tb))

    def sdot0(self, a, b): return self._bilinear(self.B.sdot0, a, b)
    def sdot1In(self, a, b): return self._bilinear(self.B.sdot1In, a, b)
    def smatmul2(self, a, b): return self._bilinear(self.B.smatmul2, a, b)
    def smatvecmul(self, m, v): return self._bilinear(self.B.smatvecmul, m, v)

    def smatmul2_sum(self, a1, b1, a2, b2):
        return self.binary("add", self.smatmul2(a1, b1), self.smatmul2(a2, b2))

    def smatmul2_sum_tanh(self, a1, b1, a2, b2, bias):
        return self.tanh_affine(self.smatmul2_sum(a1, b1, a2, b2), None, bias)
    def conv2dPreserving(self, K, A): return self._bilinear(self.B.conv2dPreserving, K, A)

    def sdot1InSumN(self, a, b, n):
        return self._bilinear(lambda x, y: self.B.sdot1InSumN(x, y, n), a, b)

    # NN blocks
    def relu(self, t):
        B = self.B
        t = self.lift(t)
        y = B.relu(t.p)
        return Dual(y, None if t.t is None else B.binary("mul", B.unary("signum", y), t.t))

    def maxpool2d(self, x, ksize, stride):
        """Tangent of the pooled value = the tangent at the window's argmax: a
        gather whose index function reads the argmax tensor (table lookup)."""
        B = self.B
        x = self.lift(x)
        y, am = B.maxpool2d(x.p, ksize, stride)
        if x.t is None:
            return Dual(y)
        N, Cc, Ho, Wo = B.shape(y)

        def f(ix):
            a = B.ix_tbl(am, ix)
            return [ix[0], ix[1], stride * ix[2] + a.quotH(ksize), stride * ix[3] + a.remH(ksize)]
        return Dual(y, B.sgather([N, Cc, Ho, Wo], x.t, f, 4))

    def tanh_affine(self, u, v, bias):
        batch = self.shape(u)[1]
        return self.unary("tanh", self.binary("add", self.binary("add", u, v) if v is not None else u,
                                              self.stranspose([1, 0], self.sreplicate(batch, bias))))

    def conv_relu_pool(self, K, bias, A, ksize):
        N, _, H, W = self.shape(A)
        pre = self.binary("add", self.conv2dPreserving(K, A),
                          self.stranspose([0, 3, 1, 2], self.sreplicateN([N, H, W], bias)))
        return self.maxpool2d(self.relu(pre), ksize, ksize)

    def softmax_xent(self, logits, expected, scale=1.0, want_grad=True):
        B = self.B
        logits = self.lift(logits)
        loss, dl = B.softmax_xent(logits.p, self.primal(expected), scale, want_grad=logits.t is not None)
        return Dual(loss, None if logits.t is None else B.sdot0(dl, logits.t)), None


def jvp(base, f: Callable, inputs: Sequence, tangents: Sequence):
    """cjvp (ADEngine.hs:544-555): (value, directional derivative of f at
    `inputs` along `tangents`); a None tangent is zero."""
    T = FwdBackend(base)
    out = T.lift(f(T, *[Dual(x, t) for x, t in zip(inputs, tangents)]))
    t = out.t if out.t is not None else T._zeros_like(out.p)
    return out.p, t

"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy) of the reference's `Concrete` backend for the hot
path: `instance BaseTensor Concrete` (src/HordeAd/Core/OpsConcrete.hs) and the
NN blocks of src/HordeAd/External/CommonShapedOps.hs.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product path (horde-ad_b200/) never does.

The arithmetic of the real reference lives in the un-vendored Hackage package
ox-arrays (>= 0.2, < 0.3; C kernels in its sublibrary strided-array-ops) on top
of orthotope 0.1.8.x (horde-ad.cabal:207-208,403): neither is under
/root/reference and there is no GHC here, so the reference cannot be executed.
This restatement is pinned instead by the reference's own literal golden
vectors (tests/test_oracle_golden.py).  What stays
unpinned at the third-party boundary: floating-point summation order inside
ox-arrays reductions (tolerance-level parity only) and argmin/argmax
tie-breaking (restated as "first extremum in row-major order").

Every function cites the reference lines it follows (relative to
/root/reference).  Tensors are numpy arrays; index functions are closures over
the symbolic index scalar (evaluated here by oracle.ix_eval, independently of
the device VM).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from . import ix_eval


def _is_float(a) -> bool:
    return np.issubdtype(a.dtype, np.floating)


def _pow_ghc(x: float, y: int) -> float:
    """`x ^ y` as GHC computes it (GHC.Real.powImpl / powImplAcc: binary exponentiation by a fixed sequence of
    multiplications) -- the `betaOne ^ tAdamNew` of updateWithGradientAdam (OptimizerTools.hs:154-155)."""
    x = float(x)
    if y <= 0:
        return 1.0
    while y % 2 == 0:
        x, y = x * x, y // 2
    if y == 1:
        return x
    z, x, y = x, x * x, y // 2
    while True:
        if y % 2 == 0:
            x, y = x * x, y // 2
        elif y == 1:
            return x * z
        else:
            z, x, y = x * z, x * x, y // 2


class ConcreteBackend:
    """`target = Concrete` (CarriersConcrete.hs:150-155,293)."""

    name = "concrete"

    # ------------------------------------------------------------ transfer
    # tsconcrete = identity (OpsConcrete.hs:120-124)
    def sconcrete(self, a):
        return np.array(a)

    def to_numpy(self, t):
        return np.array(t)

    def shape(self, t): return tuple(t.shape)
    def dtype(self, t): return t.dtype

    # srepl / treplTarget (Ops.hs:114-121, CarriersConcrete.hs:243-282)
    def srepl(self, shape, value, dtype=np.float64):
        return np.broadcast_to(np.array(value, dtype=dtype), tuple(shape))

    def sscalar(self, value, dtype=np.float64):
        return np.array(value, dtype=dtype)

    # kfromS (OpsConcrete.hs:777-812)
    def kfromS(self, t):
        assert t.size == 1
        return t.reshape(())[()]

    def sync(self):
        pass

    # tsfromVector (OpsConcrete.hs:144-149): "sfromVector: empty vector" on []
    def sfromVector(self, ts):
        if len(ts) == 0:
            raise ValueError("sfromVector: empty vector")
        return np.stack([np.asarray(t) for t in ts], axis=0)

    # tsunravelToList (OpsConcrete.hs:160-162)
    def sunravelToList(self, t): return [t[i] for i in range(t.shape[0])]

    # --------------------------------------------------------------- views
    # tstranspose (OpsConcrete.hs:648-649): PermutePrefix, out.shape[i] = sh[perm[i]]
    def stranspose(self, perm, t):
        perm = list(perm) + list(range(len(perm), t.ndim))
        return np.transpose(t, perm)

    # tsreplicateN / tsreplicate / tsreplicate0N (OpsConcrete.hs:243-246)
    def sreplicateN(self, shm, t): return np.broadcast_to(t, tuple(shm) + tuple(t.shape))
    def sreplicate(self, k, t): return self.sreplicateN([k], t)
    def sreplicate0N(self, shm, t0): return self.sreplicateN(shm, t0)

    # tsslice / tsreverse / tsappend / tsreshape (OpsConcrete.hs:638-652)
    def sslice(self, i, n, t):
        assert 0 <= i and i + n <= t.shape[0]
        return t[i:i + n]

    def sreverse(self, t): return t[::-1]
    def sappend(self, a, b): return np.concatenate([a, b], axis=0)
    def sreshape(self, sh, t): return np.reshape(t, tuple(sh))
    def scontiguous(self, t): return np.ascontiguousarray(t)

    # tindexZS (OpsConcrete.hs:1451-1490): out of bounds -> zeros of shape shn
    def sindex(self, t, ixs):
        ixs = [int(i) for i in ixs]
        for k, i in enumerate(ixs):
            if not (0 <= i < t.shape[k]):
                return np.zeros(t.shape[len(ixs):], dtype=t.dtype)
        return t[tuple(ixs)]

    # tsiota (OpsConcrete.hs:536)
    def siota(self, n, dtype=np.int64): return np.arange(n).astype(dtype)

    # --------------------------------------------------------- elementwise
    # derived Num/Floating instances (CarriersConcrete.hs:333-343); op codes Ast.hs:707-726
    def unary(self, op, t):
        t = np.asarray(t)
        with np.errstate(all="ignore"):
            if op == "negate": return -t
            if op == "abs": return np.abs(t)
            if op == "signum":
                # Haskell: signum x | x > 0 = 1 | x < 0 = -1 | otherwise = x
                return np.where(t > 0, np.ones_like(t), np.where(t < 0, -np.ones_like(t), t))
            if op == "recip": return (np.ones((), dtype=t.dtype) / t).astype(t.dtype)
            f = {"exp": np.exp, "log": np.log, "sqrt": np.sqrt, "sin": np.sin, "cos": np.cos,
                 "tan": np.tan, "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan,
                 "sinh": np.sinh, "cosh": np.cosh, "tanh": np.tanh, "asinh": np.arcsinh,
                 "acosh": np.arccosh, "atanh": np.arctanh}[op]
            return f(t).astype(t.dtype)

    def binary(self, op, a, b):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
        with np.errstate(all="ignore"):
            if op == "add": return a + b
            if op == "sub": return a - b
            if op == "mul": return a * b
            if op == "divide": return (a / b).astype(a.dtype)
            if op == "power": return np.power(a, b).astype(a.dtype)
            if op == "logbase": return (np.log(b) / np.log(a)).astype(a.dtype)  # logBase x y
            if op == "atan2": return np.arctan2(a, b).astype(a.dtype)
            if op == "max": return np.where(a < b, b, a)
            if op == "min": return np.where(b < a, b, a)
            if op in ("quot", "rem"):
                # array quotH/remH: zero divisors replaced by 1 first
                # (CarriersConcrete.hs:40-56,140-144); quot/rem truncate toward zero
                d = np.where(b == 0, np.ones_like(b), b)
                q = (np.abs(a) // np.abs(d)) * np.sign(a) * np.sign(d)
                q = q.astype(a.dtype)
                return q if op == "quot" else (a - q * d).astype(a.dtype)
        raise ValueError(op)

    # tscast = realToFrac, tsfromIntegral = fromIntegral (OpsConcrete.hs:455-532)
    def scast(self, t, dtype):
        if np.dtype(dtype) == np.bool_:
            return np.asarray(t) != 0
        return np.asarray(t).astype(dtype)

    sfromIntegral = scast

    # tsfloor (OpsConcrete.hs:455-470)
    def sfloor(self, t, dtype=np.int64): return np.floor(t).astype(dtype)

    # EqH / OrdH on whole arrays return ONE Bool (CarriersConcrete.hs:310-331);
    # <= is the Ord of the primitive vector: lexicographic over row-major order
    def all_eq(self, a, b): return bool(np.array_equal(a, b))

    def _lex(self, a, b):
        fa, fb = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
        neq = np.nonzero(fa != fb)[0]
        return None if len(neq) == 0 else (fa[neq[0]], fb[neq[0]])

    def all_leq(self, a, b):
        p = self._lex(a, b)
        return True if p is None else bool(p[0] <= p[1])

    def all_lt(self, a, b):
        p = self._lex(a, b)
        return False if p is None else bool(p[0] < p[1])

    # taddTarget (OpsConcrete.hs:770; UnwindNum.hs:66-118)
    def add_accumulate(self, acc, c): return acc + c

    def fused_map(self, expr, ins, dtype):
        n = int(np.prod(ins[0].shape))
        flat = [np.ascontiguousarray(np.broadcast_to(i, ins[0].shape)).reshape(-1) for i in ins]
        out = ix_eval.evaluate([expr], {}, n, self, inputs=flat, f32=np.dtype(dtype) == np.float32)[0]
        return out.astype(dtype).reshape(ins[0].shape)

    # ---------------------------------------------------- reduce / contract
    # tssum sums the OUTERMOST axis; tssumN repeats it per leading axis,
    # outermost first (OpsConcrete.hs:215-228).  Float accumulates in double.
    def ssum(self, t):
        t = np.asarray(t)
        if t.dtype == np.float32:
            return np.sum(t.astype(np.float64), axis=0).astype(np.float32)
        return np.sum(t, axis=0, dtype=t.dtype)

    def ssumN(self, t, n_axes):
        t = np.asarray(t)
        if n_axes == 0:
            return t
        ax = tuple(range(n_axes))
        if t.dtype == np.float32:
            return np.sum(t.astype(np.float64), axis=ax).astype(np.float32)
        return np.sum(t, axis=ax, dtype=t.dtype)

    # tssum0 (OpsConcrete.hs:230)
    def ssum0(self, t): return self.ssumN(t, np.asarray(t).ndim)

    # tsdot0 (OpsConcrete.hs:232)
    def sdot0(self, a, b):
        assert a.shape == b.shape
        return self.ssum0(self._mul_wide(a, b))

    def _mul_wide(self, a, b):
        if a.dtype == np.float32:
            return (a.astype(np.float64) * b.astype(np.float64))
        return a * b

    # tsdot1In contracts the INNERMOST axis (OpsConcrete.hs:234-235)
    def sdot1In(self, a, b):
        assert a.shape == b.shape and a.dtype == b.dtype
        r = np.einsum("...k,...k->...", a.astype(np.float64) if a.dtype == np.float32 else a,
                      b.astype(np.float64) if b.dtype == np.float32 else b)
        return r.astype(a.dtype)

    # ssumN @shm (sdot1In a b): the contracted form in TestMnistPP.hs:735
    def sdot1InSumN(self, a, b, n_axes):
        return self.ssumN(self.sdot1In(a, b), n_axes)

    # tsmatvecmul m v = sdot1In m (sreplicate v); tsmatmul2 via broadcast views
    # (OpsConcrete.hs:236-242)
    def smatvecmul(self, m, v): return self.sdot1In(m, self.sreplicate(m.shape[0], v))

    def smatmul2(self, m1, m2):
        a = self.stranspose([1, 0], self.sreplicate(m2.shape[1], m1))
        b = self.stranspose([0, 2, 1], self.sreplicate(m1.shape[0], m2))
        return self.sdot1In(a, b)

    def smatmul2_acc(self, acc, m1, m2):
        """taddTarget acc (tsmatmul2 m1 m2) (DeltaEval.hs:139-148)."""
        return acc + self.smatmul2(m1, m2)

    def smatmul2_sum(self, a1, b1, a2, b2):
        """tsmatmul2 a1 b1 + tsmatmul2 a2 b2 (the pre-activation of rnnMnistLayerS, MnistRnnShaped2.hs:55-69)."""
        return self.smatmul2(a1, b1) + self.smatmul2(a2, b2)

    def smatmul2_pair(self, a1, b1, a2, b2):
        return self.smatmul2(a1, b1), self.smatmul2(a2, b2)

    def smatmul2_acc_pair(self, acc1, a1, b1, acc2, a2, b2):
        return acc1 + self.smatmul2(a1, b1), acc2 + self.smatmul2(a2, b2)

    def smatmul2_sum_tanh(self, a1, b1, a2, b2, bias):
        """rnnMnistLayerS (MnistRnnShaped2.hs:55-69): tanh (wX x + wS s + b stretched over the batch)."""
        return np.tanh((self.smatmul2(a1, b1) + self.smatmul2(a2, b2)) + np.asarray(bias)[:, None])

    def smatmul2_list(self, pairs, acc=None):
        """taddTarget folded over the contributions `tsmatmul2 a_t b_t` of one weight (DeltaEval.hs:317-332),
        in arrival order."""
        out = acc
        for a, b in pairs:
            prod = self.smatmul2(a, b)
            out = prod if out is None else out + prod
        return out

    def smatmul2_group(self, results, products):
        """[acc_r + sum of the products naming r (in order), then tanh (. + bias_r) if given]: the transposition /
        forward rules of tsmatmul2 applied one by one (OpsConcrete.hs:234-242, DeltaEval.hs:317-332)."""
        outs = []
        for r, (acc, bias) in enumerate(results):
            out = acc
            for a, b, ri in products:
                if ri == r:
                    prod = self.smatmul2(a, b)
                    out = prod if out is None else out + prod
            outs.append(tanh_affine_fwd(out, None, bias) if bias is not None else out)
        return outs

    # tsargMax / tsargMin reduce the LAST axis, Int result of shape Init sh
    # (OpsConcrete.hs:1712-1746); first extremum (unpinned in the reference)
    def sargMax(self, t):
        if t.shape[-1] == 0:
            return np.zeros(t.shape[:-1], dtype=np.int64)
        return np.argmax(t, axis=-1).astype(np.int64)

    def sargMin(self, t):
        if t.shape[-1] == 0:
            return np.zeros(t.shape[:-1], dtype=np.int64)
        return np.argmin(t, axis=-1).astype(np.int64)

    # --------------------------------------------------- index programs
    def ix_tbl(self, t, ixs):
        from importlib import import_module
        ixm = import_module("horde_ad_b200.ix")  # the shared input language
        return ixm.tbl(t, ixs, _is_float(np.asarray(t)))

    def ix_argmax(self, t, ixs):
        from importlib import import_module
        return import_module("horde_ad_b200.ix").argmax(t, ixs)

    def ix_argmin(self, t, ixs):
        from importlib import import_module
        return import_module("horde_ad_b200.ix").argmin(t, ixs)

    def _targets(self, shm, f, shp):
        """Evaluate f over every position of shm (row-major): linear index into
        shp, or -1 when any component is out of bounds
        (ixsToLinearMaybe, OpsConcrete.hs:944-950)."""
        from importlib import import_module
        ixm = import_module("horde_ad_b200.ix")
        M = int(np.prod(shm)) if len(shm) else 1
        outs = ixm.reify(f, len(shm))
        assert len(outs) == len(shp), "index function returns the wrong number of components"
        env = {}
        if len(shm):
            grids = np.unravel_index(np.arange(M), tuple(shm))
            env = {k: g.astype(np.int64) for k, g in enumerate(grids)}
        comps = ix_eval.evaluate(outs, env, M, self)
        lin = np.zeros(M, dtype=np.int64)
        ok = np.ones(M, dtype=bool)
        for c, n in zip(comps, shp):
            c = c.astype(np.int64)
            ok &= (c >= 0) & (c < n)
            lin = lin * n + np.where(ok, c, 0)
        return np.where(ok, lin, -1)

    # tgatherZS (OpsConcrete.hs:321-323,1621-1665): out[i.., j..] = t[f(i).., j..],
    # zeros where f(i) is out of bounds (:1481-1490)
    def sgather(self, shm, t, f, shp_rank, prog=None):
        t = np.asarray(t)
        shp, shn = t.shape[:shp_rank], t.shape[shp_rank:]
        lin = self._targets(list(shm), f, list(shp))
        P = int(np.prod(shp)) if shp_rank else 1
        flat = np.ascontiguousarray(t).reshape((P,) + tuple(shn))
        out = np.zeros((len(lin),) + tuple(shn), dtype=t.dtype)
        ok = lin >= 0
        if P > 0:
            out[ok] = flat[lin[ok]]
        return out.reshape(tuple(shm) + tuple(shn))

    # tscatterZS (OpsConcrete.hs:318-320,1529-1619): out[f(i).., j..] += t[i.., j..];
    # out-of-bounds dropped; accumulation in row-major order of the source
    # positions, new + old (:1594-1601) -- np.add.at is sequential in that order
    def sscatter(self, shp, t, f, shm_rank, prog=None):
        t = np.asarray(t)
        shm, shn = t.shape[:shm_rank], t.shape[shm_rank:]
        lin = self._targets(list(shm), f, list(shp))
        P = int(np.prod(shp)) if len(shp) else 1
        M = len(lin)
        flat = np.ascontiguousarray(t).reshape((M,) + tuple(shn))
        out = np.zeros((P,) + tuple(shn), dtype=t.dtype)
        for m in range(M) if M <= 4096 else ():
            d = lin[m]
            if d >= 0:
                out[d] = flat[m] + out[d]
        if M > 4096:
            ok = lin >= 0
            np.add.at(out, lin[ok], flat[ok])
        return out.reshape(tuple(shp) + tuple(shn))

    # toneHotS (OpsConcrete.hs:1516-1525)
    def soneHot(self, shp, v, ixs):
        v = np.asarray(v)
        out = np.zeros(tuple(shp) + v.shape, dtype=v.dtype)
        if all(0 <= int(i) < n for i, n in zip(ixs, shp)):
            out[tuple(int(i) for i in ixs)] = v
        return out

    # ----------------------------------------------------- NN blocks
    # conv2dPreservingS (CommonShapedOps.hs:136-157): zero padding on the HIGH side
    def conv2dPreserving(self, K, A):
        Co, Ci, KH, KW = K.shape
        N, Ci2, H, W = A.shape
        assert Ci == Ci2
        Ap = np.zeros((N, Ci, H + KH - 1, W + KW - 1), dtype=A.dtype)
        Ap[:, :, :H, :W] = A
        B = np.zeros((N, Co, H, W), dtype=np.float64)
        for kh in range(KH):
            for kw in range(KW):
                B += np.einsum("nchw,oc->nohw", Ap[:, :, kh:kh + H, kw:kw + W].astype(np.float64),
                               K[:, :, kh, kw].astype(np.float64))
        return B.astype(A.dtype)

    # conv2dPreserving_dKrn (TestConvQuickCheck.hs:322-349)
    def conv2d_dkrn(self, A, dB, KH, KW):
        N, Ci, H, W = A.shape
        Co = dB.shape[1]
        Ap = np.zeros((N, Ci, H + KH - 1, W + KW - 1), dtype=np.float64)
        Ap[:, :, :H, :W] = A
        dK = np.zeros((Co, Ci, KH, KW), dtype=np.float64)
        for kh in range(KH):
            for kw in range(KW):
                dK[:, :, kh, kw] = np.einsum("nohw,nchw->oc", dB.astype(np.float64), Ap[:, :, kh:kh + H, kw:kw + W])
        return dK.astype(A.dtype)

    # conv2dPreserving_dInp (TestConvQuickCheck.hs:290-320)
    def conv2d_dinp(self, K, dB):
        Co, Ci, KH, KW = K.shape
        N, Co2, H, W = dB.shape
        assert Co == Co2
        Bp = np.zeros((N, Co, H + KH - 1, W + KW - 1), dtype=np.float64)
        Bp[:, :, KH - 1:, KW - 1:] = dB
        dA = np.zeros((N, Ci, H, W), dtype=np.float64)
        for kh in range(KH):
            for kw in range(KW):
                dA += np.einsum("nohw,oc->nchw", Bp[:, :, KH - 1 - kh:KH - 1 - kh + H, KW - 1 - kw:KW - 1 - kw + W],
                                K[:, :, kh, kw].astype(np.float64))
        return dA.astype(K.dtype)

    # reluS (CommonShapedOps.hs:38-44)
    def relu(self, t):
        t = np.asarray(t)
        return np.where(t <= 0, np.zeros_like(t), np.ones_like(t)) * t

    # maxPool2dUnpaddedS (CommonShapedOps.hs:241-260) via smaximum (:33-36):
    # value at kargMax of the row-major flattened window; outside = 0 (slicezS)
    def maxpool2d(self, x, ksize, stride):
        N, Cc, H, W = x.shape
        Ho, Wo = H // stride, W // stride
        xp = np.zeros((N, Cc, max(H, (Ho - 1) * stride + ksize), max(W, (Wo - 1) * stride + ksize)), dtype=x.dtype)
        xp[:, :, :H, :W] = x
        win = np.stack([xp[:, :, a:a + Ho * stride:stride, b:b + Wo * stride:stride][:, :, :Ho, :Wo]
                        for a in range(ksize) for b in range(ksize)], axis=-1)
        am = np.argmax(win, axis=-1).astype(np.int64)
        y = np.take_along_axis(win, am[..., None], axis=-1)[..., 0]
        return y, am

    def maxpool2d_bwd(self, dy, am, ksize, stride, H, W):
        N, Cc, Ho, Wo = dy.shape
        dx = np.zeros((N, Cc, H, W), dtype=dy.dtype)
        for i in range(Ho):      # row-major window order = scatter source order
            for j in range(Wo):
                a, b = am[:, :, i, j] // ksize, am[:, :, i, j] % ksize
                yy, xx = i * stride + a, j * stride + b
                ok = (yy < H) & (xx < W)
                n_idx, c_idx = np.nonzero(ok)
                np.add.at(dx, (n_idx, c_idx, yy[ok], xx[ok]), dy[:, :, i, j][ok])
        return dx

    # the tail of convMnistLayerS (MnistCnnShaped2.hs:42-62) for ksize == stride,
    # composed from the oracle's own bias add, reluS and maxPool2dUnpaddedS
    def relu_pool_fwd(self, pre, bias, ksize):
        pre = np.asarray(pre)
        v = pre if bias is None else pre + np.asarray(bias).reshape(1, -1, 1, 1)
        y, am = self.maxpool2d(self.relu(v), ksize, ksize)
        N, Cc, Ho, Wo = y.shape
        win = np.stack([v[:, :, a:a + Ho * ksize:ksize, b:b + Wo * ksize:ksize][:, :, :Ho, :Wo]
                        for a in range(ksize) for b in range(ksize)], axis=-1)
        sel = np.take_along_axis(win, am[..., None], axis=-1)[..., 0]
        code = (am | np.where(sel > 0, 0x80, 0)).astype(np.uint8).view(np.int8)
        return y, code

    def relu_pool_bwd(self, dy, code, ksize, H, W, want_bias=True):
        code = np.asarray(code).view(np.uint8)
        am = (code & 0x7f).astype(np.int64)
        g = np.where(code & 0x80, dy, np.zeros_like(dy))
        dpre = self.maxpool2d_bwd(g, am, ksize, ksize, H, W)
        return dpre, (np.sum(dpre, axis=(0, 2, 3)) if want_bias else None)

    def conv_relu_pool_fwd(self, K, bias, A, ksize):
        return self.relu_pool_fwd(self.conv2dPreserving(K, A), bias, ksize)

    # rnnMnistLayerS behind its matmuls (module-level restatements below)
    def tanh_affine_fwd(self, u, v, bias):
        return tanh_affine_fwd(u, v, bias)

    def tanh_affine_bwd(self, y, c, want_bias=True):
        dz, db = tanh_affine_bwd(y, c)
        return dz, (db if want_bias else None)

    def tanh_affine_bwd_acc(self, y, c, acc):
        dz, db = tanh_affine_bwd(y, c)
        return dz, acc + db

    def conv_relu_pool_bwd(self, K, A, dy, code, ksize, want_bias=True, want_dA=True):
        dpre, db = self.relu_pool_bwd(dy, code, ksize, A.shape[2], A.shape[3], want_bias)
        return (self.conv2d_dkrn(A, dpre, K.shape[2], K.shape[3]), db,
                self.conv2d_dinp(K, dpre) if want_dA else None)

    # lossSoftMaxCrossEntropyS (CommonShapedOps.hs:89-111): normalises over the
    # WHOLE tensor; dual = (softMax - expected) * du
    def softmax_xent(self, logits, expected, scale=1.0, want_grad=True):
        u = np.asarray(logits)
        expU = np.exp(u - np.min(u))
        sm = (1.0 / np.sum(expU.astype(np.float64))).astype(u.dtype) * expU
        loss = -np.sum((np.log(sm) * expected).astype(np.float64))
        sc = np.array(scale, dtype=u.dtype)
        return (sc * loss.astype(u.dtype)), ((sc * (sm - expected)).astype(u.dtype) if want_grad else None)

    # sfold (*) 1 x and its reverse derivative (Ops.hs:188-210,
    # OpsConcrete.hs:990-1019, DeltaEval.hs:351-389): sequential, in order
    def prod_fold_grad(self, x, dt=1.0, want_grad=True):
        x = np.asarray(x)
        n = x.shape[0]
        acc = np.empty(n + 1, dtype=x.dtype)
        acc[0] = 1
        for i in range(n):
            acc[i + 1] = acc[i] * x[i]
        if not want_grad:
            return acc[n], None
        g = np.empty(n, dtype=x.dtype)
        c = np.array(dt, dtype=x.dtype)
        for i in range(n - 1, -1, -1):
            g[i] = c * acc[i]
            c = c * x[i]
        return acc[n], g

    # updateWithGradient (OptimizerTools.hs:20-51)
    def sgd_step(self, p, g, gamma):
        p -= np.array(gamma, dtype=p.dtype) * g

    # updateWithGradientAdam (OptimizerTools.hs:134-232; defaults :83-89)
    def adam_step(self, p, m, v, g, t, alpha=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-7):
        T = p.dtype.type
        alphat = alpha * np.sqrt(1 - _pow_ghc(beta2, t)) / (1 - _pow_ghc(beta1, t))
        m[...] = T(beta1) * m + T(1 - beta1) * g
        v[...] = T(beta2) * v + T(1 - beta2) * (g * g)
        p[...] = p - (T(alphat) * m) / (np.sqrt(v) + T(epsilon))


# ---------------------------------------------------------------- recurrent layer body
def tanh_affine_fwd(u, v, bias):
    """rnnMnistLayerS behind its matmuls (example/MnistRnnShaped2.hs:55-69): tanh ((u + v) + stretched bias)."""
    z = u + v if v is not None else u
    if bias is not None:
        z = z + np.asarray(bias)[:, None]
    return np.tanh(z)


def tanh_affine_bwd(y, c):
    """Its adjoint: dz = (1 - y * y) * c (cotangent of both u and v), dbias = row sums of dz."""
    dz = (1 - y * y) * c
    return dz, dz.sum(axis=1)


# ---------------------------------------------------------------- MNIST input pipeline
def mnist_batch(pixels, labels, order, start, batch, classes=10, dtype=np.float64):
    """readMnistData + mkMnistDataBatchS (example/MnistData.hs:127-152): for the samples
    order[start : start + batch] (order None = file order) glyph = fromIntegral pixel / 255 and
    target = one-hot label; sample ids outside [0, n) give zeros (gather's out-of-bounds rule)."""
    pixels, labels = np.asarray(pixels, np.uint8), np.asarray(labels, np.uint8)
    n = pixels.shape[0]
    ids = np.arange(start, start + batch) if order is None else np.asarray(order)[start:start + batch]
    glyphs = np.zeros((batch,) + pixels.shape[1:], dtype)
    targets = np.zeros((batch, classes), dtype)
    for b, i in enumerate(ids):
        if 0 <= i < n:
            glyphs[b] = pixels[i].astype(dtype) / np.dtype(dtype).type(255)
            if labels[i] < classes:
                targets[b, labels[i]] = 1
    return glyphs, targets

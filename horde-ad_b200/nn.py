"""Target-generic NN blocks: mirror of src/HordeAd/External/CommonShapedOps.hs.

Each block exists in two forms, both written against the carrier method
surface (so they run over the device carrier, the CPU oracle, or either wrapped
in ADBackend):

* `*_vectorised`: the op sequence horde-ad's vectoriser + simplifier produce
  from the reference definitions (`kbuild` over `slicezS`/`sdot0`/`smaximum`),
  i.e. what `interpretAst` actually sends to a backend: chained affine
  `sgather`s (im2col), stride-0 `sreplicate`/`stranspose` views, `sdot1In`,
  `ssumN`, data-dependent gathers with `kargMax` and `ifH`
  (TestMnistPP.hs:735; bench/ConvVjpBench.hs:428-555).
* fused: the nodes a device-specific contraction pass (sibling of
  `contractAst`, AstTraverse.hs:290-699) would emit instead:
  one C-ABI call per block.
"""
from __future__ import annotations

import numpy as np

from . import ix as ixm


def _mul(T, a, b): return T.binary("mul", a, b)
def _add(T, a, b): return T.binary("add", a, b)
def _sub(T, a, b): return T.binary("sub", a, b)


# ------------------------------------------------------------------ relu
def reluS(T, v):
    """reluS (CommonShapedOps.hs:38-44), fused."""
    return T.relu(v)


def reluS_vectorised(T, v):
    """oneIfGtZero * v with the mask as the artifact builds it: a gather from
    the constant [0.0, 1.0] with a data-dependent index
    `ifH (0.0 <=. negate (v sindex0 ix)) 0 1` (TestMnistPP.hs:735)."""
    sh = list(T.shape(v))
    dt = T.dtype(v)
    table = T.sconcrete(np.array([0.0, 1.0], dtype=dt))
    plain = T.sprimalPart(v) if hasattr(T, "sprimalPart") else v

    def f(ix):
        x = T.ix_tbl(plain, ix)
        return [ixm.ifH(ixm.Ix.lift(0.0) <= -x, 0, 1)]
    mask = T.sgather(sh, table, f, 1)
    return _mul(T, mask, v)


# ------------------------------------------------------------------ conv
def conv2dPreservingS(T, K, A):
    """conv2dPreservingS (CommonShapedOps.hs:136-157), fused."""
    return T.conv2dPreserving(K, A)


def conv2dPreservingS_vectorised(T, K, A):
    """The vectorised form: two chained affine gathers `[i + j]` build the
    patch tensor (zeros beyond the high edge come from the gather's
    out-of-bounds rule), then `ssumN . sdot1In` over broadcast views contracts
    (c, kh, kw) (bench/ConvVjpBench.hs:428-555; TestMnistPP.hs:735)."""
    Co, Ci, KH, KW = T.shape(K)
    N, _, H, W = T.shape(A)
    # rows: [H, KH, N, Ci, W] from A transposed to [H, N, Ci, W]
    a_h = T.stranspose([2, 0, 1, 3], A)
    g1 = T.sgather([H, KH], a_h, lambda ix: [ix[0] + ix[1]], 1)
    # cols: bring W to the front: [W, H, KH, N, Ci] -> gather -> [W, KW, H, KH, N, Ci]
    g1t = T.stranspose([4, 0, 1, 2, 3], g1)
    g2 = T.sgather([W, KW], g1t, lambda ix: [ix[0] + ix[1]], 1)
    # patches as [Ci, KH, N, Co, H, W, KW] (Co by replication)
    p = T.stranspose([5, 3, 4, 2, 0, 1], g2)          # [Ci, KH, N, H, W, KW]
    p = T.stranspose([1, 2, 3, 0], T.sreplicate(Co, p))  # [Ci, KH, N, Co, H, W, KW]
    k = T.stranspose([1, 2, 0, 3], K)                 # [Ci, KH, Co, KW]
    k = T.sreplicateN([N, H, W], k)                   # [N, H, W, Ci, KH, Co, KW]
    k = T.stranspose([3, 4, 0, 5, 1, 2], k)           # [Ci, KH, N, Co, H, W, KW]
    return T.sdot1InSumN(p, k, 2) if hasattr(T, "sdot1InSumN") else T.ssumN(T.sdot1In(p, k), 2)


def convReluPoolS(T, K, bias, A, ksize):
    """convMnistLayerS (MnistCnnShaped2.hs:42-62) as one node: maxPool (relu
    (conv2dPreserving K A + bias)) with ksize == stride."""
    if hasattr(T, "conv_relu_pool"):
        return T.conv_relu_pool(K, bias, A, ksize)
    return T.relu_pool_fwd(T.conv2dPreserving(K, A), bias, ksize)[0]


# --------------------------------------------------------------- maxpool
def maxPool2dUnpaddedS(T, ksize, stride, x):
    """maxPool2dUnpaddedS (CommonShapedOps.hs:241-260), fused."""
    return T.maxpool2d(x, ksize, stride)


def maxPool2dUnpaddedS_vectorised(T, ksize, stride, x):
    """Windows by a gather with table lookups `[tblH[i, a], tblW[j, b]]`
    (tbl = stride*i + a as a constant Int tensor, as the simplifier prints it),
    then `smaximum` = gather at `kargMax` of the flattened window."""
    N, Cc, H, W = T.shape(x)
    Ho, Wo = H // stride, W // stride
    th = T.sconcrete((stride * np.arange(Ho)[:, None] + np.arange(ksize)[None, :]).astype(np.int64))
    tw = T.sconcrete((stride * np.arange(Wo)[:, None] + np.arange(ksize)[None, :]).astype(np.int64))
    xt = T.stranspose([2, 3, 0, 1], x)                # [H, W, N, C]
    win = T.sgather([Ho, Wo, ksize, ksize], xt,
                    lambda ix: [T.ix_tbl(th, [ix[0], ix[2]]), T.ix_tbl(tw, [ix[1], ix[3]])], 2)
    win = T.stranspose([4, 5, 0, 1, 2, 3], win)       # [N, C, Ho, Wo, ks, ks]
    u = T.sreshape([N, Cc, Ho, Wo, ksize * ksize], win)
    plain = T.sprimalPart(u) if hasattr(T, "sprimalPart") else u
    return T.sgather([N, Cc, Ho, Wo], u,
                     lambda ix: [ix[0], ix[1], ix[2], ix[3], T.ix_argmax(plain, ix)], 5)


# ------------------------------------------------------------------ loss
def lossSoftMaxCrossEntropyS(T, expected, d, scale=1.0):
    """lossSoftMaxCrossEntropyS (CommonShapedOps.hs:89-111), fused with its
    custom dual; `scale` folds the `recip batch_size *` of the callers
    (MnistCnnShaped2.hs:136)."""
    return T.softmax_xent(d, expected, scale)[0]


def lossSoftMaxCrossEntropyS_vectorised(T, expected, d, scale=1.0):
    """The same loss op by op (primal only; the reference attaches the dual
    with tD, so differentiation goes through the fused form)."""
    u = d
    sh = list(T.shape(u))
    flat = T.sreshape([int(np.prod(sh))], u)
    umin = T.sindex(flat, [int(T.kfromS(T.sargMin(flat)))])       # sminimum (:28-31)
    expU = T.unary("exp", _sub(T, u, T.sreplicate0N(sh, umin)))
    recip_sum = T.unary("recip", T.ssum0(expU))
    soft = _mul(T, T.sreplicate0N(sh, recip_sum), expU)
    loss = T.unary("negate", T.sdot0(T.unary("log", soft), expected))
    return _mul(T, T.sscalar(scale, T.dtype(u)), loss)


# ------------------------------------------------- dense-network pieces
def logistic(T, d):
    """logistic (CommonRankedOps.hs:55-63): y = recip (1 + exp (-d)) with the
    custom dual y * (1 - y) * dd (not the derivative AD would assemble from
    the three ops)."""
    def primal(B, x):
        one = B.srepl(list(B.shape(x)), 1, B.dtype(x))
        return B.unary("recip", B.binary("add", one, B.unary("exp", B.unary("negate", x))))

    def vjp(B, x, y, c):
        one = B.srepl(list(B.shape(y)), 1, B.dtype(y))
        return B.binary("mul", B.binary("mul", y, B.binary("sub", one, y)), c)
    if hasattr(T, "custom_unary"):
        return T.custom_unary(d, primal, vjp)
    return primal(T, d)


def softMax1(T, d):
    """softMax1 (CommonRankedOps.hs:139-145)."""
    expU = T.unary("exp", d)
    return _mul(T, T.sreplicate0N(list(T.shape(d)), T.unary("recip", T.ssum0(expU))), expU)


def lossCrossEntropyV(T, targ, res):
    """lossCrossEntropyV (CommonRankedOps.hs:89-92)."""
    return T.unary("negate", T.sdot0(T.unary("log", res), targ))


# ------------------------------------------------ folds (tmapAccumL family)
def smapAccumL(T, f, acc0, es):
    """tmapAccumL (Ops.hs:987-1125) as the Concrete backend runs it
    (tmapAccumLC, OpsConcrete.hs:990-1019): a host loop over the slices of
    `es` along its leading axis, left to right; returns (acc, stacked bs).
    Over an AD backend the unrolled ops record their own delta nodes, which
    transposes to the same cotangents as the reference's tmapAccumR rule
    (DeltaEval.hs:351-389)."""
    acc, bs = acc0, []
    for e in T.sunravelToList(es):
        acc, b = f(acc, e)
        if b is not None:
            bs.append(b)
    return acc, (T.sfromVector(bs) if bs else None)


def _step_is_product(T, f, acc0, es) -> bool:
    """Is the fold's step function `\\acc e -> acc * e` (or `e * acc`)?

    horde-ad hands `tmapAccumL` its step as a function value (an `HFun`, Ops.hs:1106-1125); a device instance can
    look inside it.  Here the step is a Python closure, and it is reified the same way index functions are
    (option A): it is applied ONCE to two 0-d probes while the library RECORDS instead of
    launching (hb_lazy_begin / hb_lazy_end); the step is the product iff the recording is exactly one
    `binary.mul` of the two probes.  Only device carriers can record; everything else folds on the host."""
    base = getattr(T, "B", T)
    if not hasattr(base, "record") or not hasattr(T, "prod_fold"):
        return False
    sh = T.shape(es)
    if len(sh) != 1 or len(T.shape(acc0)) != 0 or np.dtype(T.dtype(es)).kind != "f":
        return False
    dt = T.dtype(es)
    pa, pe = base.sscalar(2.0, dt), base.sscalar(3.0, dt)
    try:
        prog = base.record(lambda: getattr(T, "primal", lambda x: x)(f(T.sfromPrimal(pa) if hasattr(T, "sfromPrimal") else pa,
                                                                  T.sfromPrimal(pe) if hasattr(T, "sfromPrimal") else pe)),
                           rewrite=False)
    except Exception:       # the step does something that cannot be recorded (needs a value on the host, ...)
        return False
    lines = [ln for ln in prog.dump().split("--")[0].strip().splitlines() if ln]
    rec, _ = prog.stats
    ok = rec == 1 and len(lines) == 1 and lines[0].split()[1] == "binary.mul" and lines[0].count("<-const") == 2
    # the two operands must be the two probes (not the same probe twice): distinguishable by evaluating once
    if ok and not getattr(base, "trace_only", False):
        try:
            prog.run()
            ok = float(base.kfromS(prog.outs[0])) == 6.0
        except Exception:
            ok = False
    prog.close()
    return ok


def sfold(T, f, acc0, es):
    """sfold / rfold (Ops.hs:188-210): left fold of f :: acc -> e -> acc.  `sfold (*) acc0 es` (LongProdBench,
    bench/common/BenchProdTools.hs:154-182) is recognised from its step function and runs as the product-scan
    kernel (hb_prod_fold_grad) instead of n host iterations; every other step folds on the host like the reference
    (tmapAccumLC, OpsConcrete.hs:990-1019)."""
    if _step_is_product(T, f, acc0, es):
        return T.binary("mul", acc0, T.prod_fold(es))
    return smapAccumL(T, lambda a, e: (f(a, e), None), acc0, es)[0]


def sscan(T, f, acc0, es):
    """sscan (Ops.hs:66-90 of the Concrete instance): acc0 followed by every
    intermediate accumulator, stacked."""
    accs = [acc0]
    for e in T.sunravelToList(es):
        accs.append(f(accs[-1], e))
    return T.sfromVector(accs)


# -------------------------------- per-element builds (non-vectorised code)
def sbuild1(T, k, f):
    """tsbuild1 (Ops.hs:910; OpsConcrete.hs:1761-1789): stack f 0 .. f (k-1).
    Only reached by code the vectoriser did not touch; a host loop by design."""
    return T.sfromVector([f(i) for i in range(k)])


def smap0N(T, f, t):
    """tsmap0N (Ops.hs:929): per-element host function over 0-d slices."""
    sh = list(T.shape(t))
    flat = T.sreshape([int(np.prod(sh))], t)
    out = T.sfromVector([f(x) for x in T.sunravelToList(flat)])
    return T.sreshape(sh, out)


def scond(T, b: bool, u, v):
    """scond / tcond (OpsConcrete.hs:113-119): a host `if` on a host Bool."""
    return u if b else v


def szipWith0N(T, f, a, b):
    """tszipWith0N (Ops.hs:934): per-element host function over two tensors."""
    sh = list(T.shape(a))
    n = int(np.prod(sh))
    fa, fb = T.sreshape([n], a), T.sreshape([n], b)
    out = T.sfromVector([f(x, y) for x, y in zip(T.sunravelToList(fa), T.sunravelToList(fb))])
    return T.sreshape(sh, out)


def sbuild(T, shm, f):
    """tsbuild @shm (Ops.hs:913): f over every index of shm, row-major, results
    of shape shn placed at their index: shm ++ shn."""
    shm = list(shm)
    if not shm:
        return f(())
    import itertools
    parts = [f(ix) for ix in itertools.product(*[range(k) for k in shm])]
    stacked = T.sfromVector(parts)                        # [prod shm] ++ shn
    return T.sreshape(shm + list(T.shape(parts[0])), stacked)


def sfromVectorLinear(T, shm, scalars):
    """tsfromVectorLinear (Ops.hs:429): prod shm 0-d tensors -> shm."""
    return T.sreshape(list(shm), T.sfromVector(list(scalars)))


def stoListLinear(T, t):
    """tstoListLinear (Ops.hs:439): the 0-d elements of t in row-major order."""
    n = int(np.prod(T.shape(t))) if len(T.shape(t)) else 1
    return T.sunravelToList(T.sreshape([n], t))


def sminimum(T, t):
    """sminimum (CommonShapedOps.hs:28-31): t ! argMin of the flattened tensor."""
    flat = T.sreshape([int(np.prod(T.shape(t)))], t)
    return T.sindex(flat, [int(T.kfromS(T.sargMin(flat)))])


def smaximum(T, t):
    """smaximum (CommonShapedOps.hs:33-36)."""
    flat = T.sreshape([int(np.prod(T.shape(t)))], t)
    return T.sindex(flat, [int(T.kfromS(T.sargMax(flat)))])

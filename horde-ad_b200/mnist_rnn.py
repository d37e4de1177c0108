"""MnistRnnShaped2 (example/MnistRnnShaped2.hs:25-139) against the carrier
method surface: two recurrent tanh layers unrolled over the 28 rows of the
glyph (`unrollLastS` = a host fold over `sunravelToList`, :45-52), a dense
read-out, and the fused softmax-cross-entropy loss."""
from __future__ import annotations

import numpy as np

from . import nn

SIZE_MNIST_LABEL = 10


def param_shapes(out_width, h=28):
    """ADRnnMnistParametersShaped (MnistRnnShaped2.hs:25-36)."""
    return [(out_width, h), (out_width, out_width), (out_width,),
            (out_width, out_width), (out_width, out_width), (out_width,),
            (SIZE_MNIST_LABEL, out_width), (SIZE_MNIST_LABEL,)]


def init_params(out_width, seed=44, rng_range=0.4, dtype=np.float64, h=28):
    rng = np.random.Generator(np.random.Philox(seed))
    return [(2 * rng_range * (rng.random(s) - 0.5)).astype(dtype) for s in param_shapes(out_width, h)]


def rnnMnistLayerS(T, s, x, wX, wS, b, fused=False):
    """rnnMnistLayerS (MnistRnnShaped2.hs:55-69): tanh (wX x + wS s + b).  `fused`: the two adds, the
    stretched bias and tanh as the one contraction-pass node `tanh_affine`; `fused == "pair"`: additionally
    `wX x + wS s` as the node `smatmul2_sum`."""
    batch = T.shape(x)[1]
    if fused == "pair":   # the two products, their sum, the bias and tanh as ONE node (one paired GEMM launch)
        if hasattr(T, "smatmul2_sum_tanh"):
            return T.smatmul2_sum_tanh(wX, x, wS, s, b)
        return T.tanh_affine(T.smatmul2_sum(wX, x, wS, s), None, b)
    if fused:
        return T.tanh_affine(T.smatmul2(wX, x), T.smatmul2(wS, s), b)
    y = T.binary("add", T.binary("add", T.smatmul2(wX, x), T.smatmul2(wS, s)),
                 T.stranspose([1, 0], T.sreplicate(batch, b)))
    return T.unary("tanh", y)


def rnnMnistTwoS(T, s, x, layers, fused=False):
    """rnnMnistTwoS (MnistRnnShaped2.hs:73-99): state = sappend vec1 vec2."""
    (wX, wS, b), (wX2, wS2, b2) = layers
    out_width = T.shape(wS)[0]
    s1, s2 = T.sslice(0, out_width, s), T.sslice(out_width, out_width, s)
    vec1 = rnnMnistLayerS(T, s1, x, wX, wS, b, fused)
    vec2 = rnnMnistLayerS(T, s2, vec1, wX2, wS2, b2, fused)
    s3 = T.sappend(vec1, vec2)
    return T.sslice(out_width, out_width, s3), s3


def rnnMnistZeroS(T, xs, params, fused=False):
    """rnnMnistZeroS (MnistRnnShaped2.hs:103-117): zero state, unrollLastS."""
    wX, wS, b, wX2, wS2, b2, w3, b3 = params
    out_width = T.shape(wS)[0]
    batch = T.shape(xs)[2]
    if fused == "wave" and hasattr(T, "smatmul2_sum_tanh_2"):
        # the same dataflow in WAVEFRONT order: layer 1 at row t and layer 2 at row t - 1 read only values of earlier
        # waves (vec1 of row t - 1, vec2 of row t - 2), so the two layer bodies share one launch (four products, two
        # tanh epilogues).  The state `sappend vec1 vec2` of rnnMnistTwoS is never materialised: each half is only
        # ever sliced back out.  Every value is bit-identical to the row-by-row schedule's.
        rows = T.sunravelToList(xs)
        zero = T.srepl([out_width, batch], 0, T.dtype(wS))
        vec1 = T.smatmul2_sum_tanh(wX, rows[0], wS, zero, b)
        vec2 = zero
        for x in rows[1:]:
            vec1, vec2 = T.smatmul2_sum_tanh_2((wX, x, wS, vec1, b), (wX2, vec1, wS2, vec2, b2))
        out = T.smatmul2_sum_tanh(wX2, vec1, wS2, vec2, b2)
        return T.binary("add", T.smatmul2(w3, out), T.stranspose([1, 0], T.sreplicate(batch, b3)))
    if fused == "wave":
        fused = "pair"
    s = T.srepl([2 * out_width, batch], 0, T.dtype(wS))
    out = None
    for x in T.sunravelToList(xs):          # foldl' over the rows (:45-52)
        out, s = rnnMnistTwoS(T, s, x, ((wX, wS, b), (wX2, wS2, b2)), fused)
    return T.binary("add", T.smatmul2(w3, out), T.stranspose([1, 0], T.sreplicate(batch, b3)))


def rnnMnistLossFusedS(T, glyphs, labels, params, fused=False):
    """rnnMnistLossFusedS (MnistRnnShaped2.hs:119-139): glyphs [batch, h, w],
    labels [batch, 10]."""
    batch = T.shape(glyphs)[0]
    xs = T.stranspose([2, 1, 0], glyphs)
    if fused and hasattr(T, "scontiguous"):
        # the transposed glyph batch is materialised ONCE (6 MB) instead of every per-step row `xs !! t` being packed
        # inside the products that read it (twice per step: forward product and weight cotangent)
        xs = T.scontiguous(xs)
    result = rnnMnistZeroS(T, xs, params, fused)
    return nn.lossSoftMaxCrossEntropyS(T, T.stranspose([1, 0], labels), result, scale=1.0 / batch)

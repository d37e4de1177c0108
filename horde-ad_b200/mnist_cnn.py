"""MnistCnnShaped2 (example/MnistCnnShaped2.hs) written against the carrier
method surface, plus the training step the benches time.

`convMnistTwoS` / `convMnistLossFusedS` follow the reference definitions line
by line (MnistCnnShaped2.hs:42-136).  `fused=True` ("layer") selects the coarsest
nodes a device contraction pass emits (a whole conv layer: conv, then one
fused bias + relu + max-pool pass), `fused="nodes"` conv / relu / maxpool as
single C-ABI calls each, `fused=False` the vectorised gather programs the
stock pipeline produces.
"""
from __future__ import annotations

import numpy as np

from . import nn

SIZE_MNIST_LABEL = 10
SIZE_MNIST_H = 28
SIZE_MNIST_W = 28


def param_shapes(kh, kw, c_out, n_hidden, h=SIZE_MNIST_H, w=SIZE_MNIST_W):
    """ADCnnMnistParametersShaped (MnistCnnShaped2.hs:30-40); kh, kw denote
    kernel height/width MINUS ONE as in the reference."""
    return [
        (c_out, 1, kh + 1, kw + 1), (c_out,),
        (c_out, c_out, kh + 1, kw + 1), (c_out,),
        (n_hidden, c_out * (h // 4) * (w // 4)), (n_hidden,),
        (SIZE_MNIST_LABEL, n_hidden), (SIZE_MNIST_LABEL,),
    ]


def init_params(kh, kw, c_out, n_hidden, seed=44, rng_range=0.4, dtype=np.float64,
                h=SIZE_MNIST_H, w=SIZE_MNIST_W):
    """randomValue 0.4 (Adaptor.hs:166-179): 2*range*(u - 0.5), u ~ U[0,1).
    The Haskell StdGen stream is not reproducible here; Philox seeded 44."""
    rng = np.random.Generator(np.random.Philox(seed))
    return [(2 * rng_range * (rng.random(s) - 0.5)).astype(dtype) for s in param_shapes(kh, kw, c_out, n_hidden, h, w)]


def convMnistLayerS(T, ker, x, bias, fused=True):
    """convMnistLayerS (MnistCnnShaped2.hs:42-62)."""
    N, _, H, W = T.shape(x)
    if fused is True or fused == "layer":     # the whole layer as one contraction-pass node
        return nn.convReluPoolS(T, ker, bias, x, 2)
    yConv = (nn.conv2dPreservingS if fused else nn.conv2dPreservingS_vectorised)(T, ker, x)
    biasStretched = T.stranspose([0, 3, 1, 2], T.sreplicateN([N, H, W], bias))
    pre = T.binary("add", yConv, biasStretched)
    yRelu = (nn.reluS if fused else nn.reluS_vectorised)(T, pre)
    if fused:
        return T.maxpool2d(yRelu, 2, 2)
    return nn.maxPool2dUnpaddedS_vectorised(T, 2, 2, yRelu)


def _pool_value(r):
    return r[0] if isinstance(r, tuple) else r


def convMnistTwoS(T, x, params, fused=True):
    """convMnistTwoS (MnistCnnShaped2.hs:64-109): x : [batch, 1, h, w] ->
    [SizeMnistLabel, batch]."""
    ker1, bias1, ker2, bias2, wD, bD, wR, bR = params
    N = T.shape(x)[0]
    t1 = _pool_value(convMnistLayerS(T, ker1, x, bias1, fused))
    t2 = _pool_value(convMnistLayerS(T, ker2, t1, bias2, fused))
    m1 = T.sreshape([N, int(np.prod(T.shape(t2)[1:]))], t2)
    denseLayer = T.binary("add", T.smatmul2(wD, T.stranspose([1, 0], m1)),
                          T.stranspose([1, 0], T.sreplicate(N, bD)))
    hidden = (nn.reluS if fused else nn.reluS_vectorised)(T, denseLayer)
    return T.binary("add", T.smatmul2(wR, hidden), T.stranspose([1, 0], T.sreplicate(N, bR)))


def convMnistLossFusedS(T, glyphs, labels, params, fused=True, loss_scale=None):
    """convMnistLossFusedS (MnistCnnShaped2.hs:111-136): glyphs [batch, h, w],
    labels [batch, 10] one-hot; returns the 0-d loss.  `loss_scale` replaces the
    `recip batch` factor (:136) when the tensor at hand is one shard of a larger
    batch whose objective is normalised globally (train.py, objective="global")."""
    N, H, W = T.shape(glyphs)
    x = T.sreshape([N, 1, H, W], glyphs)
    result = convMnistTwoS(T, x, params, fused)
    targets = T.stranspose([1, 0], labels)
    return nn.lossSoftMaxCrossEntropyS(T, targets, result, scale=(1.0 / N if loss_scale is None else loss_scale))

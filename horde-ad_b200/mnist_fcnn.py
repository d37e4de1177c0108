"""MnistFcnnRanked2 (example/MnistFcnnRanked2.hs:25-76) against the carrier
method surface: two hidden layers with `logistic`, `softMax1` read-out and
`lossCrossEntropyV`; the second hidden layer may hold its weights in a
different precision `q` (Float in the reference's benches,
BenchMnistTools.hs:261-265) with `rcast` on both sides (:63)."""
from __future__ import annotations

import numpy as np

from . import nn

SIZE_MNIST_GLYPH = 784
SIZE_MNIST_LABEL = 10


def param_shapes(width_hidden, width_hidden2):
    """ADFcnnMnist2ParametersShaped (MnistFcnnRanked2.hs:25-33)."""
    return [(width_hidden, SIZE_MNIST_GLYPH), (width_hidden,),
            (width_hidden2, width_hidden), (width_hidden2,),
            (SIZE_MNIST_LABEL, width_hidden2), (SIZE_MNIST_LABEL,)]


def init_params(width_hidden, width_hidden2, seed=44, rng_range=1.0, r=np.float64, q=np.float32):
    """randomValue 1 (BenchMnistTools.hs:249-254); Philox instead of StdGen."""
    rng = np.random.Generator(np.random.Philox(seed))
    dts = [r, r, q, r, r, r]
    return [(2 * rng_range * (rng.random(s) - 0.5)).astype(dt) for s, dt in zip(param_shapes(width_hidden, width_hidden2), dts)]


def afcnnMnist2(T, datum, params, act_hidden=nn.logistic, act_out=nn.softMax1):
    """afcnnMnist2 (MnistFcnnRanked2.hs:52-66)."""
    hidden, bias, hidden2, bias2, readout, biasr = params
    r, q = T.dtype(hidden), T.dtype(hidden2)
    hiddenLayer1 = T.binary("add", T.smatvecmul(hidden, datum), bias)
    nonlinearLayer1 = act_hidden(T, hiddenLayer1)
    hiddenLayer2 = T.binary("add", T.scast(T.smatvecmul(hidden2, T.scast(nonlinearLayer1, q)), r), bias2)
    nonlinearLayer2 = act_hidden(T, hiddenLayer2)
    outputLayer = T.binary("add", T.smatvecmul(readout, nonlinearLayer2), biasr)
    return act_out(T, outputLayer)


def afcnnMnistLoss2(T, datum, target, params):
    """afcnnMnistLoss2 (MnistFcnnRanked2.hs:70-76)."""
    return nn.lossCrossEntropyV(T, target, afcnnMnist2(T, datum, params))

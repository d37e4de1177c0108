"""Gradient of MnistRnnShaped2 at the bench shape: tanh_affine node vs the unfused program (scratch tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ad, mnist_rnn
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
rng = np.random.Generator(np.random.Philox(42))
width, batch = 512, 1024
params = [dev.sconcrete(p) for p in mnist_rnn.init_params(width, rng_range=0.05)]
glyphs = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28)))
labels = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])
res = {}
for fused in (False, True):
    v, g = ad.grad(dev, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=fused), params)
    res[fused] = (float(dev.kfromS(v)), [dev.to_numpy(x) for x in g])
print("loss", res[False][0], res[True][0])
for a, b in zip(res[False][1], res[True][1]):
    print(a.shape, "max abs", np.abs(a).max(), "max diff", np.abs(a - b).max())

"""MnistRnnShaped2 (width 512, batch 1024): loss after k eager Adam steps, to compare builds / switches
(HB_AD_NO_MATMUL_ACC=1: weight cotangents as product + separate add).  Scratch tool."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ad, mnist_rnn
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
rng = np.random.Generator(np.random.Philox(42))
width, batch = 512, 1024
params = [dev.sconcrete(p) for p in mnist_rnn.init_params(width, rng_range=0.05)]
moms = [(dev.sconcrete(np.zeros(dev.shape(p))), dev.sconcrete(np.zeros(dev.shape(p)))) for p in params]
glyphs = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28)))
labels = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])
for t in range(1, int(sys.argv[1]) + 1 if len(sys.argv) > 1 else 7):
    val, grads = ad.grad(dev, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=True), params)
    print(t, repr(float(dev.kfromS(val))), repr(float(np.abs(dev.to_numpy(grads[1])).sum())), flush=True)
    for p, (mm, vv), g in zip(params, moms, grads):
        dev.adam_step(p, mm, vv, g, t)

"""One smatmul2 shape, a few calls (for ncu; scratch tool): python tools/gemm_one.py M K N"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
m, k, n = (int(x) for x in sys.argv[1:4])
rng = np.random.Generator(np.random.Philox(1))
A, B = dev.sconcrete(rng.uniform(-.5, .5, (m, k))), dev.sconcrete(rng.uniform(-.5, .5, (k, n)))
for _ in range(6):
    dev.smatmul2(A, B)
dev.sync()

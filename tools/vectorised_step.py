"""Times one eager MnistCnnShaped2 gradient step at the three fusion levels (scratch tool)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ad, mnist_cnn
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
rng = np.random.Generator(np.random.Philox(42))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
params = [dev.sconcrete(p) for p in mnist_cnn.init_params(4, 4, 16, 64)]
gl = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28)))
lb = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])
for fused in (True, "nodes", False):
    f = lambda: ad.grad(dev, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, gl, lb, list(ps), fused=fused), params)
    v, g = f(); dev.sync()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        v, g = f()
    loss = float(dev.kfromS(v))
    dt = (time.perf_counter() - t0) / reps
    print(f"batch {batch} fused={fused!s:6} {dt*1e3:9.2f} ms/step  {batch/dt:12.0f} samples/s  loss {loss:.6f}", flush=True)
    del v, g

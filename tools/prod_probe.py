"""LongProdBench kernel timing (prefix / suffix product scan of csrc/scan.cu) at n = 1e6, 1e7, 1e8."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ffi
from horde_ad_b200.device import DeviceBackend

dev = DeviceBackend(0)
rng = np.random.default_rng(0)
ms = C.c_float()
for n in (10 ** 6, 10 ** 7, 10 ** 8):
    x = dev.sconcrete(np.exp(rng.uniform(-1e-3, 1e-3, n)))
    for _ in range(3):
        dev.prod_fold_grad(x, 1.0)
    dev.sync()
    reps = 10
    ffi.call("hb_event_record", 2)
    for _ in range(reps):
        dev.prod_fold_grad(x, 1.0)
    ffi.call("hb_event_record", 3)
    ffi.call("hb_event_elapsed_ms", 2, 3, C.byref(ms))
    t = ms.value / reps
    pr, g = dev.prod_fold_grad(x, 1.0)
    p = float(dev.kfromS(pr))
    err = float(np.max(np.abs(dev.to_numpy(dev.binary("mul", g, x)) / p - 1.0)))
    print(f"n={n}: {t:.4f} ms  {16.0 * n / t / 1e6:.0f} GB/s of 16n  ({16.0 * n / t / 1e6 / 6558.4:.3f} of HBM)  err {err:.2e}", flush=True)

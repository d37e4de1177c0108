"""Times Float smatmul2 / conv on the tcgen05 engine (used under ncu for csrc/contract_tf32.cu)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ffi
from horde_ad_b200.device import DeviceBackend

dev = DeviceBackend(0)
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
a = dev.sconcrete(rng.uniform(-1, 1, (n, n)).astype(np.float32))
b = dev.sconcrete(rng.uniform(-1, 1, (n, n)).astype(np.float32))
ms = C.c_float()


def timed(f, tag, flops):
    for _ in range(2):
        f()
    dev.sync()
    ffi.call("hb_event_record", 2)
    for _ in range(reps):
        f()
    ffi.call("hb_event_record", 3)
    ffi.call("hb_event_elapsed_ms", 2, 3, C.byref(ms))
    t = ms.value / reps
    print(f"{tag}: {t:.4f} ms  {flops / t / 1e9:.1f} TFLOP/s", flush=True)


timed(lambda: dev.smatmul2(a, b), f"smatmul2 f32 {n}^3", 2.0 * n ** 3)
at = dev.stranspose([1, 0], a)
timed(lambda: dev.smatmul2(at, b), f"smatmul2 f32 A^T {n}^3", 2.0 * n ** 3)
K = dev.sconcrete(rng.uniform(-.5, .5, (64, 64, 3, 3)).astype(np.float32))
A = dev.sconcrete(rng.uniform(-.5, .5, (256, 64, 28, 28)).astype(np.float32))
dB = dev.sconcrete(rng.uniform(-.5, .5, (256, 64, 28, 28)).astype(np.float32))
fl = 2.0 * 256 * 64 * 28 * 28 * 64 * 9
timed(lambda: dev.conv2dPreserving(K, A), "conv fwd f32", fl)
timed(lambda: dev.conv2d_dkrn(A, dB, 3, 3), "conv dkrn f32", fl)
timed(lambda: dev.conv2d_dinp(K, dB), "conv dinp f32", fl)

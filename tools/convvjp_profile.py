"""ConvVjpBench shape (K[64,64,3,3], A,dB[256,64,28,28] f64): time the three passes, or run each once for ncu.
Usage: python tools/convvjp_profile.py [reps]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ffi
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.Generator(np.random.Philox(42))
K = dev.sconcrete(rng.uniform(-.5, .5, (64, 64, 3, 3)))
A = dev.sconcrete(rng.uniform(-.5, .5, (256, 64, 28, 28)))
dB = dev.sconcrete(rng.uniform(-.5, .5, (256, 64, 28, 28)))
fl = 2 * 256 * 64 * 28 * 28 * 64 * 9
ms = C.c_float()
for name, f in (("fwd", lambda: dev.conv2dPreserving(K, A)), ("dkrn", lambda: dev.conv2d_dkrn(A, dB, 3, 3)),
                ("dinp", lambda: dev.conv2d_dinp(K, dB))):
    for _ in range(2):
        f()
    dev.sync()
    ffi.call("hb_event_record", 2)
    for _ in range(reps):
        f()
    ffi.call("hb_event_record", 3)
    ffi.call("hb_event_elapsed_ms", 2, 3, C.byref(ms))
    t = ms.value / reps * 1e-3
    print(f"{name}: {t*1e3:.4f} ms  {fl/t/1e12:.2f} TFLOP/s", flush=True)

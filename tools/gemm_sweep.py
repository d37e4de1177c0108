"""smatmul2 timing over the GEMM shapes of the BASELINE configs, all four operand layouts (scratch tool).
HB_GEMM8=0/1 forces the four-warp / eight-warp kernel; unset = the library's own choice."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ffi
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
rng = np.random.Generator(np.random.Philox(1))
shapes = [tuple(int(x) for x in a.split(',')) for a in sys.argv[1:]] or [(512, 512, 1024), (512, 1024, 512), (512, 28, 1024), (64, 784, 4096), (64, 4096, 784), (784, 64, 4096),
          (10, 64, 4096), (128, 3136, 4096), (512, 512, 8192), (1024, 1024, 1024), (300, 784, 1), (2048, 2048, 2048)]
def timed(f, reps=20):
    for _ in range(3): f()
    ffi.call("hb_event_record", 0)
    for _ in range(reps): f()
    ffi.call("hb_event_record", 1)
    ms = C.c_float()
    ffi.call("hb_event_elapsed_ms", 0, 1, C.byref(ms))
    return ms.value / reps
for (m, k, n) in shapes:
    a = rng.uniform(-.5, .5, (m, k)); b = rng.uniform(-.5, .5, (k, n))
    ref = a @ b
    row = []
    for ta in (0, 1):
        for tb in (0, 1):
            A = dev.stranspose([1, 0], dev.sconcrete(np.ascontiguousarray(a.T))) if ta else dev.sconcrete(a)
            B = dev.stranspose([1, 0], dev.sconcrete(np.ascontiguousarray(b.T))) if tb else dev.sconcrete(b)
            got = dev.to_numpy(dev.smatmul2(A, B))
            err = float(np.max(np.abs(got - ref)) / max(1e-300, np.max(np.abs(ref))))
            ms = timed(lambda: dev.smatmul2(A, B))
            row.append(f"{ms*1e3:7.1f}us {2.0*m*k*n/(ms*1e-3)/1e12:5.1f}TF" + ("" if err < 1e-13 else f" ERR{err:.1e}"))
    print(f"[{m},{k}]x[{k},{n}]  NN {row[0]} | NT {row[1]} | TN {row[2]} | TT {row[3]}", flush=True)

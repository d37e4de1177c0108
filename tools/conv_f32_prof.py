"""Per-launch times of the Float conv paths (channels-last staging + tcgen05 engine)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ffi
from horde_ad_b200.device import DeviceBackend

dev = DeviceBackend(0)
rng = np.random.default_rng(0)
K = dev.sconcrete(rng.uniform(-.5, .5, (64, 64, 3, 3)).astype(np.float32))
A = dev.sconcrete(rng.uniform(-.5, .5, (256, 64, 28, 28)).astype(np.float32))
dB = dev.sconcrete(rng.uniform(-.5, .5, (256, 64, 28, 28)).astype(np.float32))
for name, f in (("fwd", lambda: dev.conv2dPreserving(K, A)), ("dkrn", lambda: dev.conv2d_dkrn(A, dB, 3, 3)),
                ("dinp", lambda: dev.conv2d_dinp(K, dB))):
    for _ in range(2):
        f()
    dev.sync()
    ffi.call("hb_prof_begin")
    for _ in range(3):
        f()
    ffi.call("hb_prof_end")
    buf = C.create_string_buffer(1 << 16)
    ffi.call("hb_prof_report", buf, len(buf))
    print("==", name)
    for line in buf.value.decode().splitlines():
        tag, site, cnt, tot = line.rsplit(" ", 3)
        print(f"  {float(tot) / 3:8.4f} ms  n={int(cnt) // 3}  {site}  {tag[:60]}")

"""Per-kernel timing of hb_prod_fold_grad at n = 1e8 (scratch tool)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ffi
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10**8
x = dev.sconcrete(np.exp(np.random.default_rng(0).uniform(-1e-3, 1e-3, n)))
for _ in range(3): dev.prod_fold_grad(x, 1.0)
dev.sync()
ffi.call("hb_prof_begin")
for _ in range(5): dev.prod_fold_grad(x, 1.0)
ffi.call("hb_prof_end")
buf = C.create_string_buffer(1 << 16)
ffi.call("hb_prof_report", buf, len(buf))
for line in buf.value.decode().splitlines():
    tag, site, cnt, ms = line.rsplit(" ", 3)
    print(f"{float(ms)/int(cnt):8.4f} ms  n={cnt} {tag} @ {site}")

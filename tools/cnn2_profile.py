"""Per-launch-site breakdown of one eager step of the secondary CNN config (3x3 kernels, 64 channels; scratch tool)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ffi
from horde_ad_b200.device import DeviceBackend
from horde_ad_b200.train import CnnTrainer
dev = DeviceBackend(0)
rng = np.random.Generator(np.random.Philox(42))
B = 4096
tr = CnnTrainer(dev, B, kh=2, kw=2, c_out=64, n_hidden=128, use_graph=False)
gl = np.ascontiguousarray(rng.uniform(0, 1, (B, 28, 28)))
lb = np.ascontiguousarray(np.eye(10)[rng.integers(0, 10, B)])
tr.upload(gl.ctypes.data_as(C.c_void_p), lb.ctypes.data_as(C.c_void_p))
for _ in range(2):
    tr.step_enqueue()
dev.sync()
ffi.call("hb_prof_begin"); tr.step_enqueue(); ffi.call("hb_prof_end")
buf = C.create_string_buffer(1 << 16)
ffi.call("hb_prof_report", buf, len(buf))
tot = 0
rows = []
for line in buf.value.decode().splitlines():
    tag, site, cnt, ms = line.rsplit(" ", 3)
    rows.append((float(ms), int(cnt), tag, site)); tot += float(ms)
print("total ms", tot)
for ms, cnt, tag, site in rows[:16]:
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}% n={cnt:4d} {tag} @ {site}")

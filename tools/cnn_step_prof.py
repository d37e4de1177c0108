"""Per-launch-site times (hb_prof_*) of one MnistCnnShaped2 training step replayed WITHOUT the CUDA graph, at a given
per-GPU batch (default 512: the strong-scaling shard at 8 GPUs)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ffi
from horde_ad_b200.device import DeviceBackend
from horde_ad_b200.train import CnnTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = DeviceBackend(0)
rng = np.random.default_rng(0)
tr = CnnTrainer(dev, B, use_graph=False)
gl = np.ascontiguousarray(rng.uniform(0, 1, (B, 28, 28)))
lb = np.ascontiguousarray(np.eye(10)[rng.integers(0, 10, B)])
tr.upload(gl.ctypes.data_as(C.c_void_p), lb.ctypes.data_as(C.c_void_p))
for _ in range(3):
    tr.step_enqueue()
dev.sync()
ffi.call("hb_prof_begin")
tr.step_enqueue()
ffi.call("hb_prof_end")
buf = C.create_string_buffer(1 << 16)
ffi.call("hb_prof_report", buf, len(buf))
rows = []
for line in buf.value.decode().splitlines():
    tag, site, cnt, tot = line.rsplit(" ", 3)
    rows.append((float(tot), int(cnt), site, tag))
print("batch", B, "total ms", sum(r[0] for r in rows), "launches", sum(r[1] for r in rows))
for tot, cnt, site, tag in sorted(rows, reverse=True)[:30]:
    print(f"{tot * 1e3:8.1f} us n={cnt:3d}  {site}  {tag[:80]}")

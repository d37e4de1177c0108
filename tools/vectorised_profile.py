"""Per-launch-site breakdown of one eager VECTORISED MnistCnnShaped2 gradient step (scratch tool)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ad, ffi, mnist_cnn
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
rng = np.random.Generator(np.random.Philox(42))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
params = [dev.sconcrete(p) for p in mnist_cnn.init_params(4, 4, 16, 64)]
gl = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28)))
lb = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])
f = lambda: ad.grad(dev, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, gl, lb, list(ps), fused=False), params)
f(); dev.sync()
ffi.call("hb_prof_begin"); f(); ffi.call("hb_prof_end")
buf = C.create_string_buffer(1 << 16)
ffi.call("hb_prof_report", buf, len(buf))
rows = []
tot = 0
for line in buf.value.decode().splitlines():
    tag, site, cnt, ms = line.rsplit(" ", 3)
    rows.append((float(ms), int(cnt), tag, site)); tot += float(ms)
print("total ms", round(tot, 2))
for ms, cnt, tag, site in rows[:22]:
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}% n={cnt:3d} {tag[:95]} @ {site}")

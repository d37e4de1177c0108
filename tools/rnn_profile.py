"""Per-launch-site breakdown of one eager MnistRnnShaped2 gradient step (scratch tool)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ad, ffi, mnist_rnn
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(0)
rng = np.random.Generator(np.random.Philox(42))
width, batch = 512, 1024
params = [dev.sconcrete(p) for p in mnist_rnn.init_params(width, rng_range=0.05)]
glyphs = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28)))
labels = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])
fused = {'fused': True, 'pair': 'pair'}.get(sys.argv[1] if len(sys.argv) > 1 else '', False)
f = lambda: ad.grad(dev, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=fused), params)
f(); f(); dev.sync()
ffi.call("hb_prof_begin"); f(); ffi.call("hb_prof_end")
buf = C.create_string_buffer(1 << 16)
ffi.call("hb_prof_report", buf, len(buf))
tot = 0
rows = []
for line in buf.value.decode().splitlines():
    tag, site, cnt, ms = line.rsplit(" ", 3)
    rows.append((float(ms), int(cnt), tag, site)); tot += float(ms)
print("total ms", tot)
for ms, cnt, tag, site in rows[:25]:
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}% n={cnt:4d} {tag} @ {site}")

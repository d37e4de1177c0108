"""Per-launch-site times of one MnistRnnShaped2 gradient step (width 512, batch 1024, fused pair nodes)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horde_ad_b200 import ad, ffi, mnist_rnn
from horde_ad_b200.device import DeviceBackend

dev = DeviceBackend(0)
rng = np.random.default_rng(0)
width, batch = 512, 1024
if os.environ.get("RNN_FLAT"):    # the bench's layout: leaves are views of one flat parameter array, Adam inside the step
    from horde_ad_b200.train import BucketLayout, exchange_scale
    host = mnist_rnn.init_params(width, rng_range=float(os.environ.get("RNN_RANGE", "0.05")))
    skew = int(os.environ.get("RNN_SKEW", "0"))
    if skew:      # dummy leaves between the real ones: no two weight matrices at a power-of-two distance
        host = [x for p in host for x in (p, np.zeros(skew))][:-1]
    lay = BucketLayout([p.shape for p in host])
    flat = lay.flatten(host, 0.0)[:lay.n_params]
    fp, fm, fv = dev.sconcrete(flat), dev.sconcrete(np.zeros_like(flat)), dev.sconcrete(np.zeros_like(flat))
    t_dev = dev.sconcrete(np.zeros((), np.int64))
    leaves = [dev.sreshape(list(sh), dev.sslice(off, n, fp)) for sh, n, off in zip(lay.shapes, lay.sizes, lay.offsets)]
    params = leaves[::2] if skew else leaves
    pad = dev.sconcrete(np.zeros(skew)) if skew else None
else:
    params = [dev.sconcrete(p) for p in mnist_rnn.init_params(width, rng_range=float(os.environ.get("RNN_RANGE", "0.05")))]
glyphs = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28)))
labels = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])
grad = lambda: ad.grad(dev, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=os.environ.get("RNN_FUSION", "wave")), params)


def f():
    val, grads = grad()
    if os.environ.get("RNN_FLAT"):
        gl = [x for g in grads for x in (g, pad)][:-1] if skew else list(grads)
        bucket = dev.flatten_concat(gl + [val])
        if not os.environ.get("RNN_NO_ADAM"):
            ffi.call("hb_allreduce_adam_step", bucket.h, lay.n_params, fp.h, fm.h, fv.h, t_dev.h, 1e-3, 0.9, 0.999, 1e-7, float(os.environ.get("RNN_LR_SCALE", "1.0")))
        return val, bucket
    return val, grads


for _ in range(2):
    f()
dev.sync()
ffi.call("hb_prof_begin")
f()
ffi.call("hb_prof_end")
buf = C.create_string_buffer(1 << 16)
ffi.call("hb_prof_report", buf, len(buf))
rows = []
for line in buf.value.decode().splitlines():
    tag, site, cnt, tot = line.rsplit(" ", 3)
    rows.append((float(tot), int(cnt), site, tag))
print("total ms", sum(r[0] for r in rows), "launches", sum(r[1] for r in rows))
for tot, cnt, site, tag in sorted(rows, reverse=True)[:24]:
    print(f"{tot:8.3f} ms n={cnt:4d} {tot / cnt * 1e3:7.1f} us  {site}  {tag[:70]}")

# the same step as one CUDA graph (what bench_configs.py times)
ffi.call("hb_graph_begin")
try:
    val, grads = f()
finally:
    g = C.c_void_p()
    ffi.call("hb_graph_end", C.byref(g))
for _ in range(3):
    ffi.call("hb_graph_launch", g)
dev.sync()
ms = C.c_float()
ffi.call("hb_event_record", 2)
for _ in range(10):
    ffi.call("hb_graph_launch", g)
ffi.call("hb_event_record", 3)
ffi.call("hb_event_elapsed_ms", 2, 3, C.byref(ms))
fl = 28 * 3 * 2.0 * batch * (28 * width + 3 * width * width)
print(f"graph step {ms.value / 10:.3f} ms  {fl / (ms.value / 10) / 1e9:.2f} TFLOP/s")

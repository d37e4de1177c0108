"""CPU-side tests of the device contraction pass: the library is initialised WITHOUT a device
(hb_init_trace_only), the unchanged vectorised MnistCnnShaped2 gradient program is recorded through the C ABI
and the rewriter's output is inspected.  Each case runs in a subprocess because trace-only mode is process-wide."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code: str) -> dict:
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


PRELUDE = """
import json, numpy as np
from horde_ad_b200 import ad, mnist_cnn, mnist_rnn, nn, ffi
from horde_ad_b200.device import DeviceBackend
dev = DeviceBackend(trace_only=True)
rng = np.random.Generator(np.random.Philox(42))
"""


def test_vectorised_cnn_gradient_is_rewritten_to_whole_layer_nodes():
    r = _run(PRELUDE + textwrap.dedent("""
    batch, hyper = 16, (4, 4, 16, 64)
    params = [dev.sconcrete(p) for p in mnist_cnn.init_params(*hyper)]
    gl = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28))); lb = dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])
    def build():
        loss, grads = ad.grad(dev, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, gl, lb, list(ps), fused=False), params)
        return dev.flatten_concat(list(grads) + [loss])
    p = dev.record(build)
    q = dev.record(build, rewrite=False)
    print(json.dumps({"stats": p.stats, "raw": q.stats, "dump": p.dump()}))
    """))
    rec, live = r["stats"]
    assert r["raw"][0] == rec >= 50 and 45 <= r["raw"][1] <= rec and live <= 22      # without the pass only dead nodes go
    body, log = r["dump"].split("--\n")
    assert body.count("conv_relu_pool_fwd") == 2 and body.count("conv_relu_pool_bwd") == 2
    assert "scatter_add" not in body and "dot_inner" not in body and "sum_leading f64[16,16,5,5]" not in body
    # the pass says what it recognised, with the geometry it derived from the index maps
    assert "conv2d_preserving  N=16 Ci=1 Co=16 H=28 W=28 KH=5 KW=5" in log
    assert "conv2d_dinp  N=16 Ci=16 Co=16 H=14 W=14 KH=5 KW=5" in log
    assert "conv2d_dkrn  N=16 Ci=16 Co=16 H=14 W=14 KH=5 KW=5" in log
    assert log.count("relu_pool_fwd") >= 2 and log.count("relu_pool_bwd") >= 2


def test_recording_needs_no_device_and_refuses_to_run():
    r = _run(PRELUDE + textwrap.dedent("""
    x = dev.sconcrete(np.arange(12.0).reshape(3, 4))
    assert dev.to_numpy(x).tolist() == np.arange(12.0).reshape(3, 4).tolist()     # constants live on the host
    p = dev.record(lambda: dev.ssum(dev.binary("mul", x, x)), rewrite=False)
    errs = []
    try:
        p.run()
    except ffi.HordeError as e:
        errs.append(str(e))
    try:
        dev.binary("mul", x, x)
    except ffi.HordeError as e:
        errs.append(str(e))
    print(json.dumps({"stats": p.stats, "errs": errs}))
    """))
    assert r["stats"] == [2, 2]
    assert "no device" in r["errs"][0] and "no device" in r["errs"][1]


def test_rnn_and_fcnn_programs_record_and_survive_the_pass_unchanged():
    r = _run(PRELUDE + textwrap.dedent("""
    from horde_ad_b200 import mnist_fcnn
    params = [dev.sconcrete(p) for p in mnist_rnn.init_params(16)]
    gl = dev.sconcrete(rng.uniform(0, 1, (4, 28, 28))); lb = dev.sconcrete(np.eye(10)[rng.integers(0, 10, 4)])
    def build():
        loss, grads = ad.grad(dev, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, gl, lb, list(ps), fused=False), params)
        return dev.flatten_concat(list(grads) + [loss])
    p = dev.record(build)
    print(json.dumps({"stats": p.stats, "log": p.dump().split("--")[1]}))
    """))
    rec, live = r["stats"]
    assert rec > 300 and live <= rec and "conv2d" not in r["log"]

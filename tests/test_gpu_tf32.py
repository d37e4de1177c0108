"""The Float contraction engine on the 5th-generation tensor cores (csrc/contract_tf32.cu: tcgen05.mma kind::tf32,
three-term split, TMEM accumulators) against the oracle, through the C ABI.

Bar: the reference tests Float convolutions and products at 1e-5 (TestConvQuickCheck.hs:430-464); here the bound is
the normwise one, |got - ref| <= 1e-5 * sum |a||b| per entry, AND the summation-order-free bound K eps sum|a||b| of
tests/test_gpu_parity.close_dot wherever K is large enough for it to be the looser of the two.
"""
import ctypes as C

import numpy as np
import pytest

from horde_ad_b200 import ffi

pytestmark = pytest.mark.gpu

EPS32 = 2.0 ** -23


def launched_sites(fn):
    """Runs fn() under the per-launch profiler and returns the launch-site names it used."""
    ffi.call("hb_prof_begin")
    out = fn()
    ffi.call("hb_prof_end")
    buf = C.create_string_buffer(1 << 16)
    ffi.call("hb_prof_report", buf, len(buf))
    sites = [line.rsplit(" ", 3)[1] for line in buf.value.decode().splitlines()]
    return out, sites


def check(got, ref64, absref64, K):
    assert got.dtype == np.float32 and got.shape == ref64.shape
    err = np.abs(got.astype(np.float64) - ref64)
    # (1) the reference's own Float bar, normwise
    bound1 = 1e-5 * absref64 + np.finfo(np.float32).tiny
    # (2) f32-level: three-term TF32 split error (<= 2^-21 per product) + f32 accumulation of K terms
    bound2 = (K * EPS32 + 2.0 ** -20) * absref64 + np.finfo(np.float32).tiny
    assert float(np.max(err / bound1)) <= 1.0, f"normwise 1e-5 bound exceeded: {float(np.max(err / bound1)):.3g}"
    assert float(np.max(err / bound2)) <= 1.0, f"f32-level bound exceeded: {float(np.max(err / bound2)):.3g}"
    return float(np.max(err / (absref64 + 1e-300)))


@pytest.mark.parametrize("shape", [(1024, 1024, 1024), (300, 517, 129), (128, 8, 4096), (64, 4099, 784),
                                   (2048, 64, 32), (33, 2000, 1000), (576, 20000, 64)])
def test_f32_matmul_runs_on_tcgen05_and_meets_1e_5(dev, shape):
    m, k, n = shape
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    absref = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    da, db = dev.sconcrete(a), dev.sconcrete(b)
    got, sites = launched_sites(lambda: dev.to_numpy(dev.smatmul2(da, db)))
    assert any(s.startswith("hb_contract_tf32") for s in sites), sites
    worst = check(got, ref, absref, k)
    assert worst < 3e-6, worst          # far inside 1e-5: this is not plain TF32 (which sits near 1e-3)
    # transposed / sliced / reversed views of both operands reach the same kernel through the offset tables
    at = dev.stranspose([1, 0], dev.sconcrete(np.ascontiguousarray(a.T)))
    bt = dev.stranspose([1, 0], dev.sconcrete(np.ascontiguousarray(b.T)))
    check(dev.to_numpy(dev.smatmul2(at, bt)), ref, absref, k)
    check(dev.to_numpy(dev.smatmul2(at, db)), ref, absref, k)
    if m > 4:
        check(dev.to_numpy(dev.smatmul2(dev.sslice(2, m - 3, da), db)), ref[2:m - 1], absref[2:m - 1], k)
        check(dev.to_numpy(dev.smatmul2(dev.sreverse(da), db)), ref[::-1], absref[::-1], k)
    # deterministic: the same bits on a second run (split-K partials are added in split order)
    again = dev.to_numpy(dev.smatmul2(da, db))
    assert np.array_equal(got.view(np.uint32), again.view(np.uint32))


def test_f32_split_is_exact_for_tf32_representable_and_handles_non_finite(dev):
    rng = np.random.default_rng(6)
    # small integers: every product and partial sum is exact in f32, so the result must be bit-exact
    a = rng.integers(-8, 9, (256, 96)).astype(np.float32)
    b = rng.integers(-8, 9, (96, 160)).astype(np.float32)
    got = dev.to_numpy(dev.smatmul2(dev.sconcrete(a), dev.sconcrete(b)))
    assert np.array_equal(got, (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))
    # values that need all 24 mantissa bits (plain TF32 would lose 13 of them)
    a = (1.0 + rng.integers(0, 2 ** 23, (256, 64)) * 2.0 ** -23).astype(np.float32)
    b = np.eye(64, 128, dtype=np.float32)
    got = dev.to_numpy(dev.smatmul2(dev.sconcrete(a), dev.sconcrete(b)))
    assert np.max(np.abs(got[:, :64] - a) / a) <= 2.0 ** -21
    # an Inf / NaN in one operand element reaches exactly the outputs that use it (as Inf or NaN: Inf * lo_b of
    # either sign meets Inf * hi_b in the accumulator -- the one observable difference to a plain f32 product)
    a = rng.uniform(-1, 1, (256, 64)).astype(np.float32)
    b = rng.uniform(0.5, 1, (64, 128)).astype(np.float32)
    a[7, 3] = np.inf
    a[100, 60] = np.nan
    got = dev.to_numpy(dev.smatmul2(dev.sconcrete(a), dev.sconcrete(b)))
    assert not np.any(np.isfinite(got[7])) and np.all(np.isnan(got[100]))
    mask = np.ones(256, bool)
    mask[[7, 100]] = False
    assert np.all(np.isfinite(got[mask]))


@pytest.mark.parametrize("shape", [(256, 64, 64, 28, 28, 3, 3),      # ConvVjpBench (BASELINE config 3) in Float
                                   (64, 16, 16, 14, 14, 5, 5),       # the CNN's second layer
                                   (5, 3, 7, 12, 9, 2, 3), (32, 1, 16, 28, 28, 5, 5)])
def test_f32_conv_and_vjps_on_tcgen05(dev, oracle, shape):
    N, Ci, Co, H, W, KH, KW = shape
    rng = np.random.default_rng(16)
    K = rng.uniform(-.5, .5, (Co, Ci, KH, KW)).astype(np.float32)
    A = rng.uniform(-.5, .5, (N, Ci, H, W)).astype(np.float32)
    dB = rng.uniform(-.5, .5, (N, Co, H, W)).astype(np.float32)
    K64, A64, dB64 = K.astype(np.float64), A.astype(np.float64), dB.astype(np.float64)
    dK_, dA_, ddB = dev.sconcrete(K), dev.sconcrete(A), dev.sconcrete(dB)
    big = N * H * W * Co * Ci * KH * KW >= 1 << 18
    got, sites = launched_sites(lambda: dev.to_numpy(dev.conv2dPreserving(dK_, dA_)))
    if big:
        assert any(s.startswith("hb_contract_tf32") for s in sites), sites
    check(got, oracle.conv2dPreserving(K64, A64), oracle.conv2dPreserving(np.abs(K64), np.abs(A64)), Ci * KH * KW)
    got, sites = launched_sites(lambda: dev.to_numpy(dev.conv2d_dkrn(dA_, ddB, KH, KW)))
    if big and max(Co, Ci * KH * KW) >= 32:       # a 16 x 25 result has no tile's worth of rows on either side
        assert any(s.startswith("hb_contract_tf32") for s in sites), sites
    check(got, oracle.conv2d_dkrn(A64, dB64, KH, KW), oracle.conv2d_dkrn(np.abs(A64), np.abs(dB64), KH, KW), N * H * W)
    got, sites = launched_sites(lambda: dev.to_numpy(dev.conv2d_dinp(dK_, ddB)))
    if big and Ci >= 16:                          # a one-channel result is a matrix-vector product (reduction path)
        assert any(s.startswith("hb_contract_tf32") for s in sites), sites
    check(got, oracle.conv2d_dinp(K64, dB64), oracle.conv2d_dinp(np.abs(K64), np.abs(dB64)), Co * KH * KW)


def test_f32_engine_agrees_with_the_ffma_engine(dev):
    """HB_TF32=0 keeps the FFMA register-tile kernel: both are within the bound of the exact product, hence of each
    other (checked in a subprocess because the switch is read once)."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "from horde_ad_b200.device import DeviceBackend\n"
        "dev = DeviceBackend(0)\n"
        "rng = np.random.default_rng(1)\n"
        "a = rng.uniform(-1, 1, (384, 300)).astype(np.float32); b = rng.uniform(-1, 1, (300, 200)).astype(np.float32)\n"
        "np.save(sys.argv[1], dev.to_numpy(dev.smatmul2(dev.sconcrete(a), dev.sconcrete(b))))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    outs = []
    for mode in ("0", "1"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            env = dict(os.environ, HB_TF32=mode)
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, timeout=300)
            outs.append(np.load(f.name))
    rng = np.random.default_rng(1)
    a = rng.uniform(-1, 1, (384, 300)).astype(np.float32)
    b = rng.uniform(-1, 1, (300, 200)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    absref = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    for o in outs:
        check(o, ref, absref, 300)


def test_f32_cnn_gradient_through_the_engine(dev, oracle):
    """MnistCnnShaped2 in Float: loss and the 8 gradient leaves of the whole-layer program (Float conv nodes ->
    channels-last views -> the tcgen05 engine for the 16 -> 16 channel layer, dense products) against the oracle's
    UNFUSED program evaluated in Double on the same Float inputs.  Normwise per leaf: the whole model in f32 (every
    elementwise op rounds to 24 bits on both sides of the reference's own Float runs) stays inside 2e-5."""
    from horde_ad_b200 import ad, mnist_cnn
    rng = np.random.Generator(np.random.Philox(7))
    batch, kh, kw, c_out, n_hidden = 32, 2, 2, 16, 8
    params = [p.astype(np.float32) for p in mnist_cnn.init_params(kh, kw, c_out, n_hidden)]
    glyphs = rng.uniform(0, 1, (batch, 28, 28)).astype(np.float32)
    labels = np.eye(10, dtype=np.float32)[rng.integers(0, 10, batch)]

    def run(B, fused, dt):
        ins = [B.sconcrete(p.astype(dt)) for p in params]
        g, l = B.sconcrete(glyphs.astype(dt)), B.sconcrete(labels.astype(dt))
        val, grads = ad.grad(B, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, g, l, list(ps), fused=fused), ins)
        return float(B.kfromS(val)), [B.to_numpy(x) for x in grads]
    (dl, dg), sites = launched_sites(lambda: run(dev, "nodes", np.float32))
    assert any(s.startswith("hb_contract_tf32") for s in sites), sites
    ol, og = run(oracle, False, np.float64)
    assert abs(dl - ol) <= 2e-5 * abs(ol), (dl, ol)
    for a, b in zip(dg, og):
        assert a.dtype == np.float32
        assert float(np.max(np.abs(a.astype(np.float64) - b))) <= 2e-5 * float(np.max(np.abs(b))), (a.shape,)

"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): bit-exact for data movement, indexing and
integer work; 1e-12 relative for Double, 1e-5 for Float floating point.
Every test here needs a B200 (`-m gpu`); the library has no CPU fallback.
"""
import os

import numpy as np
import pytest

from horde_ad_b200 import ad, ffi, ix as ixm, nn, mnist_cnn
from tests import programs as P

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}
ALL_DTYPES = [np.float64, np.float32, np.int64, np.int32, np.int16, np.int8, np.bool_]


def rnd(rng, shape, dtype=np.float64, lo=-0.5, hi=0.5):
    dtype = np.dtype(dtype)
    if dtype == np.bool_:
        return rng.integers(0, 2, shape).astype(np.bool_)
    if np.issubdtype(dtype, np.integer):
        return rng.integers(-100, 100, shape).astype(dtype)
    return rng.uniform(lo, hi, shape).astype(dtype)


def close(dev_arr, ref, dtype=None, scale=1.0):
    dtype = np.dtype(dtype or ref.dtype)
    got = dev_arr if isinstance(dev_arr, np.ndarray) else None
    assert got is not None
    assert got.shape == tuple(ref.shape), (got.shape, ref.shape)
    if np.issubdtype(dtype, np.floating):
        tol = RTOL[dtype] * scale
        np.testing.assert_allclose(got, ref, rtol=tol, atol=tol * max(1.0, float(np.max(np.abs(ref))) if ref.size else 1.0))
    else:
        np.testing.assert_array_equal(got, ref)


EPS = {np.dtype(np.float64): 2.0 ** -52, np.dtype(np.float32): 2.0 ** -23}


def close_dot(got, ref, absref, K, dtype):
    """Contractions (inner products of length K): the componentwise forward-error bound that holds for ANY summation
    order on both sides, |got - ref| <= K * eps * sum |a_i| |b_i|  (each side is within gamma_K = K u of the exact
    value, eps = 2 u; Higham, Accuracy and Stability of Numerical Algorithms, section 3.1), with `absref` the same
    contraction evaluated on the operands' absolute values.  No ad-hoc slack factor: where cancellation makes |ref|
    small the bound does not grow with 1 / |ref|, and it is far below BASELINE's 1e-12 / 1e-5 relative bar for the
    well-conditioned entries."""
    dtype = np.dtype(dtype)
    assert got.shape == tuple(ref.shape), (got.shape, ref.shape)
    assert got.dtype == dtype, (got.dtype, dtype)
    bound = max(1, K) * EPS[dtype] * np.asarray(absref, np.float64) + np.finfo(dtype).tiny
    err = np.abs(got.astype(np.float64) - np.asarray(ref, np.float64))
    worst = float(np.max(err / bound)) if err.size else 0.0
    assert worst <= 1.0, f"forward error {worst:.3g} x the K*eps*sum|a||b| bound (K={K}, {dtype})"


def exact(got, ref):
    assert got.shape == tuple(ref.shape) and got.dtype == ref.dtype, (got.shape, ref.shape, got.dtype, ref.dtype)
    assert np.array_equal(got.view(np.uint8) if got.dtype != np.bool_ else got,
                          np.ascontiguousarray(ref).view(np.uint8) if ref.dtype != np.bool_ else ref), "not bit-exact"


# ------------------------------------------------------------------ plumbing
@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_roundtrip_and_views(dev, oracle, dtype):
    rng = np.random.default_rng(0)
    a = rnd(rng, (3, 4, 5), dtype)
    d = dev.sconcrete(a)
    exact(dev.to_numpy(d), a)
    for perm in ([1, 0], [2, 0, 1], [0, 2, 1]):
        exact(dev.to_numpy(dev.stranspose(perm, d)), np.ascontiguousarray(oracle.stranspose(perm, a)))
    exact(dev.to_numpy(dev.sreplicateN([2, 3], d)), np.ascontiguousarray(oracle.sreplicateN([2, 3], a)))
    exact(dev.to_numpy(dev.sslice(1, 2, d)), a[1:3])
    exact(dev.to_numpy(dev.sreverse(d)), np.ascontiguousarray(a[::-1]))
    exact(dev.to_numpy(dev.sreshape([4, 15], d)), a.reshape(4, 15))
    z = dev.sconcrete(a[1, 2, 3])                       # a 0-d array stays 0-d (TKS '[] r)
    assert z.shape == () and dev.to_numpy(z).reshape(1)[0] == a[1, 2, 3]
    exact(dev.to_numpy(dev.sreshape([5, 12], dev.stranspose([2, 1, 0], d))), np.transpose(a, (2, 1, 0)).reshape(5, 12))
    exact(dev.to_numpy(dev.sindex(d, [2, 1])), a[2, 1])
    exact(dev.to_numpy(dev.sindex(d, [3, 1])), np.zeros(5, dtype))      # OOB -> zeros
    exact(dev.to_numpy(dev.sindex(d, [-1])), np.zeros((4, 5), dtype))
    exact(dev.to_numpy(dev.sappend(d, dev.sreverse(d))), np.concatenate([a, a[::-1]]))
    exact(dev.to_numpy(dev.sfromVector([d, dev.sreverse(d)])), np.stack([a, a[::-1]]))
    exact(dev.to_numpy(dev.srepl([2, 3], a.reshape(-1)[0], dtype)), np.full((2, 3), a.reshape(-1)[0], dtype))
    exact(dev.to_numpy(dev.soneHot([2, 3], dev.sindex(d, [0, 0]), [1, 2])), oracle.soneHot([2, 3], a[0, 0], [1, 2]))
    exact(dev.to_numpy(dev.soneHot([2, 3], dev.sindex(d, [0, 0]), [2, 0])), np.zeros((2, 3, 5), dtype))
    assert dev.kfromS(dev.sindex(d, [1, 2, 3])) == a[1, 2, 3]


def test_flatten_concat(dev):
    rng = np.random.default_rng(5)
    leaves = [rng.uniform(-1, 1, sh) for sh in [(16, 1, 5, 5), (16,), (3, 7), (), (1000,)]]
    d = [dev.sconcrete(x) for x in leaves]
    exact(dev.to_numpy(dev.flatten_concat(d)), np.concatenate([x.reshape(-1) for x in leaves]))
    # a non-contiguous leaf takes the per-leaf copy path
    d[2] = dev.stranspose([1, 0], dev.sconcrete(leaves[2].T.copy()))
    exact(dev.to_numpy(dev.flatten_concat(d)), np.concatenate([x.reshape(-1) for x in leaves]))
    many = [dev.sconcrete(np.full(3, float(i))) for i in range(20)]
    exact(dev.to_numpy(dev.flatten_concat(many)), np.repeat(np.arange(20.0), 3))


def test_empty_and_errors(dev):
    e = dev.sconcrete(np.zeros((0, 3)))
    assert dev.to_numpy(e).shape == (0, 3)
    assert dev.to_numpy(dev.ssum(e)).tolist() == [0, 0, 0]
    assert dev.to_numpy(dev.binary("add", e, e)).shape == (0, 3)
    with pytest.raises(ffi.HordeError, match="empty vector"):
        dev.sfromVector([])
    with pytest.raises(ffi.HordeError):
        dev.binary("add", dev.sconcrete(np.zeros(3)), dev.sconcrete(np.zeros(4)))
    with pytest.raises(ffi.HordeError):
        dev.binary("add", dev.sconcrete(np.zeros(3)), dev.sconcrete(np.zeros(3, np.float32)))
    with pytest.raises(ffi.HordeError):
        dev.unary("exp", dev.sconcrete(np.zeros(3, np.int64)))
    with pytest.raises(ffi.HordeError):
        dev.sreshape([7], dev.sconcrete(np.zeros(6)))


# --------------------------------------------------------------- elementwise
EXACT_UN = ["negate", "abs", "signum"]
DOMAIN = {"log": (0.1, 3), "sqrt": (0, 3), "asin": (-.9, .9), "acos": (-.9, .9), "acosh": (1.1, 4), "atanh": (-.9, .9)}


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("op", ffi.UNOPS)
def test_unary_float(dev, oracle, op, dtype):
    rng = np.random.default_rng(1)
    lo, hi = DOMAIN.get(op, (-2, 2))
    a = rnd(rng, (7, 33, 5), dtype, lo, hi)
    for view in (lambda T, t: t, lambda T, t: T.stranspose([2, 0, 1], t), lambda T, t: T.sreplicate(2, T.sindex(t, [1]))):
        got = dev.to_numpy(dev.unary(op, view(dev, dev.sconcrete(a))))
        ref = oracle.unary(op, view(oracle, a))
        if op in EXACT_UN or op in ("recip", "sqrt"):
            exact(got, np.ascontiguousarray(ref))   # IEEE-exact operations
        else:
            close(got, ref, dtype, scale=4)


@pytest.mark.parametrize("dtype", [np.int64, np.int32, np.int16, np.int8])
def test_int_arith(dev, oracle, dtype):
    rng = np.random.default_rng(2)
    a, b = rnd(rng, (5, 9), dtype), rnd(rng, (5, 9), dtype)
    b[0, :3] = 0  # zero divisors: array quotH/remH replace them by 1
    da, db = dev.sconcrete(a), dev.sconcrete(b)
    for op in EXACT_UN:
        exact(dev.to_numpy(dev.unary(op, da)), oracle.unary(op, a).astype(dtype))
    for op in ("add", "sub", "mul", "quot", "rem", "max", "min"):
        exact(dev.to_numpy(dev.binary(op, da, db)), oracle.binary(op, a, b).astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_binary_float(dev, oracle, dtype):
    rng = np.random.default_rng(3)
    a, b = rnd(rng, (6, 130), dtype, 0.1, 2), rnd(rng, (6, 130), dtype, 0.1, 2)
    da, db = dev.sconcrete(a), dev.sconcrete(b)
    for op in ("add", "sub", "mul", "divide", "max", "min"):
        exact(dev.to_numpy(dev.binary(op, da, db)), oracle.binary(op, a, b))
    for op in ("power", "logbase", "atan2"):
        close(dev.to_numpy(dev.binary(op, da, db)), oracle.binary(op, a, b), dtype, scale=8)
    # broadcast (stride-0) and transposed operands, long inner axis (rows kernel)
    big = rnd(rng, (4, 3, 700), dtype)
    bias = rnd(rng, (3,), dtype)
    dv = dev.stranspose([0, 2, 1], dev.sreplicateN([4, 700], dev.sconcrete(bias)))
    rv = oracle.stranspose([0, 2, 1], oracle.sreplicateN([4, 700], bias))
    exact(dev.to_numpy(dev.binary("add", dev.sconcrete(big), dv)), big + rv)
    s = dev.srepl([4, 3, 700], 0.25, dtype)
    exact(dev.to_numpy(dev.binary("mul", s, dev.sconcrete(big))), (np.array(0.25, dtype) * big))


def test_large_vector_path(dev):
    rng = np.random.default_rng(4)
    n = (1 << 22) + 3
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    exact(dev.to_numpy(dev.binary("add", dev.sconcrete(a), dev.sconcrete(b))), a + b)
    off = dev.sslice(1, n - 1, dev.sconcrete(a))     # misaligned view -> scalar fallback
    exact(dev.to_numpy(dev.unary("negate", off)), -a[1:])


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32, np.int16, np.int8])
def test_vector_rows_path_broadcast_views(dev, oracle, dtype):
    """hb_vecrows2_kernel: unit-stride or stride-0 innermost axis, arbitrary
    outer strides (stride-0 sreplicate views, sreverse, slices, permuted outer
    axes), >= 2^16 elements; bit-exact against numpy for every width."""
    rng = np.random.default_rng(41)
    N, Cc, H, W = 9, 6, 28, 64
    x = rnd(rng, (N, Cc, H, W), dtype)
    dx = dev.sconcrete(x)
    bias = rnd(rng, (Cc,), dtype)
    # per-channel operand seen through sreplicate + stranspose (innermost stride 0)
    dv = dev.stranspose([0, 3, 1, 2], dev.sreplicateN([N, H, W], dev.sconcrete(bias)))
    rv = oracle.stranspose([0, 3, 1, 2], oracle.sreplicateN([N, H, W], bias))
    exact(dev.to_numpy(dev.binary("add", dx, dv)), x + rv)
    exact(dev.to_numpy(dev.binary("sub", dv, dx)), rv - x)
    # a row broadcast along every outer axis (innermost stride 1, outer strides 0)
    row = rnd(rng, (W,), dtype)
    exact(dev.to_numpy(dev.binary("mul", dx, dev.sreplicateN([N, Cc, H], dev.sconcrete(row)))), x * row)
    # outer axes permuted / reversed / sliced, innermost axis untouched
    y = rnd(rng, (Cc, N, H, W), dtype)
    exact(dev.to_numpy(dev.binary("add", dx, dev.stranspose([1, 0, 2, 3], dev.sconcrete(y)))), x + y.transpose(1, 0, 2, 3))
    z = rnd(rng, (N + 3, Cc, H, W), dtype)
    dz = dev.sconcrete(z)
    exact(dev.to_numpy(dev.binary("add", dev.sreverse(dx), dev.sslice(2, N, dz))), x[::-1] + z[2:2 + N])
    # both operands broadcast differently: outer product [N*Cc*H] x [W]
    col = rnd(rng, (N, Cc, H), dtype)
    dcol = dev.stranspose([1, 2, 3, 0], dev.sreplicateN([W], dev.sconcrete(col)))
    exact(dev.to_numpy(dev.binary("mul", dcol, dev.sreplicateN([N, Cc, H], dev.sconcrete(row)))), col[..., None] * row)
    # innermost extent not a multiple of the vector width -> the scalar rows kernel, same answer
    x2 = rnd(rng, (N, Cc, H, W + 1), dtype)
    dv2 = dev.stranspose([0, 3, 1, 2], dev.sreplicateN([N, H, W + 1], dev.sconcrete(bias)))
    exact(dev.to_numpy(dev.binary("add", dev.sconcrete(x2), dv2)),
          x2 + oracle.stranspose([0, 3, 1, 2], oracle.sreplicateN([N, H, W + 1], bias)))


def test_full_reductions_block_finish(dev, oracle):
    """ssum0 / sdot0 over > 2^21 elements go through ~1200 partials and the
    block-wide finish kernel; two runs give the same bits (fixed tree)."""
    rng = np.random.default_rng(43)
    for dtype in (np.float64, np.float32):
        n = (1 << 22) + 40
        a, b = rnd(rng, (n,), dtype), rnd(rng, (n,), dtype)
        da, db = dev.sconcrete(a), dev.sconcrete(b)
        s1, s2 = dev.to_numpy(dev.ssum0(da)).reshape(1), dev.to_numpy(dev.ssum0(da)).reshape(1)
        exact(s1, s2)
        close(s1.reshape(()), np.asarray(a.astype(np.float64).sum()).astype(dtype), dtype, scale=100)
        d1, d2 = dev.to_numpy(dev.sdot0(da, db)).reshape(1), dev.to_numpy(dev.sdot0(da, db)).reshape(1)
        exact(d1, d2)
        close(d1.reshape(()), np.asarray(np.dot(a.astype(np.float64), b.astype(np.float64))).astype(dtype), dtype, scale=100)
    # few outputs, long reduced axis: K = 3 outputs, each through its own block of the finish kernel
    t = rng.uniform(-.5, .5, (3, 1 << 20))
    close(dev.to_numpy(dev.ssumN(dev.stranspose([1, 0], dev.sconcrete(t)), 1)), t.sum(axis=1), scale=100)


def test_casts_iota_compare(dev, oracle):
    rng = np.random.default_rng(5)
    for src in ALL_DTYPES:
        a = rnd(rng, (4, 6), src, -3, 3)
        for dst in ALL_DTYPES:
            exact(dev.to_numpy(dev.scast(dev.sconcrete(a), dst)), oracle.scast(a, dst))
    f = rng.uniform(-5, 5, (3, 8))
    for dst in (np.int64, np.int32, np.int16, np.int8):
        exact(dev.to_numpy(dev.sfloor(dev.sconcrete(f), dst)), oracle.sfloor(f, dst))
    for dt in (np.int64, np.float64, np.float32, np.int8):
        exact(dev.to_numpy(dev.siota(11, dt)), oracle.siota(11, dt))
    a = rng.integers(0, 3, (4, 5)).astype(np.float64)
    for b in (a.copy(), a + np.eye(4, 5), a - np.eye(4, 5)[::-1]):
        da, db = dev.sconcrete(a), dev.sconcrete(b)
        assert dev.all_eq(da, db) == oracle.all_eq(a, b)
        assert dev.all_leq(da, db) == oracle.all_leq(a, b)
        assert dev.all_lt(da, db) == oracle.all_lt(a, b)


def test_fused_map(dev, oracle):
    rng = np.random.default_rng(6)
    for dtype in (np.float64, np.float32):
        x, y = rnd(rng, (5, 77), dtype), rnd(rng, (5, 77), dtype)
        a, b = ixm.map_input(0, True), ixm.map_input(1, True)
        expr = ixm.ifH(a <= 0.0, ixm.Ix.lift(0.0), a) * ixm.fun1("tanh", b) + a * b
        got = dev.to_numpy(dev.fused_map(expr, [dev.stranspose([1, 0], dev.sconcrete(x.T.copy())), dev.sconcrete(y)], dtype))
        ref = oracle.fused_map(expr, [x, y], dtype)
        close(got, ref, dtype, scale=4)


def _jit_stats():
    import ctypes as C
    avail, z = C.c_int(-1), [C.c_uint64() for _ in range(3)]
    ffi.call("hb_jit_stats", C.byref(avail), C.byref(z[0]), C.byref(z[1]), C.byref(z[2]))
    return avail.value, z[0].value, z[1].value, z[2].value


def test_fused_map_generated_kernels(dev, oracle):
    """Fused chains through the kernel generator: contiguous operands and a broadcast scalar take the 16-byte
    path, transposed / odd-sized / mixed-dtype operands the scalar one; integer outputs; bit-exact where the
    chain is only + * max select (each step rounds like the reference's separate passes: no FMA contraction)."""
    if os.environ.get("HB_JIT", "1") != "0":
        assert _jit_stats()[0] == 1, "the kernel generator is not available on this box"
    rng = np.random.default_rng(61)
    a, b, c = ixm.map_input(0, True), ixm.map_input(1, True), ixm.map_input(2, True)
    for dtype in (np.float64, np.float32):
        for n in (1 << 12, 4097):
            x, y = rnd(rng, (n,), dtype), rnd(rng, (n,), dtype)
            s = np.array(0.37, dtype=dtype)
            expr = (a * b + c) * ixm.maxH(a, b)
            ins = [dev.sconcrete(x), dev.sconcrete(y), dev.sreplicateN([n], dev.sscalar(float(s), dtype))]
            got = dev.to_numpy(dev.fused_map(expr, ins, dtype))
            exact(got, ((x * y + s) * np.maximum(x, y)).astype(dtype))
            exact(got, oracle.fused_map(expr, [x, y, np.broadcast_to(s, (n,))], dtype))
        # a slice that starts off a 16-byte boundary
        x = rnd(rng, (1025,), dtype)
        got = dev.to_numpy(dev.fused_map(a * a, [dev.sslice(1, 1024, dev.sconcrete(x))], dtype))
        exact(got, x[1:] * x[1:])
    # mixed dtypes: an Int64 operand converted inside the chain, Int64 result by floor
    xi, yf = rng.integers(-50, 50, (7, 33)).astype(np.int64), rnd(rng, (7, 33), np.float64, -3, 3)
    ai, bf = ixm.map_input(0, False), ixm.map_input(1, True)
    expr = ixm.floorH(ixm.fromIntegral(ai) * bf) + ai.remH(7)
    got = dev.to_numpy(dev.fused_map(expr, [dev.sconcrete(xi), dev.stranspose([1, 0], dev.sconcrete(yf.T.copy()))], np.int64))
    exact(got, oracle.fused_map(expr, [xi, yf], np.int64))
    if os.environ.get("HB_JIT", "1") != "0":
        avail, compiled, hits, failed = _jit_stats()
        assert compiled > 0 and failed == 0, (compiled, hits, failed)


def test_interpreter_kernels_agree_with_generated_kernels():
    """HB_JIT=0 runs every index / map program on the bytecode interpreter kernels: the same parity tests must
    pass there (the path of a box without libnvrtc)."""
    import subprocess
    import sys
    if os.environ.get("HB_JIT") == "0":
        pytest.skip("already the interpreter run")
    env = dict(os.environ, HB_JIT="0")
    sel = "gather_kinds or gather_data_dependent or scatter_bit_exact or scatter_long or fused_map or row_kernels or Gather or Scatter"
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-m", "gpu", "-x", "-q", "-k", sel],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_add_accumulate_in_place(dev):
    a = np.arange(12.0).reshape(3, 4)
    acc = dev.sconcrete(a)
    keep = dev.stranspose([1, 0], acc)      # second reference to the storage -> must not mutate
    r = dev.add_accumulate(acc, dev.sconcrete(a))
    assert r is not acc
    exact(dev.to_numpy(acc), a)
    del keep
    r2 = dev.add_accumulate(acc, dev.sconcrete(a))
    assert r2 is acc                        # uniquely referenced -> in place
    exact(dev.to_numpy(acc), 2 * a)


# ---------------------------------------------------------------- reductions
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32])
def test_sums_and_dots(dev, oracle, dtype):
    rng = np.random.default_rng(7)
    a = rnd(rng, (6, 5, 40, 7), dtype)
    b = rnd(rng, (6, 5, 40, 7), dtype)
    da, db = dev.sconcrete(a), dev.sconcrete(b)
    views = [lambda T, t: t, lambda T, t: T.stranspose([3, 2, 1, 0], t), lambda T, t: T.stranspose([1, 3, 0, 2], t)]
    for v in views:
        for n in range(0, 5):
            close(dev.to_numpy(dev.ssumN(v(dev, da), n)), oracle.ssumN(v(oracle, a), n), dtype, scale=10)
        close(dev.to_numpy(dev.sdot0(v(dev, da), v(dev, db))), np.asarray(oracle.sdot0(v(oracle, a), v(oracle, b))), dtype, scale=10)
        close(dev.to_numpy(dev.sdot1In(v(dev, da), v(dev, db))), oracle.sdot1In(v(oracle, a), v(oracle, b)), dtype, scale=10)
    close(dev.to_numpy(dev.ssum0(da)).reshape(()), np.asarray(oracle.ssum0(a)), dtype, scale=10)


def test_bias_gradient_shapes(dev, oracle):
    # ssumN @[N,H,W] over a [N,H,W,C] view of [N,C,H,W] (TestMnistPP.hs:735)
    rng = np.random.default_rng(8)
    t = rng.uniform(-.5, .5, (33, 16, 14, 14))
    v = lambda T, x: T.stranspose([0, 2, 3, 1], x)
    close(dev.to_numpy(dev.ssumN(v(dev, dev.sconcrete(t)), 3)), oracle.ssumN(v(oracle, t), 3), scale=10)
    big = rng.uniform(-.5, .5, (3_000_017,))
    close(dev.to_numpy(dev.ssum0(dev.sconcrete(big))).reshape(()), np.asarray(big.sum()), scale=100)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_matmul_family(dev, oracle, dtype):
    rng = np.random.default_rng(9)
    for (m, k, n) in [(64, 784, 70), (10, 64, 33), (1, 5, 1), (130, 17, 129), (3, 1, 4), (64, 4099, 784),
                      (512, 512, 300), (65, 333, 63), (17, 2000, 19)]:
        a, b = rnd(rng, (m, k), dtype), rnd(rng, (k, n), dtype)
        ref = (a.astype(np.float64) @ b.astype(np.float64)).astype(dtype)
        absref = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
        chk = lambda got, sl=slice(None): close_dot(dev.to_numpy(got), ref[sl], absref[sl], k, dtype)
        chk(dev.smatmul2(dev.sconcrete(a), dev.sconcrete(b)))
        chk(dev.smatmul2(dev.stranspose([1, 0], dev.sconcrete(a.T.copy())), dev.sconcrete(b)))
        # every operand orientation (the four staging layouts of the GEMM kernels) and odd offsets
        bt = dev.stranspose([1, 0], dev.sconcrete(b.T.copy()))
        chk(dev.smatmul2(dev.sconcrete(a), bt))
        chk(dev.smatmul2(dev.stranspose([1, 0], dev.sconcrete(a.T.copy())), bt))
        if m > 2:
            chk(dev.smatmul2(dev.sslice(1, m - 1, dev.sconcrete(a)), dev.sconcrete(b)), slice(1, None))
        v = rnd(rng, (k,), dtype)
        close_dot(dev.to_numpy(dev.smatvecmul(dev.sconcrete(a), dev.sconcrete(v))),
                  (a.astype(np.float64) @ v.astype(np.float64)).astype(dtype),
                  np.abs(a).astype(np.float64) @ np.abs(v).astype(np.float64), k, dtype)
    # the reference's literal formulation through broadcast views (OpsConcrete.hs:237-242)
    a, b = rnd(rng, (40, 50), dtype), rnd(rng, (50, 30), dtype)
    lit = lambda T, x, y: T.sdot1In(T.stranspose([1, 0], T.sreplicate(30, x)), T.stranspose([0, 2, 1], T.sreplicate(40, y)))
    close_dot(dev.to_numpy(lit(dev, dev.sconcrete(a), dev.sconcrete(b))), lit(oracle, a, b),
              np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64), 50, dtype)


@pytest.mark.parametrize("shape", [(1024, 320, 1200),    # eight-warp kernel, one K slice, ragged N (304 tiles)
                                   (512, 512, 1024),     # one tile per SM: two K slices accumulated serially into C
                                   (500, 518, 1001),     # the same with ragged edges and an odd row stride of C
                                   (100, 4100, 200),     # deep K over 8 tiles: many slices + fixed-order reduce
                                   (512, 1024, 512),     # four-warp kernel, split-K across CTAs
                                   (128, 1030, 3136)])   # 98 tiles: three K slices accumulated serially into C
def test_dmma_gemm_variants_and_determinism(dev, shape):
    m, k, n = shape
    rng = np.random.default_rng(91)
    a, b = rnd(rng, (m, k)), rnd(rng, (k, n))
    ref = a @ b
    for ta in (False, True):
        for tb in (False, True):
            A = dev.stranspose([1, 0], dev.sconcrete(a.T.copy())) if ta else dev.sconcrete(a)
            B = dev.stranspose([1, 0], dev.sconcrete(b.T.copy())) if tb else dev.sconcrete(b)
            got = dev.to_numpy(dev.smatmul2(A, B))
            close(got, ref, np.float64, scale=k)
            for _ in range(3):       # K slices are combined in a fixed order: bit-identical from run to run
                assert np.array_equal(dev.to_numpy(dev.smatmul2(A, B)), got)


@pytest.mark.parametrize("shape", [(512, 1024, 512),     # four-warp kernel + reduce kernel that adds the old C
                                   (512, 512, 1024),     # two K slices accumulated serially, slice 0 adds the old C
                                   (1024, 320, 1200),    # eight-warp kernel, one slice
                                   (128, 1030, 3136),    # three serial slices
                                   (64, 70, 48),         # four-warp kernel, one slice
                                   (512, 1024, 28),      # thin N
                                   (3, 5, 2)])           # no GEMM path: product + add
def test_matmul2_acc(dev, shape):
    """hb_matmul2_acc = taddTarget acc (tsmatmul2 a b): in place through the GEMM epilogue when `acc` is
    uniquely referenced, a fresh array otherwise; same value either way (1e-12 relative)."""
    m, k, n = shape
    rng = np.random.default_rng(92)
    a, b, c0 = rnd(rng, (m, k)), rnd(rng, (k, n)), rnd(rng, (m, n))
    ref = c0 + a @ b
    A, B = dev.sconcrete(a), dev.stranspose([1, 0], dev.sconcrete(b.T.copy()))
    acc = dev.sconcrete(c0)
    out = dev.smatmul2_acc(acc, A, B)
    close(dev.to_numpy(out), ref, np.float64, scale=k)
    if min(m, n) >= 16:
        assert out is acc                                   # accumulated in place
    out2 = dev.smatmul2_acc(out, A, B)                      # and again (the cotangent of a weight used twice)
    close(dev.to_numpy(out2), ref + a @ b, np.float64, scale=2 * k)
    # an accumulator that somebody else still holds is not touched
    acc = dev.sconcrete(c0)
    alias = dev.sreshape([m, n], acc)                       # a second handle on the same storage
    out = dev.smatmul2_acc(acc, A, B)
    assert out is not acc
    exact(dev.to_numpy(alias), c0)
    close(dev.to_numpy(out), ref, np.float64, scale=k)
    # a transposed (non-contiguous) accumulator: product + add
    acct = dev.stranspose([1, 0], dev.sconcrete(c0.T.copy()))
    close(dev.to_numpy(dev.smatmul2_acc(acct, A, B)), ref, np.float64, scale=k)
    # deterministic
    r1 = dev.to_numpy(dev.smatmul2_acc(dev.sconcrete(c0), A, B))
    r2 = dev.to_numpy(dev.smatmul2_acc(dev.sconcrete(c0), A, B))
    assert np.array_equal(r1, r2)


@pytest.mark.parametrize("shape", [(512, 512, 28, 1024),   # the recurrent layer: K = 512 and K = 28, one tile per SM each
                                   (512, 512, 512, 1024),
                                   (512, 1024, 1024, 512),  # the weight cotangents: 64 tiles, two K slices per product
                                   (70, 33, 129, 50),       # ragged tiles, odd strides
                                   (24, 28, 24, 9)])
def test_matmul2_pairs(dev, shape):
    """hb_matmul2_sum / _pair / _acc_pair (two products per launch) against numpy, every operand orientation
    the layer and its transposition rules use; bit-identical from run to run."""
    m, k1, k2, n = shape
    rng = np.random.default_rng(94)
    a1, b1, a2, b2 = rnd(rng, (m, k1)), rnd(rng, (k1, n)), rnd(rng, (m, k2)), rnd(rng, (k2, n))
    T = lambda x: dev.stranspose([1, 0], dev.sconcrete(np.ascontiguousarray(x.T)))
    for ta in (False, True):
        for tb in (False, True):
            A1, A2 = (T(a1), T(a2)) if ta else (dev.sconcrete(a1), dev.sconcrete(a2))
            B1, B2 = (T(b1), T(b2)) if tb else (dev.sconcrete(b1), dev.sconcrete(b2))
            got = dev.to_numpy(dev.smatmul2_sum(A1, B1, A2, B2))
            close(got, a1 @ b1 + a2 @ b2, np.float64, scale=k1 + k2)
            assert np.array_equal(dev.to_numpy(dev.smatmul2_sum(A1, B1, A2, B2)), got)
            if k1 == k2:
                o1, o2 = dev.smatmul2_pair(A1, B1, A2, B2)
                close(dev.to_numpy(o1), a1 @ b1, np.float64, scale=k1)
                close(dev.to_numpy(o2), a2 @ b2, np.float64, scale=k2)
                c1, c2 = rnd(rng, (m, n)), rnd(rng, (m, n))
                acc1, acc2 = dev.sconcrete(c1), dev.sconcrete(c2)
                r1, r2 = dev.smatmul2_acc_pair(acc1, A1, B1, acc2, A2, B2)
                assert r1 is acc1 and r2 is acc2                    # both accumulated in place
                close(dev.to_numpy(r1), c1 + a1 @ b1, np.float64, scale=k1)
                close(dev.to_numpy(r2), c2 + a2 @ b2, np.float64, scale=k2)
                # one accumulator shared with somebody else: nothing is updated in place behind their back
                acc1, acc2 = dev.sconcrete(c1), dev.sconcrete(c2)
                alias = dev.sreshape([m, n], acc2)
                r1, r2 = dev.smatmul2_acc_pair(acc1, A1, B1, acc2, A2, B2)
                exact(dev.to_numpy(alias), c2)
                close(dev.to_numpy(r1), c1 + a1 @ b1, np.float64, scale=k1)
                close(dev.to_numpy(r2), c2 + a2 @ b2, np.float64, scale=k2)
    # mixed orientations between the two products, float32, broadcast operands: the single-product entry points
    got = dev.to_numpy(dev.smatmul2_sum(dev.sconcrete(a1), dev.sconcrete(b1), T(a2), dev.sconcrete(b2)))
    close(got, a1 @ b1 + a2 @ b2, np.float64, scale=k1 + k2)
    f = lambda x: dev.sconcrete(x.astype(np.float32))
    got = dev.to_numpy(dev.smatmul2_sum(f(a1), f(b1), f(a2), f(b2)))
    close(got, (a1 @ b1 + a2 @ b2).astype(np.float32), np.float32, scale=k1 + k2)


def test_grad_accumulates_weight_cotangents_in_the_gemm(dev, oracle):
    """A weight used at every step of an unrolled loop: the sweep hands its cotangent contributions to
    smatmul2_acc (ad.DeferredMatmul); same gradient as the oracle's matmul + add."""
    rng = np.random.default_rng(93)
    w, x = rng.uniform(-.3, .3, (64, 64)), rng.uniform(-.5, .5, (64, 96))

    def f(T, W, X):
        h = X
        for _ in range(4):
            h = T.unary("tanh", T.smatmul2(W, h))
        return T.ssum0(T.binary("mul", h, h))
    dv, dg = ad.grad(dev, f, [dev.sconcrete(w), dev.sconcrete(x)])
    ov, og = ad.grad(oracle, f, [w, x])
    close(dev.to_numpy(dv).reshape(()), np.asarray(ov), scale=100)
    for g, o in zip(dg, og):
        close(dev.to_numpy(g), np.asarray(o), scale=1000)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int8])
def test_transposed_views_materialise_bit_exact(dev, dtype):
    """hb_contiguous / reshape of transposed views (tiled shared-memory transpose fast path and the generic
    iterator around it): ragged tile edges, trailing axes that collapse, ones that do not."""
    rng = np.random.default_rng(77)
    for shape in [(4096, 784), (33, 65), (32, 32), (100, 7), (5, 300), (1000, 3)]:
        a = (rng.uniform(-100, 100, shape)).astype(dtype)
        t = dev.stranspose([1, 0], dev.sconcrete(a))                       # [C, R] view with unit stride along axis 0
        exact(dev.to_numpy(dev.sreshape([shape[1] * shape[0]], t)), np.ascontiguousarray(a.T).reshape(-1))
    a = rng.uniform(-1, 1, (48, 4, 6, 40)).astype(dtype)                   # [N, c, h, w] -> view [w, N, c, h]: collapses
    v = dev.stranspose([3, 0, 1, 2], dev.sconcrete(a))
    exact(dev.to_numpy(dev.sreshape([40 * 48 * 24], v)), np.ascontiguousarray(np.transpose(a, (3, 0, 1, 2))).reshape(-1))
    v = dev.stranspose([3, 1, 0, 2], dev.sconcrete(a))                     # does not collapse: generic path
    exact(dev.to_numpy(dev.sreshape([40 * 48 * 24], v)), np.ascontiguousarray(np.transpose(a, (3, 1, 0, 2))).reshape(-1))


def test_shared_cotangents_are_not_updated_in_place(dev, oracle):
    """`add` hands the same cotangent array to both operands; a later accumulation into one of them
    must not change what the other one sees (in-place taddTarget only into arrays the sweep owns)."""
    rng = np.random.default_rng(5)
    x = rnd(rng, (6, 50))

    def f(T, x):
        a, b = T.unary("exp", x), T.unary("sin", x)
        q = T.unary("cos", a)                      # a second consumer of a, created before the add
        return T.ssum0(T.binary("add", T.binary("add", a, b), q))
    _, (g,) = ad.grad(dev, f, [dev.sconcrete(x)])
    _, (go,) = ad.grad(oracle, f, [x])
    close(dev.to_numpy(g), go, np.float64, scale=10)
    ref = np.exp(x) + np.cos(x) - np.sin(np.exp(x)) * np.exp(x)
    close(dev.to_numpy(g), ref, np.float64, scale=10)


def test_argmax_argmin(dev, oracle):
    rng = np.random.default_rng(10)
    a = rng.integers(0, 4, (5, 6, 9)).astype(np.float64)     # many ties: first extremum wins
    exact(dev.to_numpy(dev.sargMax(dev.sconcrete(a))), oracle.sargMax(a))
    exact(dev.to_numpy(dev.sargMin(dev.sconcrete(a))), oracle.sargMin(a))
    v = lambda T, x: T.stranspose([2, 0, 1], x)
    exact(dev.to_numpy(dev.sargMax(v(dev, dev.sconcrete(a)))), oracle.sargMax(v(oracle, a)))


# ---------------------------------------------------------- gather / scatter
@pytest.mark.parametrize("name,prog,inp,expected", P.GATHER_GOLDEN, ids=[g[0] for g in P.GATHER_GOLDEN])
def test_reference_gather_scatter_programs(dev, oracle, name, prog, inp, expected):
    # values on VALS72-like data: bit-exact vs oracle; gradients: the reference's goldens
    data = P.VALS72 if inp.shape == (7, 2) else np.array([0.1, -3.0, 7.5, 2.25])
    exact(dev.to_numpy(prog(dev, dev.sconcrete(data))), np.ascontiguousarray(prog(oracle, data)))
    _, (g,) = ad.grad(dev, lambda T, t: prog(T, t), [dev.sconcrete(inp)])
    np.testing.assert_allclose(dev.to_numpy(g).reshape(-1), np.array(expected, dtype=float), atol=1e-10, rtol=0)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_gather_kinds_bit_exact(dev, oracle, dtype):
    rng = np.random.default_rng(11)
    src = rnd(rng, (9, 7, 5), dtype)
    d = dev.sconcrete(src)
    tbl = rng.integers(-2, 11, (6, 4)).astype(np.int64)
    cases = [
        ([4, 3], 1, lambda T: (lambda i: [i[0] + i[1]])),                       # affine in bounds -> view
        ([9, 4], 1, lambda T: (lambda i: [i[0] + i[1]])),                       # affine, high-side OOB
        ([5, 5], 2, lambda T: (lambda i: [2 * i[0] - 1, i[1] + i[0] - 2])),     # affine, both-side OOB
        ([6, 4], 2, lambda T: (lambda i: [T.ix_tbl(T.sconcrete(tbl), [i[0], i[1]]), i[1].remH(3) + i[0].quotH(2)])),
        ([6, 4], 3, lambda T: (lambda i: [ixm.ifH((i[0] <= 2) & ~(i[1].eq(1)), i[0], 8 - i[0]), abs(i[1] - 2), ixm.maxH(i[0] - i[1], 0)])),
        ([3], 0, lambda T: (lambda i: [])),                                     # replicate-like
    ]
    for shm, p, mk in cases:
        got = dev.to_numpy(dev.sgather(shm, d, mk(dev), p))
        ref = oracle.sgather(shm, src, mk(oracle), p)
        exact(got, np.ascontiguousarray(ref))
    # strided (transposed) source
    got = dev.to_numpy(dev.sgather([5, 2], dev.stranspose([2, 0, 1], d), lambda i: [i[0] - i[1]], 1))
    exact(got, oracle.sgather([5, 2], np.transpose(src, (2, 0, 1)), lambda i: [i[0] - i[1]], 1))


def test_gather_data_dependent(dev, oracle):
    rng = np.random.default_rng(12)
    u = rng.integers(0, 3, (4, 5, 6)).astype(np.float64)   # ties
    prog = lambda T, x: T.sgather([4, 5], x, lambda i: [i[0], i[1], T.ix_argmax(x, i)], 3)
    exact(dev.to_numpy(prog(dev, dev.sconcrete(u))), prog(oracle, u))
    x = rng.uniform(-.5, .5, (3, 2, 6, 8))
    exact(dev.to_numpy(nn.reluS_vectorised(dev, dev.sconcrete(x))), nn.reluS_vectorised(oracle, x))
    exact(dev.to_numpy(nn.maxPool2dUnpaddedS_vectorised(dev, 2, 2, dev.sconcrete(x))), nn.maxPool2dUnpaddedS_vectorised(oracle, 2, 2, x))


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32, np.int16, np.int8])
def test_scatter_bit_exact_order(dev, oracle, dtype):
    rng = np.random.default_rng(13)
    src = rnd(rng, (12, 10, 3), dtype)
    d = dev.sconcrete(src)
    cases = [
        ([12, 10], 2, lambda i: [i[0], i[1]]),                          # injective -> direct path
        ([5], 2, lambda i: [(i[0] + i[1]).remH(7) - 1]),                # collisions + drops
        ([3, 2], 2, lambda i: [i[0].quotH(5), i[1].remH(2)]),           # heavier collisions
        ([4], 1, lambda i: [ixm.minH(i[0], 3)]),                        # slices of rank 2
        ([2], 3, lambda i: [ixm.minH(i[0] + i[1] + i[2], 1)]),          # rank-0 slices, long lists
    ]
    for shp, m, f in cases:
        got = dev.to_numpy(dev.sscatter(shp, d, f, m))
        ref = oracle.sscatter(shp, src, f, m)
        exact(got, ref)   # same accumulation order as the reference => identical bits
    # transposed source
    got = dev.to_numpy(dev.sscatter([6], dev.stranspose([1, 0], d), lambda i: [i[0].quotH(2) + i[1].remH(2)], 2))
    exact(got, oracle.sscatter([6], np.transpose(src, (1, 0, 2)), lambda i: [i[0].quotH(2) + i[1].remH(2)], 2))


def test_scatter_long_segments(dev, oracle):
    rng = np.random.default_rng(14)
    src = rng.uniform(-.5, .5, (5000,))
    f = lambda i: [ixm.ifH(i[0] <= 4000, 0, i[0].remH(3) + 1)]     # 4001 contributions to slot 0
    exact(dev.to_numpy(dev.sscatter([4], dev.sconcrete(src), f, 1)), oracle.sscatter([4], src, f, 1))
    src2 = rng.uniform(-.5, .5, (300, 4))
    f2 = lambda i: [i[0].remH(2)]                                    # 150 slices per destination
    exact(dev.to_numpy(dev.sscatter([2], dev.sconcrete(src2), f2, 1)), oracle.sscatter([2], src2, f2, 1))


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int16, np.int8])
@pytest.mark.parametrize("W", [8, 7, 16, 40])
def test_gather_scatter_row_kernels_window_order_and_16_byte_paths(dev, oracle, dtype, W):
    """The im2col-style gather out[i,j,n,c,:] = x[n,c,i+j,:] and its adjoint: overlapping windows are walked
    window-axes-inside (every row carries its own result offset), runs move by 16-byte accesses when the run,
    the strides and the base allow it (W = 7: never; sliced sources with odd bases: never)."""
    rng = np.random.default_rng(16)
    N, Cc, H, KH = 3, 4, 6, 3
    x = rnd(rng, (N, Cc, H, W), dtype)
    f = lambda i: [i[0] + i[1]]
    for perm in ([2, 0, 1, 3], [2, 1, 0, 3]):
        xh, d = np.transpose(x, perm), dev.stranspose(perm, dev.sconcrete(x))
        got = dev.sgather([H, KH], d, f, 1)
        ref = oracle.sgather([H, KH], xh, f, 1)
        exact(dev.to_numpy(got), np.ascontiguousarray(ref))
        exact(dev.to_numpy(dev.sscatter([H], got, f, 2)), oracle.sscatter([H], np.ascontiguousarray(ref), f, 2))
    # contiguous source: the trailing axes merge into one long run (> 32 accesses)
    exact(dev.to_numpy(dev.sgather([N + 1, 2], dev.sconcrete(x), f, 1)), np.ascontiguousarray(oracle.sgather([N + 1, 2], x, f, 1)))
    # odd base offset (a slice of a flat buffer) and a both-sides out-of-bounds two-component window
    flat = rnd(rng, (1 + H * H * W,), dtype)
    v, dv = flat[1:].reshape(H, H, W), dev.sreshape([H, H, W], dev.sslice(1, H * H * W, dev.sconcrete(flat)))
    g = lambda i: [i[0] + i[1] - 1, i[2] - i[0] + 1]
    got = dev.sgather([H, 3, H], dv, g, 2)
    ref = oracle.sgather([H, 3, H], v, g, 2)
    exact(dev.to_numpy(got), np.ascontiguousarray(ref))
    exact(dev.to_numpy(dev.sscatter([H, H], got, g, 3)), oracle.sscatter([H, H], np.ascontiguousarray(ref), g, 3))
    # data-dependent index (offset table) over slices: table offsets are multiples of the access width or not
    tbl = rng.integers(-1, H + 1, (5,)).astype(np.int64)
    for src_np, src_d in ((v, dv), (x[0], dev.sindex(dev.sconcrete(x), [0]))):
        h = lambda T: (lambda i: [T.ix_tbl(T.sconcrete(tbl), [i[0]])])
        exact(dev.to_numpy(dev.sgather([5], src_d, h(dev), 1)), np.ascontiguousarray(oracle.sgather([5], src_np, h(oracle), 1)))
    # scatter with untouched destinations (zeros written by the sum kernel), with and without collisions
    y = rnd(rng, (4, 6, W), dtype)
    for k in (lambda i: [2 * i[0]], lambda i: [i[0].quotH(2) * 3]):
        exact(dev.to_numpy(dev.sscatter([9], dev.sconcrete(y), k, 1)), oracle.sscatter([9], y, k, 1))


def test_gather_scatter_adjoint(dev):
    # checkAdjoint (bench/ConvVjpBench.hs:564-584) at the gather48 shapes (:428-439)
    rng = np.random.default_rng(15)
    x = dev.sconcrete(rng.uniform(-.5, .5, (50, 3, 3, 50)))
    f = lambda i: [i[0] + i[1]]
    g = dev.sgather([48, 3], x, f, 1)
    y = dev.sconcrete(rng.uniform(-.5, .5, dev.shape(g)))
    lhs = dev.kfromS(dev.sdot0(g, y))
    rhs = dev.kfromS(dev.sdot0(x, dev.sscatter([50], y, f, 2)))
    assert abs(lhs - rhs) <= 1e-12 * max(1.0, abs(lhs))


# -------------------------------------------------------------------- conv
@pytest.mark.parametrize("name,K,A,wrt,expected,eps", P.CONV_GOLDEN, ids=[c[0] for c in P.CONV_GOLDEN])
@pytest.mark.parametrize("form", ["fused", "vectorised"])
def test_conv_reference_goldens(dev, form, name, K, A, wrt, expected, eps):
    conv = nn.conv2dPreservingS if form == "fused" else nn.conv2dPreservingS_vectorised
    Kt, At = dev.sconcrete(K), dev.sconcrete(A)
    if wrt == "A":
        _, (g,) = ad.grad(dev, lambda T, a: conv(T, Kt, a), [At])
    else:
        _, (g,) = ad.grad(dev, lambda T, k: conv(T, k, At), [Kt])
    np.testing.assert_allclose(dev.to_numpy(g).reshape(-1), np.array(expected), atol=eps, rtol=0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(3, 3, 3, 6, 6, 3, 3), (2, 5, 7, 9, 4, 2, 4), (5, 1, 16, 12, 12, 5, 5), (1, 1, 1, 1, 1, 1, 1),
                                   (4, 16, 16, 14, 14, 5, 5), (2, 3, 2, 5, 0, 3, 3), (3, 2, 2, 3, 3, 4, 5),
                                   # DMMA kernels: row bands + channel chunks, several output-channel tiles,
                                   # odd widths (8-byte cp.async path), ragged last band, tap-k / tap-column modes
                                   (3, 64, 64, 28, 28, 3, 3), (2, 20, 72, 9, 7, 2, 3), (2, 8, 40, 30, 30, 3, 3),
                                   (9, 1, 16, 28, 28, 5, 5), (301, 16, 16, 14, 14, 5, 5), (3, 2, 9, 31, 17, 3, 6),
                                   (2, 12, 24, 50, 20, 1, 1), (1, 4, 4, 40, 230, 3, 3),
                                   # the reference's own ConvVjpBench sweep (bench/ConvVjpBench.hs:850-859)
                                   (3, 3, 3, 24, 24, 3, 3), (3, 3, 3, 96, 96, 3, 3), (3, 3, 3, 192, 192, 3, 3)])
def test_conv_fwd_vjp_vs_oracle(dev, oracle, dtype, shape):
    N, Ci, Co, H, W, KH, KW = shape
    rng = np.random.default_rng(16)
    K, A, dB = rnd(rng, (Co, Ci, KH, KW), dtype), rnd(rng, (N, Ci, H, W), dtype), rnd(rng, (N, Co, H, W), dtype)
    dK_, dA_, dB_ = dev.sconcrete(K), dev.sconcrete(A), dev.sconcrete(dB)
    aK, aA, adB = np.abs(K), np.abs(A), np.abs(dB)
    fwd_abs = oracle.conv2dPreserving(aK, aA)
    close_dot(dev.to_numpy(dev.conv2dPreserving(dK_, dA_)), oracle.conv2dPreserving(K, A), fwd_abs, Ci * KH * KW, dtype)
    close_dot(dev.to_numpy(dev.conv2d_dkrn(dA_, dB_, KH, KW)), oracle.conv2d_dkrn(A, dB, KH, KW),
              oracle.conv2d_dkrn(aA, adB, KH, KW), N * H * W, dtype)
    close_dot(dev.to_numpy(dev.conv2d_dinp(dK_, dB_)), oracle.conv2d_dinp(K, dB), oracle.conv2d_dinp(aK, adB),
              Co * KH * KW, dtype)
    if N * H * W and dtype == np.float64:
        close_dot(dev.to_numpy(nn.conv2dPreservingS_vectorised(dev, dK_, dA_)), oracle.conv2dPreserving(K, A), fwd_abs,
                  Ci * KH * KW, dtype)


def test_conv_non_finite_input_stays_local(dev, oracle):
    """An Inf in the input reaches exactly the outputs whose window covers it (as
    on the oracle): the zero padding of the k dimension in the tap-k DMMA mode
    must not turn 0 * Inf into NaN elsewhere."""
    rng = np.random.default_rng(33)
    K, A = rnd(rng, (16, 1, 5, 5)), rnd(rng, (3, 1, 28, 28))
    A[1, 0, 10, 12] = np.inf
    got = dev.to_numpy(dev.conv2dPreserving(dev.sconcrete(K), dev.sconcrete(A)))
    ref = oracle.conv2dPreserving(K, A)
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    np.testing.assert_allclose(got[np.isfinite(ref)], ref[np.isfinite(ref)], rtol=1e-11, atol=1e-13)


def test_conv_padding_taps_and_non_finite_weights():
    """The DMMA kernels do not compute the taps that fall entirely into the zero padding.  That is
    only observable with a non-finite WEIGHT: the reference's gather-then-multiply formulation turns
    0 * Inf at a padded tap into NaN (the oracle does the same); HB_CONV_FULL_PADDING=1 computes every
    tap and reproduces that pattern exactly.  (Own process: the switch is read by hb_init.)"""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import numpy as np
        from horde_ad_b200.device import DeviceBackend
        from oracle.concrete import ConcreteBackend
        dev, orc = DeviceBackend(0), ConcreteBackend()
        rng = np.random.default_rng(3)
        K, A = rng.uniform(-.5, .5, (16, 16, 5, 5)), rng.uniform(-.5, .5, (4, 16, 14, 14))
        K[3, 2, 4, 1] = np.inf                                   # bottom tap row
        got = dev.to_numpy(dev.conv2dPreserving(dev.sconcrete(K), dev.sconcrete(A)))
        ref = orc.conv2dPreserving(K, A)
        same_nan = np.array_equal(np.isnan(got), np.isnan(ref))
        fin = np.isfinite(ref)
        close = np.allclose(got[fin], ref[fin], rtol=1e-11, atol=1e-12)
        extra_finite = int((np.isnan(ref) & ~np.isnan(got)).sum())
        print("RESULT", int(same_nan), int(close), extra_finite, int(np.isnan(ref).sum()))
    """)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    def run(env_extra):
        env = dict(os.environ, PYTHONPATH=root, **env_extra)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd=root)
        assert out.returncode == 0, out.stderr[-2000:]
        return [int(x) for x in out.stdout.split("RESULT")[1].split()]
    same_nan, close, extra, nnan = run({"HB_CONV_FULL_PADDING": "1"})
    assert same_nan == 1 and close == 1 and nnan > 0           # every tap computed: the reference's NaN pattern
    same_nan, close, extra, nnan = run({"HB_CONV_FULL_PADDING": "0"})
    assert close == 1 and extra > 0 and extra < nnan            # default: finite wherever the Inf tap only met padding


def test_conv_full_size_adjoint_property(dev):
    """ConvVjpBench at BASELINE's size: size-independent checks (adjointness,
    linearity) instead of an oracle run."""
    rng = np.random.Generator(np.random.Philox(42))
    N, C, H, W, KH, KW = 256, 64, 28, 28, 3, 3
    K = dev.sconcrete(rng.uniform(-.5, .5, (C, C, KH, KW)))
    A = dev.sconcrete(rng.uniform(-.5, .5, (N, C, H, W)))
    Y = dev.sconcrete(rng.uniform(-.5, .5, (N, C, H, W)))
    B = dev.conv2dPreserving(K, A)
    lhs = dev.kfromS(dev.sdot0(B, Y))
    r1 = dev.kfromS(dev.sdot0(A, dev.conv2d_dinp(K, Y)))
    r2 = dev.kfromS(dev.sdot0(K, dev.conv2d_dkrn(A, Y, KH, KW)))
    assert abs(lhs - r1) <= 1e-10 * abs(lhs) and abs(lhs - r2) <= 1e-10 * abs(lhs)
    B2 = dev.conv2dPreserving(K, dev.binary("add", A, A))
    assert dev.all_eq(B2, dev.binary("add", B, B))      # exact: scaling by 2 commutes with rounding


# ---------------------------------------------------------------- NN blocks
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_nn_blocks(dev, oracle, dtype):
    rng = np.random.default_rng(17)
    x = rnd(rng, (5, 3, 9, 10), dtype)
    x[0, 0, :2, :2] = 0.25     # ties inside a window
    exact(dev.to_numpy(dev.relu(dev.sconcrete(x))), oracle.relu(x))
    for ks, st in [(2, 2), (3, 2), (2, 3), (3, 3)]:
        y, am = dev.maxpool2d(dev.sconcrete(x), ks, st)
        ry, ram = oracle.maxpool2d(x, ks, st)
        exact(dev.to_numpy(y), ry)
        exact(dev.to_numpy(am), ram)
        dy = rnd(rng, ry.shape, dtype)
        got = dev.to_numpy(dev.maxpool2d_bwd(dev.sconcrete(dy), am, ks, st, 9, 10))
        ref = oracle.maxpool2d_bwd(dy, ram, ks, st, 9, 10)
        if ks <= st:
            exact(got, ref)
        else:
            close(got, ref, dtype)
    lg, ex = rnd(rng, (10, 37), dtype, -2, 2), np.eye(10, dtype=dtype)[:, rng.integers(0, 10, 37)]
    l, dl = dev.softmax_xent(dev.sconcrete(lg), dev.sconcrete(ex), 1 / 37)
    rl, rdl = oracle.softmax_xent(lg, ex, 1 / 37)
    close(dev.to_numpy(l).reshape(()), np.asarray(rl), dtype, scale=100)
    close(dev.to_numpy(dl), rdl, dtype, scale=100)
    if dtype == np.float64:
        close(dev.to_numpy(nn.lossSoftMaxCrossEntropyS_vectorised(dev, dev.sconcrete(ex), dev.sconcrete(lg), 1 / 37)).reshape(()),
              np.asarray(rl), dtype, scale=100)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,ks", [((5, 3, 10, 12), 2), ((3, 4, 9, 11), 2), ((2, 2, 9, 10), 3), ((130, 16, 28, 28), 2),
                                      ((2, 3, 4, 4), 1)])
def test_fused_relu_pool(dev, oracle, dtype, shape, ks):
    """hb_relu_pool_fwd/bwd == reluS . (+bias) then maxPool2dUnpaddedS and their
    adjoints (bit-exact forward incl. ties and -0.0; cotangents exact, bias
    cotangent to reduction tolerance)."""
    rng = np.random.default_rng(21)
    pre, bias = rnd(rng, shape, dtype), rnd(rng, (shape[1],), dtype)
    pre[0, 0, :2, :2] = 0.25 - bias[0]
    pre[0, 1, :2, :2] = -1.0          # an all-inactive window
    y, code = dev.relu_pool_fwd(dev.sconcrete(pre), dev.sconcrete(bias), ks)
    ry, rcode = oracle.relu_pool_fwd(pre, bias, ks)
    exact(dev.to_numpy(y), ry.astype(dtype))
    exact(dev.to_numpy(code), rcode)
    # the same through the unfused device chain
    N, Cc, H, W = shape
    stretched = dev.stranspose([0, 3, 1, 2], dev.sreplicateN([N, H, W], dev.sconcrete(bias)))
    chain, _ = dev.maxpool2d(dev.relu(dev.binary("add", dev.sconcrete(pre), stretched)), ks, ks)
    exact(dev.to_numpy(y), dev.to_numpy(chain))
    dy = rnd(rng, ry.shape, dtype)
    dp, db = dev.relu_pool_bwd(dev.sconcrete(dy), code, ks, H, W)
    rdp, rdb = oracle.relu_pool_bwd(dy, rcode, ks, H, W)
    exact(dev.to_numpy(dp), rdp.astype(dtype))
    close(dev.to_numpy(db), rdb.astype(dtype), dtype, scale=100)
    y2, _ = dev.relu_pool_fwd(dev.sconcrete(pre), None, ks)
    exact(dev.to_numpy(y2), oracle.relu_pool_fwd(pre, None, ks)[0].astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,ks", [((37, 1, 16, 28, 28, 5, 5), 2), ((9, 16, 16, 14, 14, 5, 5), 2), ((3, 3, 40, 12, 20, 3, 3), 2),
                                      ((2, 5, 7, 9, 11, 2, 3), 2), ((2, 4, 4, 12, 12, 3, 3), 3), ((5, 8, 64, 28, 28, 3, 3), 2),
                                      ((4, 2, 9, 30, 6, 3, 2), 2),
                                      # warp-autonomous one-channel kernels (conv_strip.cu): widths 26 / 28, odd item counts
                                      ((5, 1, 9, 28, 26, 3, 4), 2), ((6, 1, 8, 4, 28, 2, 2), 2), ((131, 1, 16, 28, 28, 5, 5), 2),
                                      ((3, 1, 5, 6, 28, 1, 1), 2),
                                      # ... with several groups of 16 output channels per strip
                                      ((21, 1, 64, 28, 28, 3, 3), 2), ((7, 1, 32, 28, 28, 5, 5), 2), ((5, 1, 24, 10, 26, 2, 4), 2),
                                      ((4, 1, 40, 28, 28, 5, 5), 2), ((150, 1, 17, 28, 28, 3, 3), 2)])
def test_fused_conv_layer(dev, oracle, dtype, shape, ks):
    """hb_conv_relu_pool_fwd/bwd (a whole convMnistLayerS as one node): the fused
    DMMA epilogue / pooled-cotangent dKrn path and its unfused fallbacks agree
    with conv -> (+bias) -> relu -> maxpool and the adjoints on the oracle."""
    N, Ci, Co, H, W, KH, KW = shape
    rng = np.random.default_rng(31)
    K, A, bias = rnd(rng, (Co, Ci, KH, KW), dtype), rnd(rng, (N, Ci, H, W), dtype), rnd(rng, (Co,), dtype)
    dK_, dA_, db_ = dev.sconcrete(K), dev.sconcrete(A), dev.sconcrete(bias)
    y, code = dev.conv_relu_pool_fwd(dK_, db_, dA_, ks)
    ry, rcode = oracle.conv_relu_pool_fwd(K, bias, A, ks)
    aK, aA, ab = np.abs(K), np.abs(A), np.abs(bias)
    # relu and max are 1-Lipschitz: the pooled error is bounded by the largest pre-activation bound of the window
    y_abs = oracle.conv_relu_pool_fwd(aK, ab, aA, ks)[0]
    close_dot(dev.to_numpy(y), ry.astype(dtype), y_abs, Ci * KH * KW + 1, dtype)
    got_code = dev.to_numpy(code)
    if dtype == np.float64:
        exact(got_code, rcode)
    dy = rnd(rng, ry.shape, dtype)
    for want_dA in (False, True):
        dk, dbias, da = dev.conv_relu_pool_bwd(dK_, dA_, dev.sconcrete(dy), code, ks, want_bias=True, want_dA=want_dA)
        rdk, rdb, rda = oracle.conv_relu_pool_bwd(K, A, dy, got_code, ks, True, want_dA)
        # the routing by `code` is linear and sign-free, so the same adjoint on absolute values is sum |a||b|
        adk, adb, ada = oracle.conv_relu_pool_bwd(aK, aA, np.abs(dy), got_code, ks, True, want_dA)
        close_dot(dev.to_numpy(dk), rdk.astype(dtype), adk, N * H * W, dtype)
        close_dot(dev.to_numpy(dbias), rdb.astype(dtype), adb, N * H * W, dtype)
        if want_dA:
            close_dot(dev.to_numpy(da), rda.astype(dtype), ada, Co * KH * KW, dtype)
        else:
            assert da is None
    # no bias, no bias cotangent
    y2, code2 = dev.conv_relu_pool_fwd(dK_, None, dA_, ks)
    close_dot(dev.to_numpy(y2), oracle.conv_relu_pool_fwd(K, None, A, ks)[0].astype(dtype),
              oracle.conv_relu_pool_fwd(aK, None, aA, ks)[0], Ci * KH * KW, dtype)
    dk2, none_b, _ = dev.conv_relu_pool_bwd(dK_, dA_, dev.sconcrete(dy), code2, ks, want_bias=False, want_dA=False)
    assert none_b is None
    c2 = dev.to_numpy(code2)
    close_dot(dev.to_numpy(dk2), oracle.conv_relu_pool_bwd(K, A, dy, c2, ks, False, False)[0].astype(dtype),
              oracle.conv_relu_pool_bwd(aK, aA, np.abs(dy), c2, ks, False, False)[0], N * H * W, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1000, 1003, 3])     # 16-byte vectors only / vectors + scalar tail / scalar only
def test_optimizers(dev, oracle, dtype, n):
    rng = np.random.default_rng(18)
    p, g = rnd(rng, (n,), dtype), rnd(rng, (n,), dtype)
    m, v = np.zeros(n, dtype), np.zeros(n, dtype)
    dp, dm, dv = dev.sconcrete(p), dev.sconcrete(m), dev.sconcrete(v)
    for t in range(1, 4):
        g = rnd(rng, (n,), dtype)
        dev.adam_step(dp, dm, dv, dev.sconcrete(g), t)
        oracle.adam_step(p, m, v, g, t)
    close(dev.to_numpy(dp), p, dtype, scale=10)
    close(dev.to_numpy(dm), m, dtype, scale=10)
    close(dev.to_numpy(dv), v, dtype, scale=10)
    dev.sgd_step(dp, dev.sconcrete(g), 0.02)
    oracle.sgd_step(p, g, 0.02)
    close(dev.to_numpy(dp), p, dtype, scale=10)


def test_prod_fold(dev, oracle):
    # testListProd golden (TestAdaptorSimplified.hs:1043-1053)
    val, (g,) = ad.grad(dev, lambda T, t: T.prod_fold(t), [dev.sconcrete(np.array([1.0, 2.0, 3.0, 4.0]))])
    assert float(dev.kfromS(val)) == 24.0
    np.testing.assert_allclose(dev.to_numpy(g), [24, 12, 8, 6], atol=1e-10)
    rng = np.random.default_rng(19)
    for n in (1, 7, 2048, 2049, 10_000, 123_457):
        # the reference's LongProdBench distribution 0.55 + U[0,1) overflows past ~9e4;
        # exp(U[-a,a]) keeps the product finite
        x = np.exp(rng.uniform(-1e-3, 1e-3, n)) if n > 10_000 else 0.55 + rng.uniform(0, 1, n)
        pr, gr = dev.prod_fold_grad(dev.sconcrete(x), 1.5)
        rp, rg = oracle.prod_fold_grad(x, 1.5)
        tol = 4e-16 * (n + 10)
        np.testing.assert_allclose(dev.kfromS(pr), rp, rtol=tol)
        np.testing.assert_allclose(dev.to_numpy(gr), rg, rtol=tol)
    z = np.array([2.0, 0.0, 3.0, 5.0])     # exact zeros: no division anywhere
    _, gz = dev.prod_fold_grad(dev.sconcrete(z), 1.0)
    assert dev.to_numpy(gz).tolist() == [0.0, 30.0, 0.0, 0.0]


# ------------------------------------------------------------- whole model
@pytest.mark.parametrize("fused", [True, "nodes", False])
def test_mnist_cnn_gradient_vs_oracle(dev, oracle, fused):
    """MnistCnnShaped2 loss and all 8 gradient leaves: the device at every fusion level against the oracle's
    UNFUSED program (the reference's own formulation: gather chains, sdot1In, kargMax gathers), both through the
    same dual-number pipeline (cgrad)."""
    rng = np.random.Generator(np.random.Philox(42))
    batch, kh, kw, c_out, n_hidden = 6, 2, 2, 4, 8
    params = mnist_cnn.init_params(kh, kw, c_out, n_hidden)
    glyphs = rng.uniform(0, 1, (batch, 28, 28))
    labels = np.eye(10)[rng.integers(0, 10, batch)]

    def run(B, fused):
        ins = [B.sconcrete(p) for p in params]
        g, l = B.sconcrete(glyphs), B.sconcrete(labels)
        val, grads = ad.grad(B, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, g, l, list(ps), fused=fused), ins)
        return float(B.kfromS(val)), [B.to_numpy(x) for x in grads]
    dl, dg = run(dev, fused)
    ol, og = run(oracle, False)       # always the oracle's UNFUSED (vectorised, reference-formulation) program
    np.testing.assert_allclose(dl, ol, rtol=1e-12)
    for a, b in zip(dg, og):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("q", [np.float32, np.float64])
def test_mnist_fcnn_gradient_vs_oracle(dev, oracle, q):
    """MnistFcnnRanked2 (BASELINE config 1): loss and the 6 gradient leaves, with
    the reference's mixed precision (Double net, Float second hidden layer)."""
    from horde_ad_b200 import mnist_fcnn
    rng = np.random.Generator(np.random.Philox(11))
    datum, target = rng.uniform(0, 1, 784), np.eye(10)[7]
    params = mnist_fcnn.init_params(30, 10, q=q)

    def run(B):
        ins = [B.sconcrete(p) for p in params]
        d, t = B.sconcrete(datum), B.sconcrete(target)
        val, grads = ad.grad(B, lambda T, *ps: mnist_fcnn.afcnnMnistLoss2(T, d, t, list(ps)), ins)
        return float(B.kfromS(val)), [B.to_numpy(x) for x in grads]
    dl, dg = run(dev)
    ol, og = run(oracle)
    tol = 1e-12 if q == np.float64 else 2e-5
    np.testing.assert_allclose(dl, ol, rtol=tol * 10)
    for a, b in zip(dg, og):
        assert a.dtype == b.dtype
        np.testing.assert_allclose(a, b, rtol=tol * 100, atol=tol * 10)


@pytest.mark.parametrize("fused,batch,width", [(False, 9, 24), (True, 9, 24), ("pair", 9, 24), ("pair", 200, 128),
                                               ("wave", 9, 24), ("wave", 200, 128), ("wave", 64, 512)])
def test_mnist_rnn_gradient_vs_oracle(dev, oracle, fused, batch, width):
    """MnistRnnShaped2 (BASELINE config 5) at a small width: 28 unrolled steps of
    two tanh layers (smatmul2 on the DMMA GEMM), loss and the 8 gradient leaves;
    `fused`: the layer body as the tanh_affine node (`"pair"`: `wX x + wS s` as the smatmul2_sum node too:
    paired GEMM launches forward and in both transposition rules; `"wave"`: the wavefront schedule, layer 1 of row t
    and layer 2 of row t - 1 in one group launch), checked against the oracle's UNFUSED program."""
    from horde_ad_b200 import mnist_rnn
    rng = np.random.Generator(np.random.Philox(12))
    glyphs, labels = rng.uniform(0, 1, (batch, 28, 28)), np.eye(10)[rng.integers(0, 10, batch)]
    params = mnist_rnn.init_params(width, rng_range=0.2 if width < 100 else 0.05)

    def run(B, fused):
        ins = [B.sconcrete(p) for p in params]
        g, l = B.sconcrete(glyphs), B.sconcrete(labels)
        val, grads = ad.grad(B, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, g, l, list(ps), fused=fused), ins)
        return float(B.kfromS(val)), [B.to_numpy(x) for x in grads]
    dl, dg = run(dev, fused)
    ol, og = run(oracle, False)
    np.testing.assert_allclose(dl, ol, rtol=1e-12)
    for a, b in zip(dg, og):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(512, 1024), (7, 3), (1, 1), (33, 257), (5, 0)])
def test_tanh_affine_node(dev, dtype, shape):
    """hb_tanh_affine_fwd/bwd (the body of rnnMnistLayerS) against the oracle; the forward value is the
    same sequence of roundings as the unfused maps up to tanh itself."""
    from oracle import concrete as OC
    rng = np.random.default_rng(21)
    u, v, b, c = rnd(rng, shape, dtype), rnd(rng, shape, dtype), rnd(rng, shape[:1], dtype), rnd(rng, shape, dtype)
    y = dev.tanh_affine_fwd(dev.sconcrete(u), dev.sconcrete(v), dev.sconcrete(b))
    close(dev.to_numpy(y), OC.tanh_affine_fwd(u, v, b).astype(dtype), dtype, scale=4)
    y0 = dev.tanh_affine_fwd(dev.stranspose([1, 0], dev.sconcrete(u.T.copy())), dev.sconcrete(v), None)
    close(dev.to_numpy(y0), OC.tanh_affine_fwd(u, v, None).astype(dtype), dtype, scale=4)
    yh = dev.to_numpy(y)
    dz, db = dev.tanh_affine_bwd(y, dev.sconcrete(c))
    rdz, rdb = OC.tanh_affine_bwd(yh.astype(np.float64), c.astype(np.float64))
    close(dev.to_numpy(dz), rdz.astype(dtype), dtype, scale=4)
    close(dev.to_numpy(db), rdb.astype(dtype), dtype, scale=4 * max(1, shape[1]))
    dz2, none = dev.tanh_affine_bwd(y, dev.sconcrete(c), want_bias=False)
    assert none is None and np.array_equal(dev.to_numpy(dz2), dev.to_numpy(dz))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_matmul2_group_edge_cases(dev, oracle, dtype):
    """Empty reduced axes, empty results, broadcast (stride-0) operands, more than four products, an accumulator
    shared by two results, errors for inconsistent arguments."""
    rng = np.random.default_rng(37)
    f64 = lambda x: x.astype(np.float64)
    # K = 0 next to K = 5; an empty result
    for m, n, ks in [(4, 5, (0, 5)), (0, 5, (2, 2)), (3, 0, (2, 2))]:
        a = [rnd(rng, (m, k), dtype) for k in ks]
        b = [rnd(rng, (k, n), dtype) for k in ks]
        outs = dev.smatmul2_group([(None, None), (None, None)], [(dev.sconcrete(a[0]), dev.sconcrete(b[0]), 0), (dev.sconcrete(a[1]), dev.sconcrete(b[1]), 1)])
        for o, x, y in zip(outs, a, b):
            close(dev.to_numpy(o), (f64(x) @ f64(y)).astype(dtype), dtype, scale=16)
    # a broadcast zero state (the first time step of the recurrence) and a broadcast bias-like operand
    m, n, k = 64, 96, 32
    a, b = rnd(rng, (m, k), dtype), rnd(rng, (k, n), dtype)
    zero = dev.srepl([k, n], 0.0, dtype)
    ones = dev.srepl([k, n], 1.0, dtype)
    o0, o1 = dev.smatmul2_group([(None, None), (None, None)], [(dev.sconcrete(a), dev.sconcrete(b), 0), (dev.sconcrete(a), zero, 0),
                                                              (dev.sconcrete(a), ones, 1)])
    close(dev.to_numpy(o0), (f64(a) @ f64(b)).astype(dtype), dtype, scale=8 * k)
    close(dev.to_numpy(o1), (f64(a) @ np.ones((k, n))).astype(dtype), dtype, scale=8 * k)
    # six products into two results (more than one launch's worth)
    aa = [rnd(rng, (m, k), dtype) for _ in range(6)]
    bb = [rnd(rng, (k, n), dtype) for _ in range(6)]
    idx = [0, 1, 0, 1, 1, 0]
    outs = dev.smatmul2_group([(None, None), (None, None)], [(dev.sconcrete(x), dev.sconcrete(y), i) for x, y, i in zip(aa, bb, idx)])
    for r in range(2):
        ref = sum(f64(x) @ f64(y) for x, y, i in zip(aa, bb, idx) if i == r)
        close(dev.to_numpy(outs[r]), ref.astype(dtype), dtype, scale=8 * 6 * k)
    # the same accumulator array named by two results: both results start from its value, it is left alone
    acc = rnd(rng, (m, n), dtype)
    held = dev.sconcrete(acc)
    o0, o1 = dev.smatmul2_group([(held, None), (held, None)], [(dev.sconcrete(aa[0]), dev.sconcrete(bb[0]), 0), (dev.sconcrete(aa[1]), dev.sconcrete(bb[1]), 1)])
    exact(dev.to_numpy(held), acc)
    close(dev.to_numpy(o0), (f64(acc) + f64(aa[0]) @ f64(bb[0])).astype(dtype), dtype, scale=8 * k)
    close(dev.to_numpy(o1), (f64(acc) + f64(aa[1]) @ f64(bb[1])).astype(dtype), dtype, scale=8 * k)
    # errors: a result without a product, mismatched shapes, accumulator and epilogue together
    with pytest.raises(Exception):
        dev.smatmul2_group([(None, None), (None, None)], [(dev.sconcrete(aa[0]), dev.sconcrete(bb[0]), 0)])
    with pytest.raises(Exception):
        dev.smatmul2_group([(None, None)], [(dev.sconcrete(aa[0]), dev.sconcrete(bb[0]), 0), (dev.sconcrete(aa[0][:5]), dev.sconcrete(bb[0]), 0)])
    with pytest.raises(Exception):
        dev.smatmul2_group([(dev.sconcrete(acc), dev.sconcrete(acc[:, 0].copy()))], [(dev.sconcrete(aa[0]), dev.sconcrete(bb[0]), 0)])


def test_layer_body_tanh_is_within_ulps_of_the_correctly_rounded_value(dev):
    """hb_tanh_f64 (common.cuh: the branch-free tanh of hb_tanh_affine_fwd and of the grouped products' epilogue)
    against numpy's over the whole range, in ulps; special values."""
    rng = np.random.default_rng(41)
    parts = [rng.uniform(-1, 1, 200000), rng.uniform(-25, 25, 200000), rng.uniform(-0.6, 0.6, 100000),
             np.sign(rng.uniform(-1, 1, 100000)) * 10.0 ** rng.uniform(-300, 1.5, 100000),
             np.linspace(0.3, 0.4, 24001), (np.arange(1, 60) * np.log(2) / 2)[:, None].repeat(3, 1).ravel() * (1 + np.array([-1e-15, 0, 1e-15]).repeat(59).reshape(3, 59).T.ravel())]
    x = np.concatenate(parts)
    x = x[: x.size // 8 * 8].reshape(8, -1)
    got = dev.to_numpy(dev.tanh_affine_fwd(dev.sconcrete(x), None, None))
    ref = np.tanh(x)
    ulp = np.abs(got - ref) / np.spacing(np.abs(ref))
    assert float(ulp.max()) <= 4.0, float(ulp.max())
    assert np.all(np.abs(got) <= 1.0) and np.all(np.sign(got) == np.sign(x))
    sp = np.array([[0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 20.0, -20.0, 1e300, -1e300, 19.0]])
    g = dev.to_numpy(dev.tanh_affine_fwd(dev.sconcrete(sp), None, None))[0]
    assert g[0] == 0 and not np.signbit(g[0]) and g[1] == 0 and np.signbit(g[1])
    assert g[2] == 1 and g[3] == -1 and np.isnan(g[4]) and g[5] == 5e-324 and g[6] == -5e-324
    assert g[7] == 1 and g[8] == -1 and g[9] == 1 and g[10] == -1 and abs(g[11] - np.tanh(19.0)) <= 2 ** -52


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(512, 28, 512, 1024), (512, 512, 512, 1024), (64, 40, 24, 130), (3, 2, 5, 4)])
def test_matmul2_sum_tanh_is_the_unfused_pair_bit_for_bit(dev, oracle, dtype, shape):
    """hb_matmul2_sum_tanh (bias + tanh in the epilogue of the paired product's last K slice) == hb_matmul2_sum
    followed by hb_tanh_affine_fwd, bit for bit (same sums, same tanh); and both against the oracle."""
    m, k1, k2, n = shape
    rng = np.random.default_rng(23)
    a1, b1 = rnd(rng, (m, k1), dtype), rnd(rng, (k1, n), dtype)
    a2, b2 = rnd(rng, (m, k2), dtype), rnd(rng, (k2, n), dtype)
    bias = rnd(rng, (m,), dtype)
    d = [dev.sconcrete(x) for x in (a1, b1, a2, b2, bias)]
    fused = dev.to_numpy(dev.smatmul2_sum_tanh(*d))
    two = dev.to_numpy(dev.tanh_affine_fwd(dev.smatmul2_sum(*d[:4]), None, d[4]))
    exact(fused, two)
    ref = oracle.smatmul2_sum_tanh(*[x.astype(np.float64) for x in (a1, b1, a2, b2, bias)])
    close(fused, ref.astype(dtype), dtype, scale=8 * (k1 + k2))
    # operands as transposed views (the orientations the RNN's backward pass hands over)
    dt = [dev.stranspose([1, 0], dev.sconcrete(np.ascontiguousarray(x.T))) for x in (a1, b1, a2, b2)]
    exact(dev.to_numpy(dev.smatmul2_sum_tanh(*dt, d[4])), dev.to_numpy(dev.tanh_affine_fwd(dev.smatmul2_sum(*dt), None, d[4])))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(512, 1024), (33, 257), (7, 1), (64, 130)])
def test_tanh_affine_bwd_acc_is_the_adjoint_plus_add(dev, oracle, dtype, shape):
    """hb_tanh_affine_bwd_acc (bias cotangent accumulated by the pass that produces the row sums) == hb_tanh_affine_bwd
    followed by the addition, bit for bit; in place exactly when the accumulator is uniquely referenced."""
    rng = np.random.default_rng(29)
    y, c, acc = np.tanh(rnd(rng, shape, dtype)), rnd(rng, shape, dtype), rnd(rng, shape[:1], dtype)
    dy, dc = dev.sconcrete(y), dev.sconcrete(c)
    dz0, db0 = dev.tanh_affine_bwd(dy, dc)
    want = dev.to_numpy(dev.binary("add", dev.sconcrete(acc), db0))
    oz, ob = oracle.tanh_affine_bwd_acc(y.astype(np.float64), c.astype(np.float64), acc.astype(np.float64))
    close(dev.to_numpy(dz0), oz.astype(dtype), dtype)
    close(want, ob.astype(dtype), dtype, scale=shape[1])
    # uniquely referenced accumulator: updated in place, the same handle comes back
    a1 = dev.sconcrete(acc)
    dz1, out1 = dev.tanh_affine_bwd_acc(dy, dc, a1)
    assert out1 is a1
    exact(dev.to_numpy(dz1), dev.to_numpy(dz0))
    exact(dev.to_numpy(out1), want)
    # an accumulator somebody else holds a view of is left alone
    a2 = dev.sconcrete(acc)
    view = dev.sreverse(a2)
    dz2, out2 = dev.tanh_affine_bwd_acc(dy, dc, a2)
    assert out2 is not a2
    exact(dev.to_numpy(a2), acc)
    exact(dev.to_numpy(view), acc[::-1])
    exact(dev.to_numpy(out2), want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("m,n,ks", [(512, 1024, (512, 512, 512)), (130, 70, (33, 64, 5)), (64, 64, (28, 512, 512, 512)),
                                    (8, 3, (2, 1))])
def test_matmul2_group_against_the_oracle(dev, oracle, dtype, m, n, ks):
    """hb_matmul2_group: products of one [M, N] into named results in one launch -- accumulators (in place when
    uniquely referenced), tanh epilogues, operands in either orientation -- against the rules applied one by one."""
    rng = np.random.default_rng(31)
    a = [rnd(rng, (m, k), dtype) for k in ks]
    b = [rnd(rng, (k, n), dtype) for k in ks]
    acc, bias = rnd(rng, (m, n), dtype), rnd(rng, (m,), dtype)
    f64 = lambda x: x.astype(np.float64)
    da, db = [dev.sconcrete(x) for x in a], [dev.sconcrete(x) for x in b]
    # weights as transposed views (the backward pass hands over wS^T), cotangents as they are
    dat = [dev.stranspose([1, 0], dev.sconcrete(np.ascontiguousarray(x.T))) for x in a]
    scale = 8 * sum(ks)
    for ops in (da, dat):
        # (1) the backward step: the first two products into result 0, the rest into result 1; result 1 accumulated
        idx = [0, 0] + [1] * (len(ks) - 2)
        prods = [(ops[g], db[g], idx[g]) for g in range(len(ks))]
        oprods = [(f64(a[g]), f64(b[g]), idx[g]) for g in range(len(ks))]
        nres = 2 if len(ks) > 2 else 1
        held = dev.sconcrete(acc)
        res = [(None, None), (held, None)][:nres]
        ores = [(None, None), (f64(acc), None)][:nres]
        outs = dev.smatmul2_group(res, prods)
        del res
        refs = oracle.smatmul2_group(ores, oprods)
        for o, r in zip(outs, refs):
            close(dev.to_numpy(o), r.astype(dtype), dtype, scale=scale)
        if nres == 2 and dtype == np.float64:
            assert outs[1] is held                      # updated in place
        # (2) the forward wavefront: two layers' tanh (wX x + wS s + b), every result with its epilogue
        if len(ks) >= 3:
            idx = [0, 0] + [1] * (len(ks) - 2)
            b2 = dev.sconcrete(bias[::-1].copy())
            outs = dev.smatmul2_group([(None, dev.sconcrete(bias)), (None, b2)], [(ops[g], db[g], idx[g]) for g in range(len(ks))])
            refs = oracle.smatmul2_group([(None, f64(bias)), (None, f64(bias[::-1]))], oprods)
            for o, r in zip(outs, refs):
                close(dev.to_numpy(o), r.astype(dtype), dtype, scale=scale)
            # the first result is exactly hb_matmul2_sum_tanh of its two products (same kernel, same chain)
            if dtype == np.float64 and ks[0] >= 32:
                one = dev.smatmul2_sum_tanh(ops[0], db[0], ops[1], db[1], dev.sconcrete(bias))
                close(dev.to_numpy(outs[0]), dev.to_numpy(one), dtype, scale=scale)
    # an accumulator somebody else holds is left alone
    held = dev.sconcrete(acc)
    view = dev.stranspose([1, 0], held)
    out, = dev.smatmul2_group([(held, None)], [(da[0], db[0], 0), (da[1], db[1], 0)])
    assert out is not held
    exact(dev.to_numpy(view), acc.T)
    close(dev.to_numpy(out), (f64(acc) + f64(a[0]) @ f64(b[0]) + f64(a[1]) @ f64(b[1])).astype(dtype), dtype, scale=scale)
    # deterministic: the same launch twice (K slices of a tile are combined in chain order)
    one, = dev.smatmul2_group([(dev.sconcrete(acc), None)], [(da[0], db[0], 0), (da[1], db[1], 0)])
    two, = dev.smatmul2_group([(dev.sconcrete(acc), None)], [(da[0], db[0], 0), (da[1], db[1], 0)])
    exact(dev.to_numpy(one), dev.to_numpy(two))
    close(dev.to_numpy(one), dev.to_numpy(out), dtype, scale=scale)


@pytest.mark.parametrize("name,mk,expected", P.VSTACK_GOLDEN, ids=[v[0] for v in P.VSTACK_GOLDEN])
def test_vstack_known_answers_on_the_device(dev, name, mk, expected):
    """vstackABC / vstackBuild (TestAdaptorSimplified.hs:785-835, 949) through the C ABI, and the
    reference's 1e6-element agreement test (testVstackConcatConcrete, :850-853) for the slice/append form."""
    a, b, c = (dev.sconcrete(x) for x in mk(10))
    assert dev.to_numpy(P.vstackABC(dev, a, b, c)).tolist() == expected
    assert dev.to_numpy(P.vstackBuild(dev, a, b, c)).tolist() == expected
    n = 10**6
    ha, hb, hc = mk(n)
    got = dev.to_numpy(P.vstackABC(dev, dev.sconcrete(ha), dev.sconcrete(hb), dev.sconcrete(hc)))
    ref = np.concatenate([[ha[0] + hb[1]], ha[1:n - 1] + hb[2:n] + hc[:n - 2], [ha[n - 1] + hc[n - 2]]])
    exact(got, ref)
    f32 = dev.to_numpy(P.vstackABC(dev, *(dev.sconcrete(x.astype(np.float32)) for x in (ha, hb, hc))))
    exact(f32.astype(np.float64), ref)         # "trustedResult": the Float run, cast, equals the Double run


@pytest.mark.parametrize("name,prog,expected", P.GATHER_BUILD_GOLDEN, ids=[g[0] for g in P.GATHER_BUILD_GOLDEN])
def test_gather_scatter_build_variants_on_the_device(dev, name, prog, expected):
    """The rbuild1 variants of the reference's gather / scatter tests (TestGatherSimplified.hs:161-398,
    968-1232) through the C ABI."""
    t = np.tile(np.array([0.0, 1.0]), (7, 1))
    _, (g,) = ad.grad(dev, prog, [dev.sconcrete(t)])
    assert dev.to_numpy(g).reshape(-1).tolist() == expected


GATHER_COND_CASES = [(P.gatherCond, (7, 2), P.GATHER_COND_GOLDEN), (P.gatherCond2, (7, 2), P.GATHER_COND_GOLDEN),
                     (P.gatherCond3, (7, 2), P.GATHER_COND34_GOLDEN), (P.gatherCond4, (7, 2), P.GATHER_COND34_GOLDEN),
                     (P.gatherCond5, (2, 4, 2), P.GATHER_COND56_GOLDEN), (P.gatherCond6, (2, 4, 2), P.GATHER_COND56_GOLDEN)]


@pytest.mark.parametrize("prog,shape,expected", GATHER_COND_CASES, ids=[c[0].__name__ for c in GATHER_COND_CASES])
def test_gather_with_conditional_index_on_the_device(dev, oracle, prog, shape, expected):
    """testGatherCond .. testGatherCond6 and their rbuild1 variants (TestGatherSimplified.hs:716-915):
    `ifH (i ==. 3) 0 j` compiled into the index program, index components that run out of bounds;
    values bit-exact vs the oracle, gradients = the reference's literals."""
    t = np.broadcast_to(np.array([0.0, 1.0]), shape).copy()
    data = np.arange(t.size, dtype=np.float64).reshape(shape) * 1.5 - 3
    exact(dev.to_numpy(prog(dev, dev.sconcrete(data))), np.ascontiguousarray(prog(oracle, data)))
    _, (g,) = ad.grad(dev, prog, [dev.sconcrete(t)])
    assert dev.to_numpy(g).reshape(-1).tolist() == expected
    build = lambda T, x: nn.sbuild1(T, 4, lambda i: prog(T, T.binary("mul", x, T.srepl(list(shape), float(i)))))
    _, (g,) = ad.grad(dev, build, [dev.sconcrete(t)])
    assert dev.to_numpy(g).reshape(-1).tolist() == [6 * x for x in expected]


def test_gather_transpose33_on_the_device(dev):
    """testGatherTranspose33 / Build33 (TestGatherSimplified.hs:564-606) through the C ABI: rank-10 views,
    permutations shorter than the rank, reshapes of non-contiguous views, the DMMA GEMM."""
    for f, expected in ((P.gatherTranspose33, P.GT33_GOLDEN), (P.gatherTransposeBuild33, P.GT33_BUILD_GOLDEN)):
        _, (g,) = ad.grad(dev, f, [dev.sconcrete(P.T128)])
        np.testing.assert_allclose(dev.to_numpy(g).reshape(-1), np.array(expected), atol=1e-10, rtol=0)


def test_cnn_opp3_on_the_device(dev):
    """testCNNOPP3 / testCNNOPP3b (TestGatherSimplified.hs:1365-1391) through the C ABI: out-of-bounds
    index -> 0, stacking of 0-d results, host conditionals; known folded value and gradient."""
    g = np.full((3, 3, 3, 3), 7.0)
    prog = lambda T, u: P.maxPool2dUnpadded3(T, P.conv2dPreserving3(T, u))
    v, (gr,) = ad.grad(dev, prog, [dev.sconcrete(g)])
    assert dev.to_numpy(v).reshape(-1).tolist() == P.CNNOPP3_GOLDEN
    expect = np.zeros((3, 3, 3, 3))
    expect[0, 0, 0, 1] = expect[0, 1, 1, 1] = 2.0
    exact(dev.to_numpy(gr), expect)
    g2 = dev.sconcrete(np.full((2, 2, 2, 2), 7.0))                       # testCNNOPP6 / testCNNOPP7
    assert dev.to_numpy(P.maxPool2dUnpadded3(dev, P.conv2dPreserving3z(dev, g2))).reshape(-1).tolist() == P.CNNOPP6_GOLDEN
    assert dev.to_numpy(P.maxPool2dUnpadded3y(dev, P.conv2dPreserving3y(dev, g2))).reshape(-1).tolist() == P.CNNOPP7_GOLDEN


def test_folds_builds_and_cond(dev, oracle):
    """tmapAccumL / tfold / tscan as host loops over device slices, tsbuild1,
    tsmap0N, scond (non-vectorised code paths: they must work, not be fast)."""
    rng = np.random.default_rng(23)
    es, a0 = rng.uniform(-1, 1, (6, 3, 4)), rng.uniform(-1, 1, (3, 4))

    def prog(T, acc0, xs):
        step = lambda a, e: (T.binary("add", T.unary("sin", a), T.binary("mul", a, e)), T.unary("cos", a))
        acc, bs = nn.smapAccumL(T, step, acc0, xs)
        sc = nn.sscan(T, lambda a, e: T.binary("mul", a, e), acc0, xs)
        return T.binary("add", T.ssum0(acc), T.binary("add", T.ssum0(bs), T.ssum0(sc)))

    def run(B):
        val, grads = ad.grad(B, prog, [B.sconcrete(a0), B.sconcrete(es)])
        return float(B.kfromS(val)), [B.to_numpy(g) for g in grads]
    dv, dg = run(dev)
    ov, og = run(oracle)
    np.testing.assert_allclose(dv, ov, rtol=1e-12)
    for a, b in zip(dg, og):
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-14)
    # fold golden on the device: testSin0Fold2 (TestRevFwdFold.hs:545-550)
    _, (g,) = ad.grad(dev, lambda T, x: nn.sfold(T, lambda a, _e: T.unary("sin", a), x, T.srepl([5], 42.0)), [dev.sscalar(1.1)])
    np.testing.assert_allclose(float(dev.kfromS(g)), 0.12389721944941383, atol=1e-10)
    # a fold whose step is a fold: testSin0FoldNestedS1 (TestRevFwdFold.hs:2316-2326)
    def nested(T, x0):
        inner = lambda x, a: nn.sfold(T, lambda x2, a2: T.binary("mul", T.binary("mul", T.sscalar(0.7), x2), a2),
                                      a, T.sreplicate(7, x))
        return nn.sfold(T, inner, x0, T.sreplicate(3, x0))
    _, (g,) = ad.grad(dev, nested, [dev.sscalar(1.1)])
    np.testing.assert_allclose(float(dev.kfromS(g)), 2.0504979297616553e-43, rtol=1e-12)
    t = dev.sconcrete(a0)
    b1 = nn.sbuild1(dev, 4, lambda i: dev.binary("mul", t, dev.srepl([3, 4], float(i))))
    exact(dev.to_numpy(b1), np.stack([a0 * i for i in range(4)]))
    m0 = nn.smap0N(dev, lambda x: dev.unary("negate", x), t)
    exact(dev.to_numpy(m0), -a0)
    assert nn.scond(dev, dev.all_leq(t, t), t, m0) is t
    z = nn.szipWith0N(dev, lambda x, y: dev.binary("mul", x, y), t, m0)
    exact(dev.to_numpy(z), a0 * -a0)
    bld = nn.sbuild(dev, [2, 3], lambda ix: dev.srepl([4], float(10 * ix[0] + ix[1])))
    exact(dev.to_numpy(bld), np.array([[[10.0 * i + j] * 4 for j in range(3)] for i in range(2)]))
    lin = nn.stoListLinear(dev, t)
    assert len(lin) == 12 and dev.kfromS(lin[5]) == a0.reshape(-1)[5]
    exact(dev.to_numpy(nn.sfromVectorLinear(dev, [3, 4], lin)), a0)
    assert dev.kfromS(nn.sminimum(dev, t)) == a0.min() and dev.kfromS(nn.smaximum(dev, t)) == a0.max()


@pytest.mark.parametrize("fused", [True, "nodes", False])
def test_forward_reverse_duality_on_device(dev, fused):
    """<cgrad f, ds> = cjvp f ds on the device (TestConvQuickCheck.hs:182-268): a
    size-independent check that needs no oracle; the forward mode exercises the
    gather-at-argmax index programs, the reverse mode their scatters."""
    rng = np.random.Generator(np.random.Philox(14))
    batch = 64 if fused is not False else 8
    hyper = (4, 4, 16, 64) if fused is not False else (2, 2, 4, 8)
    glyphs, labels = dev.sconcrete(rng.uniform(0, 1, (batch, 28, 28))), dev.sconcrete(np.eye(10)[rng.integers(0, 10, batch)])
    host = mnist_cnn.init_params(*hyper)
    params = [dev.sconcrete(p) for p in host]
    ds_host = [rng.uniform(-1, 1, p.shape) for p in host]
    ds = [dev.sconcrete(d) for d in ds_host]
    f = lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, glyphs, labels, list(ps), fused=fused)
    _, grads = ad.grad(dev, f, params)
    _, tan = ad.jvp(dev, f, params, ds)
    lhs = sum(float(np.sum(dev.to_numpy(g) * d)) for g, d in zip(grads, ds_host))
    np.testing.assert_allclose(float(dev.kfromS(tan)), lhs, rtol=1e-9)


def test_graph_capture_replay(dev):
    rng = np.random.default_rng(20)
    x = dev.sconcrete(rng.uniform(-.5, .5, (64, 32)))
    w = dev.sconcrete(rng.uniform(-.5, .5, (32, 16)))
    import ctypes as C
    def step():
        return dev.ssum0(dev.unary("tanh", dev.smatmul2(x, w)))
    eager = dev.kfromS(step())
    ffi.call("hb_graph_begin")
    out = step()
    g = C.c_void_p()
    ffi.call("hb_graph_end", C.byref(g))
    for _ in range(3):
        ffi.call("hb_graph_launch", g)
    assert dev.kfromS(out) == eager
    ffi.call("hb_graph_free", g)


@pytest.mark.parametrize("name,f,inputs,expected", P.ADAPTOR_REV_GOLDEN, ids=[g[0] for g in P.ADAPTOR_REV_GOLDEN])
def test_relu_and_bar_relu_known_answers_on_the_device(dev, name, f, inputs, expected):
    """The reference's literal gradients (TestAdaptorSimplified.hs:1247-1251, 1300-1345, 1404-1462, 1850-1895,
    eps 1e-10) reproduced by the CUDA path: relu mask, per-element maxH through tsmap0N, atan2/sin chains."""
    _, grads = ad.grad(dev, lambda T, *xs: T.ssum0(f(T, *xs)), [dev.sconcrete(np.asarray(x, np.float64)) for x in inputs])
    for g, e in zip(grads, expected):
        np.testing.assert_allclose(dev.to_numpy(g), e, atol=1e-10, rtol=0)


def test_bar_relu_vjp_and_forward_known_answers_on_the_device(dev):
    """testBarReluDt / testBarReluMaxDt (vjp, cotangent 42.2) and testBarReluMax3CFwd / FwdS (cjvp / jvp),
    TestAdaptorSimplified.hs:1843-1848, 1872-1877, 1896-1925."""
    for name, f, x, dt, expected in P.ADAPTOR_VJP_GOLDEN:
        _, (g,) = ad.grad(dev, f, [dev.sconcrete(x)], dt=dev.sconcrete(dt))
        np.testing.assert_allclose(dev.to_numpy(g), expected, atol=1e-10, rtol=1e-11, err_msg=name)
    for name, f, x, dx, expected in P.ADAPTOR_FWD_GOLDEN:
        _, t = ad.jvp(dev, f, [dev.sconcrete(x)], [dev.sconcrete(dx)])
        np.testing.assert_allclose(dev.to_numpy(t), expected, atol=1e-10, rtol=1e-11, err_msg=name)


@pytest.mark.parametrize("name,f,x,expected", P.SCAN_GOLDEN, ids=[g[0] for g in P.SCAN_GOLDEN])
def test_scan_known_answers_on_the_device(dev, name, f, x, expected):
    """testSin0Scan1/2/3/4/6 (TestRevFwdFold.hs:934-1000) reproduced by the CUDA path, eps 1e-10."""
    _, (g,) = ad.grad(dev, lambda T, t: T.ssum0(f(T, t)), [dev.sconcrete(x)])
    np.testing.assert_allclose(dev.to_numpy(g), expected, atol=1e-10, rtol=0)


"""The device contraction pass (hb_lazy_*, csrc/lazy.cu + lazy_rewrite.cu) on the device:
the UNCHANGED vectorised program -- chained i+j gathers, sdot1In over broadcast views, [0.0,1.0]-table relu masks,
kargMax gathers and their adjoint scatters, exactly what horde-ad's interpreter sends -- recorded through the C ABI,
rewritten, replayed, and compared with (a) the same program run eagerly op by op, (b) the hand-fused program and
(c) the CPU oracle's UNFUSED program."""
import ctypes as C

import numpy as np
import pytest

from horde_ad_b200 import ad, ffi, mnist_cnn, nn

pytestmark = pytest.mark.gpu


def _cnn_inputs(batch, hyper, seed=42):
    rng = np.random.Generator(np.random.Philox(seed))
    params = mnist_cnn.init_params(*hyper)
    glyphs = rng.uniform(0, 1, (batch, 28, 28))
    labels = np.eye(10)[rng.integers(0, 10, batch)]
    return params, glyphs, labels


def _bucket(B, params, glyphs, labels, fused):
    loss, grads = ad.grad(B, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, glyphs, labels, list(ps), fused=fused), params)
    return B.flatten_concat(list(grads) + [loss]) if hasattr(B, "flatten_concat") else \
        np.concatenate([np.asarray(g).reshape(-1) for g in grads] + [np.asarray(loss).reshape(1)])


def test_replay_without_the_pass_is_the_eager_program_bit_for_bit(dev):
    params, glyphs, labels = _cnn_inputs(5, (2, 2, 3, 6))
    dp = [dev.sconcrete(p) for p in params]
    g, l = dev.sconcrete(glyphs), dev.sconcrete(labels)
    eager = dev.to_numpy(_bucket(dev, dp, g, l, False))
    prog = dev.record(lambda: _bucket(dev, dp, g, l, False), rewrite=False)
    rec, live = prog.stats
    assert rec >= live > 40
    prog.run()
    got = dev.to_numpy(prog.outs[0])
    assert np.array_equal(got, eager)
    prog.run()                                   # replays rebind their outputs
    assert np.array_equal(dev.to_numpy(prog.outs[0]), eager)


@pytest.mark.parametrize("fused,batch,width", [("wave", 40, 64), ("pair", 40, 64), ("wave", 7, 24)])
def test_recorded_rnn_gradient_replays_to_the_eager_values(dev, oracle, fused, batch, width):
    """MnistRnnShaped2's gradient with the grouped products (hb_matmul2_group), the list products (hb_matmul2_list)
    and the accumulating tanh adjoint (hb_tanh_affine_bwd_acc) RECORDED through the C ABI and replayed: every one of
    those entry points has a lazy form.  Replayed accumulators are never unique (the recorder holds them), so the
    replay takes the value-semantics paths: equal to the eager run to rounding, and to the oracle's unfused program."""
    from horde_ad_b200 import mnist_rnn
    rng = np.random.Generator(np.random.Philox(17))
    glyphs, labels = rng.uniform(0, 1, (batch, 28, 28)), np.eye(10)[rng.integers(0, 10, batch)]
    params = mnist_rnn.init_params(width, rng_range=0.1)

    def bucket(B, fused):
        ins = [B.sconcrete(p) for p in params]
        g, l = B.sconcrete(glyphs), B.sconcrete(labels)
        val, grads = ad.grad(B, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, g, l, list(ps), fused=fused), ins)
        return B.flatten_concat(list(grads) + [val])
    eager = dev.to_numpy(bucket(dev, fused))
    prog = dev.record(lambda: bucket(dev, fused))
    prog.run()
    got = dev.to_numpy(prog.outs[0])
    np.testing.assert_allclose(got, eager, rtol=1e-10, atol=1e-14)
    prog.run()
    assert np.array_equal(dev.to_numpy(prog.outs[0]), got)          # deterministic
    ov, og = ad.grad(oracle, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=False), params)
    ref = np.concatenate([np.asarray(x).reshape(-1) for x in og] + [np.asarray(ov).reshape(1)])
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("batch,hyper", [(6, (2, 2, 4, 8)), (33, (4, 4, 16, 16)), (8, (2, 3, 5, 7))])
def test_contraction_pass_on_the_vectorised_cnn_gradient(dev, oracle, batch, hyper):
    params, glyphs, labels = _cnn_inputs(batch, hyper)
    dp = [dev.sconcrete(p) for p in params]
    g, l = dev.sconcrete(glyphs), dev.sconcrete(labels)
    prog = dev.record(lambda: _bucket(dev, dp, g, l, False))
    text = prog.dump()
    rec, live = prog.stats
    # both conv layers and both adjoints were recognised, whole-layer nodes emitted, gathers / scatters gone
    assert text.count("conv_relu_pool_fwd f64") == 2 and text.count("conv_relu_pool_bwd f64") == 2, text
    body = text.split("--")[0]
    names = [ln.split()[1] for ln in body.strip().splitlines()]
    assert names.count("gather") <= 1, text        # only the dense layer's relu mask (reused by its adjoint)
    assert "scatter_add" not in names and "dot_inner" not in names and "dot_inner_sum_leading" not in names, text
    assert live <= 24 < rec
    prog.run()
    got = dev.to_numpy(prog.outs[0])
    eager_unfused = dev.to_numpy(_bucket(dev, dp, g, l, False))        # the recorded program, run op by op
    fused = dev.to_numpy(_bucket(dev, dp, g, l, True))                # the hand-fused model of round 1
    ref = _bucket(oracle, [oracle.sconcrete(p) for p in params], glyphs, labels, False)   # oracle, UNFUSED
    scale = np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))
    assert np.max(np.abs(got - ref) / scale) < 1e-10
    assert np.max(np.abs(got - eager_unfused) / scale) < 1e-10
    # same kernels as the hand-fused program: bit-identical
    assert np.array_equal(got, fused)
    # data-movement parts of the recorded program are exact: max-pool routing picks the same elements
    nz_ref, nz_got = ref != 0, got != 0
    assert np.array_equal(nz_ref, nz_got)


def test_pass_without_layer_fusion_keeps_conv_and_pool_nodes_apart(dev, oracle):
    params, glyphs, labels = _cnn_inputs(7, (2, 2, 4, 8))
    dp = [dev.sconcrete(p) for p in params]
    g, l = dev.sconcrete(glyphs), dev.sconcrete(labels)
    prog = dev.record(lambda: _bucket(dev, dp, g, l, False), fuse_layers=False)
    body = prog.dump().split("--")[0]
    for name in ("conv2d_preserving f64", "conv2d_dkrn f64", "conv2d_dinp f64", "relu_pool_fwd f64", "relu_pool_bwd f64"):
        assert name in body, (name, body)
    prog.run()
    ref = _bucket(oracle, [oracle.sconcrete(p) for p in params], glyphs, labels, False)
    got = dev.to_numpy(prog.outs[0])
    scale = np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))
    assert np.max(np.abs(got - ref) / scale) < 1e-10


def test_conv_vjp_patterns_alone(dev, oracle):
    """ConvVjpBench's three programs (bench/ConvVjpBench.hs:428-555): the vectorised forward, and the dual-number
    adjoints w.r.t. kernel and input, each recorded on its own and rewritten to one conv node."""
    rng = np.random.default_rng(3)
    N, Ci, Co, H, W, KH, KW = 3, 5, 6, 9, 11, 3, 2
    K, A, dB = rng.uniform(-.5, .5, (Co, Ci, KH, KW)), rng.uniform(-.5, .5, (N, Ci, H, W)), rng.uniform(-.5, .5, (N, Co, H, W))
    dK_, dA_, dB_ = dev.sconcrete(K), dev.sconcrete(A), dev.sconcrete(dB)
    fwd = dev.record(lambda: nn.conv2dPreservingS_vectorised(dev, dK_, dA_))
    assert "conv2d_preserving f64" in fwd.dump() and fwd.stats[1] == 1, fwd.dump()
    fwd.run()
    np.testing.assert_allclose(dev.to_numpy(fwd.outs[0]), oracle.conv2dPreserving(K, A), rtol=1e-12, atol=1e-13)

    def vjp():
        _, (gk, ga) = ad.grad(dev, lambda T, k, a: nn.conv2dPreservingS_vectorised(T, k, a), [dK_, dA_], dt=dB_)
        return [gk, ga]
    bwd = dev.record(vjp)
    text = bwd.dump()
    assert "conv2d_dkrn f64" in text and "conv2d_dinp f64" in text, text
    bwd.run()
    np.testing.assert_allclose(dev.to_numpy(bwd.outs[0]), oracle.conv2d_dkrn(A, dB, KH, KW), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(dev.to_numpy(bwd.outs[1]), oracle.conv2d_dinp(K, dB), rtol=1e-12, atol=1e-13)


def test_unmatched_programs_are_left_alone(dev, oracle):
    """Patterns that only LOOK like the fused forms must not be rewritten: a stride-3 window over a 2-wide pool,
    a gather with an offset, a `valid` (non-preserving) window."""
    rng = np.random.default_rng(8)
    x = rng.uniform(-1, 1, (4, 3, 12, 12))
    dx = dev.sconcrete(x)
    # windows at stride 3 with 2x2 extent: not maxPool k k
    th = dev.sconcrete((3 * np.arange(4)[:, None] + np.arange(2)[None, :]).astype(np.int64))

    def odd_pool():
        xt = dev.stranspose([2, 3, 0, 1], nn.reluS_vectorised(dev, dx))
        win = dev.sgather([4, 4, 2, 2], xt, lambda ix: [dev.ix_tbl(th, [ix[0], ix[2]]), dev.ix_tbl(th, [ix[1], ix[3]])], 2)
        u = dev.sreshape([4, 3, 4, 4, 4], dev.stranspose([4, 5, 0, 1, 2, 3], win))
        return dev.sgather([4, 3, 4, 4], u, lambda ix: [ix[0], ix[1], ix[2], ix[3], dev.ix_argmax(u, ix)], 5)
    p = dev.record(odd_pool)
    assert "relu_pool_fwd" not in p.dump()
    p.run()
    got = dev.to_numpy(p.outs[0])
    r = np.maximum(x, 0)
    ref = np.stack([[r[:, :, 3 * i:3 * i + 2, 3 * j:3 * j + 2].max(axis=(2, 3)) for j in range(4)] for i in range(4)])
    np.testing.assert_array_equal(got, np.transpose(ref, (2, 3, 0, 1)))
    # shifted window i + j + 1: not conv2dPreservingS
    K = dev.sconcrete(rng.uniform(-1, 1, (2, 3, 2, 2)))

    def shifted():
        g1 = dev.sgather([12, 2], dev.stranspose([2, 0, 1, 3], dx), lambda ix: [ix[0] + ix[1] + 1], 1)
        g2 = dev.sgather([12, 2], dev.stranspose([4, 0, 1, 2, 3], g1), lambda ix: [ix[0] + ix[1]], 1)
        pt = dev.stranspose([1, 2, 3, 0], dev.sreplicate(2, dev.stranspose([5, 3, 4, 2, 0, 1], g2)))
        k = dev.stranspose([3, 4, 0, 5, 1, 2], dev.sreplicateN([4, 12, 12], dev.stranspose([1, 2, 0, 3], K)))
        return dev.ssumN(dev.sdot1In(pt, k), 2)
    q = dev.record(shifted)
    assert "conv2d" not in q.dump().split("--")[0]
    q.run()
    eager = dev.to_numpy(shifted())
    assert np.array_equal(dev.to_numpy(q.outs[0]), eager)


def test_lazy_program_inside_a_cuda_graph_follows_new_inputs(dev, oracle):
    params, glyphs, labels = _cnn_inputs(6, (2, 2, 4, 8))
    dp = [dev.sconcrete(p) for p in params]
    g, l = dev.sconcrete(glyphs), dev.sconcrete(labels)
    prog = dev.record(lambda: _bucket(dev, dp, g, l, False))
    prog.run()
    first = dev.to_numpy(prog.outs[0])
    dev.sync()
    ffi.call("hb_graph_begin")
    try:
        prog.run()
    finally:
        gr = C.c_void_p()
        ffi.call("hb_graph_end", C.byref(gr))
    ffi.call("hb_graph_launch", gr)
    assert np.array_equal(dev.to_numpy(prog.outs[0]), first)
    g2 = np.ascontiguousarray(glyphs[::-1])
    dev.upload_into(g, g2.ctypes.data_as(C.c_void_p))
    dev.sync()
    ffi.call("hb_graph_launch", gr)
    second = dev.to_numpy(prog.outs[0])
    ref = _bucket(oracle, [oracle.sconcrete(p) for p in params], g2, labels, False)
    scale = np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))
    assert np.max(np.abs(second - ref) / scale) < 1e-10 and not np.array_equal(first, second)
    ffi.call("hb_graph_free", gr)


def test_values_are_refused_while_recording(dev):
    x = dev.sconcrete(np.arange(6.0))
    ffi.call("hb_lazy_begin")
    try:
        y = dev.binary("add", x, x)
        with pytest.raises(ffi.HordeError, match="lazy value"):
            dev.to_numpy(y)
        with pytest.raises(ffi.HordeError, match="cannot be recorded"):
            dev.all_eq(x, x)
    finally:
        ffi.call("hb_lazy_abort")
    assert dev.to_numpy(dev.binary("add", x, x)).tolist() == [0, 2, 4, 6, 8, 10]


def test_sfold_product_is_recognised_from_its_step_function(dev, oracle):
    """`sfold (*) 1 x` (LongProdBench, bench/common/BenchProdTools.hs:154-182): the step closure is reified by
    recording it once (nn._step_is_product) and the fold runs as the product-scan kernel; the oracle runs the
    reference's n sequential host iterations.  Other steps are NOT mistaken for a product."""
    rng = np.random.default_rng(5)
    x = np.exp(rng.uniform(-1e-2, 1e-2, 3000))
    for step in (lambda T: (lambda a, e: T.binary("mul", a, e)), lambda T: (lambda a, e: T.binary("mul", e, a))):
        n0 = C.c_uint64()
        ffi.call("hb_launch_count", C.byref(n0))
        dv, (dg,) = ad.grad(dev, lambda T, t: nn.sfold(T, step(T), T.sscalar(1.0), t), [dev.sconcrete(x)])
        n1 = C.c_uint64()
        ffi.call("hb_launch_count", C.byref(n1))
        assert n1.value - n0.value < 40, "the fold ran as a host loop"          # not 3000 iterations
        ov, (og,) = ad.grad(oracle, lambda T, t: nn.sfold(T, step(T), T.sscalar(1.0), t), [oracle.sconcrete(x)])
        np.testing.assert_allclose(float(dev.kfromS(dv)), float(ov), rtol=1e-12)
        np.testing.assert_allclose(dev.to_numpy(dg), np.asarray(og), rtol=1e-11)
    assert nn._step_is_product(dev, lambda a, e: dev.binary("mul", a, e), dev.sscalar(1.0), dev.sconcrete(x))
    assert not nn._step_is_product(dev, lambda a, e: dev.binary("add", a, e), dev.sscalar(1.0), dev.sconcrete(x))
    assert not nn._step_is_product(dev, lambda a, e: dev.binary("mul", a, a), dev.sscalar(1.0), dev.sconcrete(x))
    assert not nn._step_is_product(dev, lambda a, e: dev.binary("mul", dev.binary("mul", a, e), e), dev.sscalar(1.0), dev.sconcrete(x))
    xs = x[:40]
    f = lambda T: (lambda a, e: T.binary("add", T.binary("mul", a, e), e))
    dv, (dg,) = ad.grad(dev, lambda T, t: nn.sfold(T, f(T), T.sscalar(1.0), t), [dev.sconcrete(xs)])
    ov, (og,) = ad.grad(oracle, lambda T, t: nn.sfold(T, f(T), T.sscalar(1.0), t), [oracle.sconcrete(xs)])
    np.testing.assert_allclose(dev.to_numpy(dg), np.asarray(og), rtol=1e-12)


@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 100_003, 3_000_017])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_product_scan_kernel(dev, n, dtype):
    """The one-launch persistent product scan (phase A tile products, grid barrier, phase B gradient): values against
    a float64 host fold, the size-independent identity g_i * x_i = prod, exact zeros, and equality with the
    four-launch pipeline of round 1 is not required (different association), only the tolerance."""
    rng = np.random.default_rng(n)
    x = np.exp(rng.uniform(-3e-3, 3e-3, n)).astype(dtype)
    pr, g = dev.prod_fold_grad(dev.sconcrete(x), 1.5)
    p64 = float(np.prod(x.astype(np.float64)))
    tol = 1e-12 if dtype == np.float64 else 3e-5
    np.testing.assert_allclose(float(dev.kfromS(pr)), p64, rtol=tol * max(1, np.sqrt(n) / 30))
    got = dev.to_numpy(g).astype(np.float64)
    np.testing.assert_allclose(got * x.astype(np.float64), 1.5 * p64, rtol=tol * max(1, np.sqrt(n) / 30))
    if n >= 7:
        x[3] = 0.0
        _, g0 = dev.prod_fold_grad(dev.sconcrete(x), 1.0)
        g0 = dev.to_numpy(g0)
        assert np.count_nonzero(g0) == 1 and g0[3] != 0        # no division: exact zeros everywhere else

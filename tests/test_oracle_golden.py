"""Pins the CPU oracle against the reference's own literal golden vectors
(SURVEY.md 8c).  Everything here runs without a GPU."""
import numpy as np
import pytest

from horde_ad_b200 import ad, nn
from tests import programs as P


@pytest.mark.parametrize("name,prog,inp,expected", P.GATHER_GOLDEN, ids=[g[0] for g in P.GATHER_GOLDEN])
def test_gather_scatter_gradients(oracle, name, prog, inp, expected):
    # assertEqualUpToEpsilon' 1e-10 (rev' ... ) -- TestGatherSimplified.hs:154-398,961-1232
    _, (g,) = ad.grad(oracle, lambda T, t: prog(T, t), [oracle.sconcrete(inp)])
    np.testing.assert_allclose(oracle.to_numpy(g).reshape(-1), np.array(expected, dtype=float), atol=1e-10, rtol=0)


@pytest.mark.parametrize("nested,fused", P.GATHER_EQUAL_PAIRS)
def test_gather_nested_equals_fused_bit_exact(oracle, nested, fused):
    # interpretAstPrimal ... @?= ... (TestGatherSimplified.hs:208-235): bit-exact
    t = oracle.sconcrete(P.VALS72)
    a, b = oracle.to_numpy(nested(oracle, t)), oracle.to_numpy(fused(oracle, t))
    assert a.shape == b.shape and np.array_equal(a, b)


def test_bar_relu_gradient(oracle):
    # testBarReluADVal320 (TestGatherSimplified.hs:1283-1288), eps 1e-10
    x = oracle.sconcrete(0.001 * P.T128)
    _, (g,) = ad.grad(oracle, lambda T, t: P.barRelu(T, t), [x])
    np.testing.assert_allclose(oracle.to_numpy(g).reshape(-1), P.BARRELU_GOLDEN, atol=1e-10, rtol=0)


@pytest.mark.parametrize("name,K,A,wrt,expected,eps", P.CONV_GOLDEN, ids=[c[0] for c in P.CONV_GOLDEN])
@pytest.mark.parametrize("form", ["fused", "vectorised"])
def test_conv_gradients(oracle, form, name, K, A, wrt, expected, eps):
    # TestConvSimplified.hs:196-228,287-297
    conv = nn.conv2dPreservingS if form == "fused" else nn.conv2dPreservingS_vectorised
    Kt, At = oracle.sconcrete(K), oracle.sconcrete(A)
    if wrt == "A":
        _, (g,) = ad.grad(oracle, lambda T, a: conv(T, Kt, a), [At])
    else:
        _, (g,) = ad.grad(oracle, lambda T, k: conv(T, k, At), [Kt])
    np.testing.assert_allclose(oracle.to_numpy(g).reshape(-1), np.array(expected), atol=eps, rtol=0)


def test_readme_foo_known_answers(oracle):
    # README.md:75-76,89-90,104-105
    def foo(T, x, y, z):
        w = T.binary("mul", x, T.unary("sin", y))
        return T.binary("add", T.binary("atan2", z, w), T.binary("mul", z, w))
    ins = [oracle.srepl([2, 2], v) for v in (1.1, 2.2, 3.3)]
    val, grads = ad.grad(oracle, foo, ins)
    np.testing.assert_allclose(oracle.to_numpy(val), 4.242393641025528, rtol=1e-15)
    for g, e in zip(grads, (2.4396285219055063, -1.953374825727421, 0.9654825811012627)):
        np.testing.assert_allclose(oracle.to_numpy(g), e, rtol=1e-14)


def test_list_prod_gradient(oracle):
    # testListProd (TestAdaptorSimplified.hs:1043-1053): grad of the product of [1,2,3,4]
    x = oracle.sconcrete(np.array([1.0, 2.0, 3.0, 4.0]))
    val, (g,) = ad.grad(oracle, lambda T, t: T.prod_fold(t), [x])
    assert float(oracle.to_numpy(val)) == 24.0
    np.testing.assert_allclose(oracle.to_numpy(g), [24, 12, 8, 6], atol=1e-10)


def test_conv_adjoint_identity(oracle):
    # checkAdjoint (bench/ConvVjpBench.hs:564-584): <conv(K, A), Y> = <A, dInp(K, Y)> = <K, dKrn(A, Y)>
    rng = np.random.default_rng(0)
    K, A = rng.uniform(-.5, .5, (3, 2, 3, 2)), rng.uniform(-.5, .5, (4, 2, 5, 6))
    Y = rng.uniform(-.5, .5, (4, 3, 5, 6))
    lhs = np.sum(oracle.conv2dPreserving(K, A) * Y)
    np.testing.assert_allclose(lhs, np.sum(A * oracle.conv2d_dinp(K, Y)), rtol=1e-13)
    np.testing.assert_allclose(lhs, np.sum(K * oracle.conv2d_dkrn(A, Y, 3, 2)), rtol=1e-13)


def test_vectorised_blocks_match_fused(oracle):
    rng = np.random.default_rng(1)
    x = rng.uniform(-.5, .5, (3, 2, 6, 8))
    np.testing.assert_array_equal(nn.reluS_vectorised(oracle, x), nn.reluS(oracle, x))
    y, _ = oracle.maxpool2d(x, 2, 2)
    np.testing.assert_array_equal(nn.maxPool2dUnpaddedS_vectorised(oracle, 2, 2, x), y)
    K = rng.uniform(-.5, .5, (4, 2, 3, 3))
    np.testing.assert_allclose(nn.conv2dPreservingS_vectorised(oracle, K, x), oracle.conv2dPreserving(K, x), rtol=1e-13, atol=1e-15)
    lg, ex = rng.uniform(-1, 1, (10, 5)), np.eye(10)[:, :5]
    np.testing.assert_allclose(nn.lossSoftMaxCrossEntropyS_vectorised(oracle, ex, lg, 0.2),
                               nn.lossSoftMaxCrossEntropyS(oracle, ex, lg, 0.2), rtol=1e-14)


@pytest.mark.parametrize("fused", [True, "nodes"])
def test_cnn_fusion_levels_agree_on_the_oracle(oracle, fused):
    """The layer-level node, the per-op nodes and the vectorised gather program
    are the same function (loss and all 8 gradient leaves)."""
    from horde_ad_b200 import mnist_cnn
    rng = np.random.Generator(np.random.Philox(5))
    batch, hyper = 3, (1, 2, 3, 4)
    params = mnist_cnn.init_params(*hyper)
    glyphs, labels = rng.uniform(0, 1, (batch, 28, 28)), np.eye(10)[rng.integers(0, 10, batch)]

    def run(f):
        return ad.grad(oracle, lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, glyphs, labels, list(ps), fused=f), params)
    v0, g0 = run(False)
    v1, g1 = run(fused)
    np.testing.assert_allclose(float(v1), float(v0), rtol=1e-13)
    for a, b in zip(g1, g0):
        np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=1e-10, atol=1e-14)


def test_fold_goldens(oracle):
    """rfold known answers (TestRevFwdFold.hs:516-596): pins tmapAccumL/tfold and
    its reverse rule; eps 1e-10 as in the reference."""
    o = oracle
    x0 = o.sscalar(1.1)
    sin = lambda T, v: T.unary("sin", v)
    cases = [
        # testSin0FoldB3 / B4: constant step functions
        (lambda T, x: nn.sfold(T, lambda _x, _a: T.sscalar(7.0), T.sscalar(5.0), T.sreplicate(1, x)), 0.0),
        (lambda T, x: nn.sfold(T, lambda _x, _a: T.sscalar(7.0), x, T.srepl([1], 42.0)), 0.0),
        # testSin0Fold2: five nested sins of the accumulator
        (lambda T, x: nn.sfold(T, lambda a, _e: sin(T, a), x, T.srepl([5], 42.0)), 0.12389721944941383),
        # testSin0Fold3
        (lambda T, a0: nn.sfold(T, lambda _x, a: sin(T, a), T.sscalar(84.0), T.sreplicate(3, a0)), 0.4535961214255773),
        # testSin0Fold4
        (lambda T, a0: nn.sfold(T, lambda x, a: T.binary("atan2", sin(T, x), sin(T, a)),
                                T.binary("mul", T.sscalar(2.0), a0), T.sreplicate(3, a0)), -0.7053476446727861),
    ]
    for f, expected in cases:
        _, (g,) = ad.grad(o, f, [x0])
        np.testing.assert_allclose(float(o.to_numpy(g)), expected, atol=1e-10, rtol=0)

    # testSin0FoldNestedS1 (TestRevFwdFold.hs:2316-2326): a fold whose step is a fold; the reference's eps
    # is 1e-10 absolute, far above the value itself, so the known answer is also checked relatively
    def nested(T, a0):
        inner = lambda x, a: nn.sfold(T, lambda x2, a2: T.binary("mul", T.binary("mul", T.sscalar(0.7), x2), a2),
                                      a, T.sreplicate(7, x))
        return nn.sfold(T, inner, a0, T.sreplicate(3, a0))
    _, (g,) = ad.grad(o, nested, [x0])
    np.testing.assert_allclose(float(o.to_numpy(g)), 2.0504979297616553e-43, atol=1e-10, rtol=0)
    np.testing.assert_allclose(float(o.to_numpy(g)), 2.0504979297616553e-43, rtol=1e-12)

    # testSin0Fold6 (= 6) and testSin0Fold7 (= 250): tensor-valued accumulators
    def fold6(T, a0):
        acc0 = T.sreplicate(2, T.sreplicate(1, a0))
        step = lambda x, a: T.stranspose([1, 0], T.binary("add", T.stranspose([1, 0], x), T.sreplicate(1, T.sreplicate(2, a))))
        return T.ssum0(nn.sfold(T, step, acc0, T.sreplicate(2, a0)))

    def fold7(T, a0):
        acc0 = T.sreplicate(2, T.sreplicate(5, a0))
        step = lambda x, _a: T.stranspose([1, 0], T.sreplicate(5, T.ssum(T.stranspose([1, 0], x))))
        return T.ssum0(nn.sfold(T, step, acc0, T.sreplicate(2, a0)))
    for f, expected in ((fold6, 6.0), (fold7, 250.0)):
        _, (g,) = ad.grad(o, f, [x0])
        np.testing.assert_allclose(float(o.to_numpy(g)), expected, atol=1e-10, rtol=0)


def test_fcnn_and_rnn_gradients_match_finite_differences(oracle):
    """MnistFcnnRanked2 (logistic/softMax1/lossCrossEntropyV, Float hidden layer)
    and MnistRnnShaped2 on the oracle: the dual-number gradient agrees with
    central differences of the primal (their losses, unlike the CNN's fused
    whole-tensor softmax with its custom dual, are plain compositions)."""
    from horde_ad_b200 import mnist_fcnn
    rng = np.random.Generator(np.random.Philox(3))
    datum, target = rng.uniform(0, 1, 784), np.eye(10)[3]
    params = mnist_fcnn.init_params(6, 5, rng_range=0.3, q=np.float64)
    f = lambda ps: float(np.asarray(ad.grad(oracle, lambda T, *p: mnist_fcnn.afcnnMnistLoss2(T, datum, target, list(p)), ps)[0]))
    _, grads = ad.grad(oracle, lambda T, *p: mnist_fcnn.afcnnMnistLoss2(T, datum, target, list(p)), params)
    for k in (0, 2, 5):
        idx = tuple(rng.integers(0, s) for s in params[k].shape)
        e = 1e-6
        pp = [p.copy() for p in params]; pp[k][idx] += e
        pm = [p.copy() for p in params]; pm[k][idx] -= e
        np.testing.assert_allclose(np.asarray(grads[k])[idx], (f(pp) - f(pm)) / (2 * e), rtol=1e-6, atol=1e-9)


def test_committed_golden_fixture_is_in_sync(oracle):
    """tests/golden/reference_goldens.json (written by tests/golden/make_golden.py)
    carries the same literals the programs above are checked against, and the
    oracle reproduces the conv entries straight from the file."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))
    assert [e["name"] for e in g["gather_scatter_gradients"]] == [x[0] for x in P.GATHER_GOLDEN]
    for e in g["conv_gradients"]:
        K = np.array(e["K"]).reshape(e["K_shape"])
        A = np.array(e["A"]).reshape(e["A_shape"])
        if e["wrt"] == "A":
            got = oracle.conv2d_dinp(K, np.ones((A.shape[0], K.shape[0], A.shape[2], A.shape[3])))
        else:
            got = oracle.conv2d_dkrn(A, np.ones((A.shape[0], K.shape[0], A.shape[2], A.shape[3])), K.shape[2], K.shape[3])
        np.testing.assert_allclose(got.reshape(-1), e["expected_gradient"], atol=e["eps"], rtol=0)


def test_forward_reverse_duality_on_the_oracle(oracle):
    """<grad f, ds> = jvp f ds (the reference's QuickCheck property,
    TestConvQuickCheck.hs:182-268) for the CNN, FCNN and a gather program."""
    from horde_ad_b200 import mnist_cnn, mnist_fcnn
    rng = np.random.Generator(np.random.Philox(8))
    glyphs, labels = rng.uniform(0, 1, (3, 28, 28)), np.eye(10)[rng.integers(0, 10, 3)]
    params = mnist_cnn.init_params(1, 2, 3, 4)
    ds = [rng.uniform(-1, 1, p.shape) for p in params]
    for fused in (True, "nodes", False):
        f = lambda T, *ps: mnist_cnn.convMnistLossFusedS(T, glyphs, labels, list(ps), fused=fused)
        _, grads = ad.grad(oracle, f, params)
        _, tan = ad.jvp(oracle, f, params, ds)
        lhs = sum(float(np.sum(np.asarray(g) * d)) for g, d in zip(grads, ds))
        np.testing.assert_allclose(float(np.asarray(tan)), lhs, rtol=1e-10)
    datum, target = rng.uniform(0, 1, 784), np.eye(10)[2]
    fp = mnist_fcnn.init_params(5, 4, rng_range=0.3, q=np.float64)
    fds = [rng.uniform(-1, 1, p.shape) for p in fp]
    g = lambda T, *ps: mnist_fcnn.afcnnMnistLoss2(T, datum, target, list(ps))
    _, grads = ad.grad(oracle, g, fp)
    _, tan = ad.jvp(oracle, g, fp, fds)
    np.testing.assert_allclose(float(np.asarray(tan)), sum(float(np.sum(np.asarray(a) * d)) for a, d in zip(grads, fds)), rtol=1e-10)


def test_tanh_affine_node_equals_the_unfused_layer_on_the_oracle(oracle):
    """rnnMnistLayerS with its body as one node (tanh_affine) against the literal program, in reverse mode
    (all 8 gradient leaves) and in forward mode (duality <grad f, ds> = jvp f ds)."""
    from horde_ad_b200 import mnist_rnn
    rng = np.random.Generator(np.random.Philox(31))
    batch, width = 5, 12
    glyphs, labels = rng.uniform(0, 1, (batch, 28, 28)), np.eye(10)[rng.integers(0, 10, batch)]
    params = mnist_rnn.init_params(width, rng_range=0.2)
    f = lambda fused: (lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=fused))
    v0, g0 = ad.grad(oracle, f(False), params)
    v1, g1 = ad.grad(oracle, f(True), params)
    np.testing.assert_allclose(v1, v0, rtol=1e-14)
    for a, b in zip(g1, g0):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-16)
    ds = [rng.uniform(-1, 1, p.shape) for p in params]
    _, t = ad.jvp(oracle, f(True), params, ds)
    np.testing.assert_allclose(float(t), sum(float(np.sum(g * d)) for g, d in zip(g1, ds)), rtol=1e-10)
    # ... and with `wX x + wS s` as the one node smatmul2_sum (paired products, paired weight cotangents)
    v2, g2 = ad.grad(oracle, f("pair"), params)
    np.testing.assert_allclose(v2, v0, rtol=1e-14)
    for a, b in zip(g2, g0):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-16)
    _, t2 = ad.jvp(oracle, f("pair"), params, ds)
    np.testing.assert_allclose(float(t2), float(t), rtol=1e-12)


def test_cnn_opp3_folded_value_and_gradient(oracle):
    """testCNNOPP3 / testCNNOPP3b (TestGatherSimplified.hs:1365-1391): the toy conv/max-pool built from
    rbuild, rindex and ifH folds to [14, 0, 14, 0, 0, ...] on a glyph of sevens, and its simplified gradient
    artifact is `soneHot x [0,0,0,1] + soneHot x [0,1,1,1]` with x = the sum of two cotangent entries."""
    g = np.full((3, 3, 3, 3), 7.0)
    prog = lambda T, u: P.maxPool2dUnpadded3(T, P.conv2dPreserving3(T, u))
    v, (gr,) = ad.grad(oracle, prog, [g])
    assert np.asarray(v).shape == (2, 2, 2, 2) and np.asarray(v).reshape(-1).tolist() == P.CNNOPP3_GOLDEN
    expect = np.zeros((3, 3, 3, 3))
    expect[0, 0, 0, 1] = expect[0, 1, 1, 1] = 2.0
    assert np.array_equal(np.asarray(gr), expect)


def test_cnn_opp6_and_opp7_folded_values(oracle):
    """testCNNOPP6 / testCNNOPP7 (TestGatherSimplified.hs:1580-1650; the same literals at
    TestConvSimplified.hs:1507,1546): folded values of the doubled-index variants."""
    g = np.full((2, 2, 2, 2), 7.0)
    v6 = P.maxPool2dUnpadded3(oracle, P.conv2dPreserving3z(oracle, g))
    v7 = P.maxPool2dUnpadded3y(oracle, P.conv2dPreserving3y(oracle, g))
    assert np.asarray(v6).reshape(-1).tolist() == P.CNNOPP6_GOLDEN
    assert np.asarray(v7).reshape(-1).tolist() == P.CNNOPP7_GOLDEN


@pytest.mark.parametrize("name,mk,expected", P.VSTACK_GOLDEN, ids=[v[0] for v in P.VSTACK_GOLDEN])
def test_vstack_known_answers(oracle, name, mk, expected):
    """testTrustVstackConcatRepl10 / Iota10 / ReplIota10 and the folded value of vstackBuild
    (TestAdaptorSimplified.hs:815-835, 949): slices + appends and rbuild1 agree with the literals."""
    a, b, c = mk(10)
    assert np.asarray(P.vstackABC(oracle, a, b, c)).tolist() == expected
    assert np.asarray(P.vstackBuild(oracle, a, b, c)).tolist() == expected


def test_gather_transpose33_gradients(oracle):
    """testGatherTranspose33 / testGatherTransposeBuild33 (TestGatherSimplified.hs:564-606): the chain of
    21 transpositions + 2 reshapes + matmul; pins the permutation convention (also for permutations
    shorter than the rank) against the reference's 128-element literals, eps 1e-10 as there."""
    for f, expected in ((P.gatherTranspose33, P.GT33_GOLDEN), (P.gatherTransposeBuild33, P.GT33_BUILD_GOLDEN)):
        _, (g,) = ad.grad(oracle, f, [P.T128])
        np.testing.assert_allclose(np.asarray(g).reshape(-1), np.array(expected), atol=1e-10, rtol=0)


def test_gather_reshape22_gradients(oracle):
    """testGatherReshape22 / testGatherReshapeBuild22 (TestGatherSimplified.hs:439-468): five nested
    reshapes are the identity on the gradient (ones), 0+1+2+3 = 6 under rbuild1 4."""
    t = np.tile(np.array([0.0, 1.0]), (6, 1))
    resh = lambda T, x: T.sreshape([2, 6], T.sreshape([3, 1, 2, 1, 1, 2], T.sreshape([1, 12, 1], T.sreshape([3, 1, 1, 4], T.sreshape([2, 2, 3], x)))))
    _, (g,) = ad.grad(oracle, resh, [t])
    assert np.array_equal(np.asarray(g), np.ones((6, 2)))
    build = lambda T, x: nn.sbuild1(T, 4, lambda i: resh(T, T.binary("mul", x, T.srepl([6, 2], float(i)))))
    _, (g,) = ad.grad(oracle, build, [t])
    assert np.array_equal(np.asarray(g), np.full((6, 2), 6.0))


GATHER_COND_CASES = [(P.gatherCond, (7, 2), P.GATHER_COND_GOLDEN), (P.gatherCond2, (7, 2), P.GATHER_COND_GOLDEN),
                     (P.gatherCond3, (7, 2), P.GATHER_COND34_GOLDEN), (P.gatherCond4, (7, 2), P.GATHER_COND34_GOLDEN),
                     (P.gatherCond5, (2, 4, 2), P.GATHER_COND56_GOLDEN), (P.gatherCond6, (2, 4, 2), P.GATHER_COND56_GOLDEN)]


@pytest.mark.parametrize("prog,shape,expected", GATHER_COND_CASES, ids=[c[0].__name__ for c in GATHER_COND_CASES])
def test_gather_with_conditional_index_gradients(oracle, prog, shape, expected):
    """testGatherCond .. testGatherCond6 and their rbuild1 variants (TestGatherSimplified.hs:716-915)."""
    t = np.broadcast_to(np.array([0.0, 1.0]), shape).copy()
    _, (g,) = ad.grad(oracle, prog, [t])
    assert np.asarray(g).reshape(-1).tolist() == expected
    build = lambda T, x: nn.sbuild1(T, 4, lambda i: prog(T, T.binary("mul", x, T.srepl(list(shape), float(i)))))
    _, (g,) = ad.grad(oracle, build, [t])
    assert np.asarray(g).reshape(-1).tolist() == [6 * x for x in expected]


@pytest.mark.parametrize("name,prog,expected", P.GATHER_BUILD_GOLDEN, ids=[g[0] for g in P.GATHER_BUILD_GOLDEN])
def test_gather_scatter_build_variants(oracle, name, prog, expected):
    """testGather(Nested)Build1/2/12 and testScatter(Nested)Build1/2/12 (TestGatherSimplified.hs:161-398,
    968-1232): the gather / scatter programs under rbuild1 with host conditionals and indexing."""
    t = np.tile(np.array([0.0, 1.0]), (7, 1))
    _, (g,) = ad.grad(oracle, prog, [t])
    assert np.asarray(g).reshape(-1).tolist() == expected


@pytest.mark.parametrize("name,f,inputs,expected", P.ADAPTOR_REV_GOLDEN, ids=[g[0] for g in P.ADAPTOR_REV_GOLDEN])
def test_relu_and_bar_relu_known_answers(oracle, name, f, inputs, expected):
    """testReluSimpler/3, testReluMax/3, testBarRelu/3, testBarReluMax/30/31 (TestAdaptorSimplified.hs:1247-1251,
    1300-1345, 1404-1462, 1850-1895), eps 1e-10."""
    _, grads = ad.grad(oracle, lambda T, *xs: T.ssum0(f(T, *xs)), [oracle.sconcrete(x) for x in inputs])
    for g, e in zip(grads, expected):
        np.testing.assert_allclose(oracle.to_numpy(g), e, atol=1e-10, rtol=0)


@pytest.mark.parametrize("name,f,x,dt,expected", P.ADAPTOR_VJP_GOLDEN, ids=[g[0] for g in P.ADAPTOR_VJP_GOLDEN])
def test_bar_relu_vjp_known_answers(oracle, name, f, x, dt, expected):
    """testBarReluDt / testBarReluMaxDt (TestAdaptorSimplified.hs:1843-1848, 1872-1877): vjp with cotangent 42.2."""
    _, (g,) = ad.grad(oracle, f, [oracle.sconcrete(x)], dt=oracle.sconcrete(dt))
    np.testing.assert_allclose(oracle.to_numpy(g), expected, atol=1e-10, rtol=1e-11)


@pytest.mark.parametrize("name,f,x,dx,expected", P.ADAPTOR_FWD_GOLDEN, ids=[g[0] for g in P.ADAPTOR_FWD_GOLDEN])
def test_bar_relu_forward_known_answers(oracle, name, f, x, dx, expected):
    """testBarReluMax3CFwd / testBarReluMax3FwdS (TestAdaptorSimplified.hs:1896-1925): cjvp / jvp."""
    _, t = ad.jvp(oracle, f, [oracle.sconcrete(x)], [oracle.sconcrete(dx)])
    np.testing.assert_allclose(oracle.to_numpy(t), expected, atol=1e-10, rtol=1e-11)


@pytest.mark.parametrize("name,f,x,expected", P.SCAN_GOLDEN, ids=[g[0] for g in P.SCAN_GOLDEN])
def test_scan_known_answers(oracle, name, f, x, expected):
    """testSin0Scan1/2/3/4/6 (TestRevFwdFold.hs:934-1000): rscan = a host loop over the slices (tscan,
    OpsConcrete.hs:66-90) whose unrolled ops transpose to the reference's gradients, eps 1e-10."""
    _, (g,) = ad.grad(oracle, lambda T, t: T.ssum0(f(T, t)), [oracle.sconcrete(x)])
    np.testing.assert_allclose(oracle.to_numpy(g), expected, atol=1e-10, rtol=0)


@pytest.mark.parametrize("dtype", [np.float64, np.int64, np.int32, np.int16, np.int8])
def test_overleaf_gather_with_wraparound(oracle, dtype):
    """testOverleaf, testOverleafInt64p / CIntp / Int16p / Int8p (TestAdaptorSimplified.hs:507-585): gradient of
    ssum0 (sgather1 @50 v (\\i -> [remH i 28])) = sscatter1 with collisions = [2 x 22, 1 x 6], in the tensor's own
    (integer) arithmetic; testOverleafGrad1 (:515-519): a one-element vector collects all 50."""
    v = np.arange(28).astype(dtype)
    val, (g,) = ad.grad(oracle, P.overleaf, [oracle.sconcrete(v)])
    got = oracle.to_numpy(g)
    assert got.dtype == np.dtype(dtype)
    np.testing.assert_array_equal(got, np.array(P.OVERLEAF_GRAD).astype(dtype))
    assert oracle.to_numpy(val) == np.array((np.arange(50) % 28).sum()).astype(dtype)      # Int8 wraps like the reference's
    _, (g1,) = ad.grad(oracle, P.overleaf, [oracle.sconcrete(np.array([27.0]))])
    np.testing.assert_array_equal(oracle.to_numpy(g1), [50.0])


def test_high_rank_logistic_and_bar_relu_known_answers(oracle):
    """testLogistic0 / 5 / 52, testBarReluADVal / ADVal3 / ADValDt (TestHighRankSimplified.hs:527-560, 696-712): rank-5
    and rank-7 tensors through logistic's hand-written dual (CommonRankedOps.hs:55-63) and the atan2 / sin / relu chain."""
    g = lambda f, x, dt=None: oracle.to_numpy(ad.grad(oracle, f, [oracle.sconcrete(x)], dt=dt)[1][0])
    ssum = lambda f: (lambda T, t: T.ssum0(f(T, t)))
    np.testing.assert_allclose(g(ssum(nn.logistic), np.array(3.0)), 4.5176659730912e-2, atol=1e-10, rtol=0)
    np.testing.assert_allclose(g(ssum(nn.logistic), P.T16), P.LOGISTIC5_GRAD, atol=1e-10, rtol=0)
    np.testing.assert_allclose(g(ssum(lambda T, t: nn.logistic(T, nn.logistic(T, t))), P.T16), P.LOGISTIC52_GRAD, atol=1e-10, rtol=0)
    np.testing.assert_allclose(g(ssum(P.barRelu), P.T48), P.BARRELU_T48_GRAD, atol=1e-10, rtol=0)
    np.testing.assert_allclose(g(ssum(P.barRelu), 0.001 * P.T48), P.BARRELU_T48_SCALED_GRAD, atol=1e-10, rtol=0)
    dt = oracle.sconcrete(np.full(P.T16.shape, 42.2))
    np.testing.assert_allclose(g(P.barRelu, P.T16, dt), P.BARRELU_T16_DT_GRAD, atol=1e-6, rtol=0)


def test_map_accum_empty_and_single_step_known_answers(oracle):
    """testSin0rmapAccumRD0R0 / 0R2 / 0R3 / 0S / 01SN (TestRevFwdFold.hs:1264-1281, 1542-1625): tmapAccum over an
    EMPTY sequence returns the initial accumulator (gradient 1), zero-length tensors flow through replicate /
    reverse / + with a zero gradient, and a single step with a pair accumulator gives cos 1.1."""
    x0 = oracle.sconcrete(np.array(1.1))
    # 0R0: rreverse (rreverse (rreplicate 0 x0) + rreplicate 0 x0), rev' = gradient of the (empty) sum
    f0 = lambda T, x: T.ssum0(T.sreverse(T.binary("add", T.sreverse(T.sreplicate(0, x)), T.sreplicate(0, x))))
    v, (g,) = ad.grad(oracle, f0, [x0])
    assert float(oracle.to_numpy(v)) == 0.0 and float(oracle.to_numpy(g)) == 0.0
    # 0S / 0R2 / 0R3: tproject1 (tmapAccumR g x0 (empty)) with g x _a = (sin x, sin x)
    def f1(T, x):
        acc, bs = nn.smapAccumL(T, lambda a, _e: (T.unary("sin", a), T.unary("sin", a)), x, T.srepl([0], 0.0))
        assert bs is None
        return acc
    _, (g,) = ad.grad(oracle, f1, [x0])
    np.testing.assert_allclose(oracle.to_numpy(g), 1.0, atol=1e-10, rtol=0)
    # 01SN: one step, accumulator = (3, x0), g xh _a = let x = snd xh in ((sin x, sin x), sin x); result = the stacked bs
    def f2(T, x):
        step = lambda xh, _a: ((T.unary("sin", xh[1]), T.unary("sin", xh[1])), T.unary("sin", xh[1]))
        _acc, bs = nn.smapAccumL(T, step, (T.sscalar(3.0), x), T.srepl([1], 0.0))
        return T.ssum0(bs)
    _, (g,) = ad.grad(oracle, f2, [x0])
    np.testing.assert_allclose(oracle.to_numpy(g), 0.4535961214255773, atol=1e-10, rtol=0)


@pytest.mark.parametrize("name,f,x,expected", P.BUILD_GOLDEN, ids=[g[0] for g in P.BUILD_GOLDEN])
def test_high_rank_build_known_answers(oracle, name, f, x, expected):
    """testBraidedBuilds/1, testConcatBuild2/22/0m/1m (TestHighRankSimplified.hs:565-690): rbuild / rfromList /
    rindex / rslice host loops incl. an index that runs out of bounds (zeros), eps 1e-10."""
    _, (g,) = ad.grad(oracle, lambda T, t: T.ssum0(f(T, t)), [oracle.sconcrete(x)])
    np.testing.assert_allclose(oracle.to_numpy(g), expected, atol=1e-10, rtol=0)


def test_det_square_known_answer(oracle):
    """testDetSquare3 (TestGatherSimplified.hs:1825-1838, eps 1e-5): the gradient of the permutation-sum determinant
    of a 6 x 6 matrix (720 data-dependent gathers + product folds) = the reference's literal = det(a) * inv(a)^T."""
    v, (g,) = ad.grad(oracle, P.detSquare, [oracle.sconcrete(P.DET6_IN)])
    np.testing.assert_allclose(float(oracle.to_numpy(v)), np.linalg.det(P.DET6_IN), rtol=1e-10)
    np.testing.assert_allclose(oracle.to_numpy(g), P.DET6_GRAD, atol=1e-5, rtol=0)
    np.testing.assert_allclose(oracle.to_numpy(g), np.linalg.det(P.DET6_IN) * np.linalg.inv(P.DET6_IN).T, rtol=1e-8, atol=1e-12)


def test_fold_goldens_with_tensor_steps(oracle):
    """testSin0Fold5 (1.2992412552109085) and testSin0Fold8 (-2.200311410593445), TestRevFwdFold.hs:572-582, 598-608:
    fold steps built from rsum / rtr / rreplicate / atan2H over rank-1 and rank-2 intermediates, eps 1e-10."""
    o = oracle
    tr = lambda T, t: T.stranspose([1, 0] + list(range(2, len(T.shape(t)))), t)
    sin = lambda T, v: T.unary("sin", v)

    def fold5(T, a0):
        def step(x, a):                                   # x : [], a : [2, 5]
            inner = T.ssum(sin(T, T.ssum(tr(T, T.sreplicate(7, a)))))        # [7,2,5] -> [2,7,5] -> [7,5] -> sin -> [5]
            return T.ssum(T.binary("atan2", sin(T, T.sreplicate(5, x)), inner))
        return nn.sfold(T, step, T.binary("mul", T.sscalar(2.0), a0), T.sreplicate(3, T.sreplicate(2, T.sreplicate(5, a0))))

    def fold8(T, a0):
        def step(x, a):                                   # x : [2, 5], a : []
            num = T.ssum(tr(T, sin(T, x)))                                    # [5,2] -> [2]
            den = T.sreplicate(2, sin(T, T.ssum(T.sreplicate(7, a))))         # [2]
            return tr(T, T.sreplicate(5, T.binary("atan2", num, den)))        # [5,2] -> [2,5]
        acc0 = T.sreplicate(2, T.sreplicate(5, T.binary("mul", T.sscalar(2.0), a0)))
        return T.ssum0(nn.sfold(T, step, acc0, T.sreplicate(3, a0)))
    for f, expected in ((fold5, 1.2992412552109085), (fold8, -2.200311410593445)):
        _, (g,) = ad.grad(o, f, [o.sscalar(1.1)])
        np.testing.assert_allclose(float(o.to_numpy(g)), expected, atol=1e-10, rtol=0)


def test_foo_no_go_known_answer(oracle):
    """testFooNoGo0 (TestAdaptorSimplified.hs:1747-1752, eps 1e-6)."""
    _, (g,) = ad.grad(oracle, lambda T, t: T.ssum0(P.fooNoGo(T, t)), [oracle.sconcrete(P.FOONOGO_IN)])
    np.testing.assert_allclose(oracle.to_numpy(g), P.FOONOGO_GRAD, atol=1e-6, rtol=0)


def test_foo_float_rank5_known_answer(oracle):
    """testFoo (TestHighRankSimplified.hs:102-106): grad (rsum0 @5 @Float . foo) (t16, t16, t16) in single precision,
    eps 1e-3 as in the reference."""
    exp = [[-4.6947093,1.5697206,-1.6332961,0.34882763,1.5697206,-1.0,-0.9784988,-0.9158946,6.6326222,3.6699238,7.85237,-2.9069107,17.976654,0.3914159,32.98194,19.807974],
           [6.943779,-1.436789,33.67549,0.22397964,-1.436789,-1.0,-0.975235,-0.90365005,147.06645,-73.022705,-9.238474,-10.042692,-980.2843,-7.900571,-14.451739,436.9084],
           [-4.8945336,2.067469,-1.7196897,1.3341143,2.067469,1.0,0.99846554,0.99536234,6.6943173,3.7482092,7.977362,-3.1475093,18.000969,0.48736274,33.01224,19.845064]]
    t16 = P.T16.astype(np.float32)
    _, grads = ad.grad(oracle, lambda T, x, y, z: T.ssum0(P._foo(T, x, y, z)), [oracle.sconcrete(t16)] * 3)
    for g, e in zip(grads, exp):
        got = oracle.to_numpy(g)
        assert got.dtype == np.float32
        np.testing.assert_allclose(got.reshape(-1), np.array(e, np.float32), atol=1e-3, rtol=1e-5)


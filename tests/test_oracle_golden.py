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
    assert float(val) == 24.0
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

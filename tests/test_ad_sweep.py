"""Host-side logic of the reverse sweep (horde-ad_b200/ad.py) that the device path shares: deferred products
accumulated by `smatmul2_acc` / `smatmul2_acc_pair` (interior values) or kept as a list and evaluated once by
`smatmul2_list` (inputs: weights of an unrolled loop), zero-padded slice cotangents joined by one `sappend`.
Runs on the CPU oracle with a call-counting wrapper; values are checked against the plain rules."""
import numpy as np
import pytest

from horde_ad_b200 import ad
from oracle.concrete import ConcreteBackend


class Counting(ConcreteBackend):
    """ConcreteBackend that counts the calls of the methods the sweep chooses between."""

    def __init__(self):
        super().__init__()
        self.calls = {}
        self.groups = []

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def smatmul2_group(self, results, products):
        """Counted by its shape (results, products, accumulated results); the products inside are not counted."""
        self.groups.append((len(results), len(products), sum(a is not None for a, _ in results)))
        return ConcreteBackend.smatmul2_group(ConcreteBackend(), results, products)


for _name in ("smatmul2", "smatmul2_acc", "smatmul2_acc_pair", "smatmul2_pair", "smatmul2_sum", "smatmul2_list",
              "sappend", "srepl", "add_accumulate"):
    def _wrap(name):
        base = getattr(ConcreteBackend, name)

        def method(self, *a, **k):
            self._count(name)
            return base(self, *a, **k)
        return method
    setattr(Counting, _name, _wrap(_name))


def fd_grad(f, x, eps=1e-6):
    g = np.zeros_like(x)
    for i in np.ndindex(*x.shape):
        d = np.zeros_like(x)
        d[i] = eps
        g[i] = (f(x + d) - f(x - d)) / (2 * eps)
    return g


def test_two_halves_of_a_state_become_one_append():
    rng = np.random.default_rng(0)
    s = rng.uniform(-1, 1, (6, 3))

    def f(T, t):
        a, b = T.sslice(0, 2, t), T.sslice(2, 4, t)                      # two windows that tile the leading axis
        return T.binary("add", T.ssum0(T.unary("sin", a)), T.ssum0(T.binary("mul", b, b)))
    B = Counting()
    _, (g,) = ad.grad(B, f, [s])
    np.testing.assert_allclose(g, np.concatenate([np.cos(s[:2]), 2 * s[2:]]), rtol=1e-14)
    # one append, no zero blocks (the only srepl is the cotangent of ones of the result), nothing added
    assert B.calls.get("sappend") == 1
    assert B.calls.get("srepl", 0) == 1 and B.calls.get("add_accumulate", 0) == 0


@pytest.mark.parametrize("windows", [[(0, 2), (3, 3)],            # a gap
                                     [(0, 4), (2, 4)],            # an overlap
                                     [(1, 2)],                    # a single interior window
                                     [(0, 2), (2, 2), (4, 2)],    # three windows that tile the axis
                                     [(4, 2), (0, 2), (2, 2)]])   # ... arriving out of order
def test_slice_windows_against_finite_differences(windows):
    rng = np.random.default_rng(1)
    s = rng.uniform(-1, 1, (6, 2))

    def prog(T, t):
        out = None
        for k, (i, n) in enumerate(windows):
            term = T.ssum0(T.unary("sin", T.binary("mul", T.sslice(i, n, t), T.srepl([n, 2], float(k + 1)))))
            out = term if out is None else T.binary("add", out, term)
        return out
    B = ConcreteBackend()
    _, (g,) = ad.grad(B, prog, [s])
    ref = fd_grad(lambda x: sum(np.sin(x[i:i + n] * (k + 1)).sum() for k, (i, n) in enumerate(windows)), s)
    np.testing.assert_allclose(g, ref, rtol=1e-7, atol=1e-9)


def test_weight_used_at_every_step_accumulates_in_the_product():
    rng = np.random.default_rng(2)
    w, x = rng.uniform(-.3, .3, (5, 5)), rng.uniform(-.5, .5, (5, 4))

    def f(T, W, X):
        h = X
        for _ in range(3):
            h = T.unary("tanh", T.smatmul2(W, h))
        return T.ssum0(h)
    B = Counting()
    _, (gw, gx) = ad.grad(B, f, [w, x])
    # W is an INPUT used three times: its three cotangent products are kept as a list and evaluated once, when the
    # sweep reaches the input (one product over the concatenated reduced axes on the device)
    assert B.calls.get("smatmul2_list") == 1 and B.calls.get("smatmul2_acc", 0) == 0
    # an INTERIOR value used three times (W = 2 * V): its cotangent products wait until something needs a deferred
    # product, then go out together with the state's (one launch for two results), accumulated from the second on
    B2 = Counting()
    _, (gv, _) = ad.grad(B2, lambda T, V, X: f(T, T.binary("mul", V, T.srepl([5, 5], 2.0)), X), [w / 2, x])
    assert B2.groups == [(2, 2, 0), (2, 2, 1)] and B2.calls.get("smatmul2_acc") == 1
    assert B2.calls.get("smatmul2_list", 0) == 0
    np.testing.assert_allclose(gv, 2 * gw, rtol=1e-13)

    def val(W):
        h = x
        for _ in range(3):
            h = np.tanh(W @ h)
        return h.sum()
    np.testing.assert_allclose(gw, fd_grad(val, w), rtol=1e-7, atol=1e-9)


def test_paired_layer_node_pairs_its_cotangents():
    rng = np.random.default_rng(3)
    wx, ws = rng.uniform(-.3, .3, (4, 4)), rng.uniform(-.3, .3, (4, 4))
    x0, s0 = rng.uniform(-.5, .5, (4, 3)), rng.uniform(-.5, .5, (4, 3))

    def f(T, WX, WS, X, S):
        x, s = X, S
        for _ in range(3):                                   # both weights used three times
            y = T.unary("tanh", T.smatmul2_sum(WX, x, WS, s))
            x, s = s, y
        return T.ssum0(y)
    B = Counting()
    v, grads = ad.grad(B, f, [wx, ws, x0, s0])
    assert B.calls.get("smatmul2_sum") == 3
    # dWX, dWS and dS (s0 is x of the second step): one list product each at the end; dX: a single product
    assert B.calls.get("smatmul2_list") == 3
    assert B.calls.get("smatmul2_acc_pair", 0) == 0
    assert B.groups[0] == (2, 2, 0)                          # (dx, ds) of the last step share one launch

    def val(WX, WS, X, S):
        x, s = X, S
        for _ in range(3):
            y = np.tanh(WX @ x + WS @ s)
            x, s = s, y
        return y.sum()
    args = [wx, ws, x0, s0]
    for k in range(4):
        ref = fd_grad(lambda a: val(*[a if j == k else args[j] for j in range(4)]), args[k])
        np.testing.assert_allclose(grads[k], ref, rtol=1e-7, atol=1e-9)


def test_two_layer_recurrence_launches_three_products_per_step():
    """rnnMnistTwoS's shape (MnistRnnShaped2.hs:71-93): at every step of the sweep the cotangent of the first layer's
    state gets wS1^T d1' (recurrence) + wX2^T d2 (second layer), the second layer's state gets wS2^T d2 -- three
    deferred products, two results, ONE group launch."""
    rng = np.random.default_rng(5)
    n, b, steps = 4, 3, 4
    ws = [rng.uniform(-.3, .3, (n, n)) for _ in range(4)]
    bs = [rng.uniform(-.1, .1, (n,)) for _ in range(2)]
    xs = rng.uniform(-.5, .5, (steps, n, b))

    def f(T, wx1, ws1, wx2, ws2, b1, b2):
        h1 = h2 = T.srepl([n, b], 0.0)
        for t in range(steps):
            h1 = T.smatmul2_sum_tanh(wx1, T.sconcrete(xs[t]), ws1, h1, b1)
            h2 = T.smatmul2_sum_tanh(wx2, h1, ws2, h2, b2)
        return T.ssum0(h2)
    B = Counting()
    _, grads = ad.grad(B, f, ws + bs)
    # last time step: no recurrence cotangent yet; first: the initial states are constants
    assert B.groups == [(2, 2, 0)] + [(2, 3, 0)] * (steps - 2) + [(1, 2, 0)], B.groups
    assert B.calls.get("smatmul2_list") == 4

    def val(wx1, ws1, wx2, ws2, b1, b2):
        h1 = h2 = np.zeros((n, b))
        for t in range(steps):
            h1 = np.tanh(wx1 @ xs[t] + ws1 @ h1 + b1[:, None])
            h2 = np.tanh(wx2 @ h1 + ws2 @ h2 + b2[:, None])
        return h2.sum()
    args = ws + bs
    for k in range(6):
        ref = fd_grad(lambda a: val(*[a if j == k else args[j] for j in range(6)]), args[k])
        np.testing.assert_allclose(grads[k], ref, rtol=1e-7, atol=1e-9)


def test_wavefront_schedule_is_the_row_by_row_program():
    """mnist_rnn's fused="wave" (layer 1 of row t and layer 2 of row t - 1 as one smatmul2_group forward) computes
    the values and gradients of the row-by-row program of rnnMnistTwoS (MnistRnnShaped2.hs:71-117): on the oracle the
    two schedules agree to the last bit forward and to rounding in the gradients (other summation order of the shared
    cotangents)."""
    from horde_ad_b200 import mnist_rnn
    rng = np.random.Generator(np.random.Philox(3))
    batch, width = 5, 12
    glyphs, labels = rng.uniform(0, 1, (batch, 28, 28)), np.eye(10)[rng.integers(0, 10, batch)]
    params = mnist_rnn.init_params(width, rng_range=0.2)

    def run(fused):
        B = Counting()
        val, grads = ad.grad(B, lambda T, *ps: mnist_rnn.rnnMnistLossFusedS(T, glyphs, labels, list(ps), fused=fused), params)
        return B, float(val), grads
    Bw, vw, gw = run("wave")
    Bp, vp, gp = run("pair")
    _, v0, g0 = run(False)
    assert vw == vp
    np.testing.assert_allclose(vw, v0, rtol=1e-13)
    for a, b, c in zip(gw, gp, g0):
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-15)
        np.testing.assert_allclose(a, c, rtol=1e-9, atol=1e-14)
    # forward: 27 waves of (2 results, 4 products); backward: 26 sweep steps of (2 results, 3 products)
    assert Bw.groups.count((2, 4, 0)) == 27 and Bw.groups.count((2, 3, 0)) == 26, Bw.groups
    assert Bp.groups.count((2, 3, 0)) == 26 and Bp.groups.count((2, 4, 0)) == 0

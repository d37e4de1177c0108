"""Objective functions transcribed from the reference's own test-suite, written
against the carrier method surface so the same program runs on the oracle
(CPU tests) and on the device backend (GPU tests).

Sources (relative to the horde-ad tree):
  test/simplified/TestGatherSimplified.hs:142-398   gather programs
  test/simplified/TestGatherSimplified.hs:961-1232  scatter programs
  test/simplified/TestGatherSimplified.hs:1262-1288 barRelu
  test/simplified/TestConvSimplified.hs:183-297     conv gradients
  test/tool/CrossTesting.hs:457-473                 fixtures t16b, t128, t128b, t128c
"""
import numpy as np

from horde_ad_b200 import ix as ixm, nn

minH = ixm.minH

T16B = np.array([5, 2, 6, 1, -2, 0, 0.1, -0.2, 13.1, 9, 8, -4, 582934, 2.99432, -335, 26],
                dtype=np.float64).reshape(2, 2, 2, 2)
T128 = np.array([29.1,32.1,40.1,29.0,53.99432,97.1,58.8943200001,18.1,29.1,32.1,40.1,32.0,53.99432,97.1,25.8943200001, 5, 2, 6, 1, -2, 97.1,58.8943200001,97.1,55.8943200001,97.1,58.8943200001,18.1,29.1,32.1,40.1,32.1,32.1,40.1,53.0,53.99432, -0.00001, 0.1, -0.2, 13.1, 9, 8, -4, 29, 2.99432, -335, 26, 2, 2, 2, 2, -0.2,-0.2,-0.2,-0.2,25.0003,25.0003,25.0003,25.0003,-0.2,-0.2,-0.2,-0.2,25.0003,25.0003,25.0003,25.0003,40.1,8.0,11.0,-3.0,25.89432,28.79432,-39.09999999999997,25.8,40.1,8.0,11.0,-3.0,25.89432,28.79432,-19.09999999999997,25.8, 8.1,29.1,32.1,40.1,32.1,40.1,292.0,53.99432,97.1,55.8943200001,97.1,85.8943200001,97.1,85.8943200001,18.1,29.1,32.1,40.1,32.1,40.1,32.1,40.1,22.0,53.99432,97.1,82.8943200001,97.1,22.8943200001,97.1,58.8943200001,18.1,29.1,32.1,40.1,32.1,40.1,32.1,40.1,89.0,53.99432,97.1,56.8943200001,97.1,52.8943200001,97.1,55.8943200001],
                dtype=np.float64).reshape(1, 2, 2, 1, 2, 2, 2, 2, 2, 1)
T128B = T128.reshape(4, 2, 4, 4)
T128C = T128.reshape(2, 2, 8, 4)

# rreplicate 7 $ ringestData [2] [0, 1]
IN72 = np.tile(np.array([0.0, 1.0]), (7, 1))
# vals of the interpreter-consistency tests (TestGatherSimplified.hs:212)
VALS72 = np.array([-1, 0, 2.0, 5.0, 11.0, -17.0, 23.0, 29.0, -35.0, 41.0, 47.0, 33.0, 0.1, 0.007]).reshape(7, 2)


# ---- gathers (TestGatherSimplified.hs:142-398) ---------------------------
def gatherNested1(T, t):
    inner = T.sgather([4], t, lambda i: [i[0]], 1)
    return T.sgather([2], inner, lambda i: [i[0] + i[0], i[0]], 2)


def gather1(T, t):
    return T.sgather([2], T.sslice(0, 4, t), lambda i: [i[0] + i[0], i[0]], 2)


def gatherNested02(T, t):
    inner = T.sgather([2], t, lambda k: [k[0] + k[0]], 1)
    return T.sgather([1], inner, lambda i: [i[0] + i[0] + i[0]], 1)


def gatherNested2(T, t):
    inner = T.sgather([2, 3, 4], t, lambda k: [k[0] + k[1] + k[2]], 1)
    return T.sgather([2, 3], inner, lambda i: [i[0], i[1], i[0] + i[1], i[0]], 4)


def gather2(T, t):
    return T.sgather([2, 3], t, lambda i: [i[0] + i[1] + i[0] + i[1], i[0]], 2)


def gatherNested12(T, t):
    inner = T.sgather([2, 3, 4], t, lambda k: [k[0] + k[1] + k[2], k[0]], 2)
    return T.sgather([2], inner, lambda i: [i[0], i[0] + i[0]], 2)


def gather12(T, t):
    return T.sgather([2, 4], t, lambda i: [i[0] + i[0] + i[0] + i[1], i[0]], 2)


# ---- scatters (TestGatherSimplified.hs:961-1232) -------------------------
def scatterNested1(T, t):
    inner = T.sscatter([7], t, lambda k: [k[0]], 1)
    return T.sscatter([2], inner, lambda i: [i[1].quotH(1 + i[0])], 2)


def scatter1(T, t):
    return T.sscatter([2], t, lambda i: [minH(i[1] + 2 * i[0], 1)], 2)


def scatterNested2(T, t):
    inner = T.sscatter([2, 3, 4], t, lambda k: [minH(k[0], 1), minH(k[0], 2), minH(k[0], 3)], 1)
    return T.sscatter([2, 3], inner, lambda i: [minH(i[0] + i[1], 1), minH(i[3] + i[0], 2)], 4)


def scatter2(T, t):
    return T.sscatter([2, 3], t, lambda i: [minH(i[0] + i[1] + i[0] + i[1], 1), minH(i[0], 2)], 2)


def scatterNested12(T, t):
    inner = T.sscatter([2, 3, 4], t, lambda k: [minH(k[0], 1), minH(k[1] + k[0], 2), minH(k[0], 3)], 2)
    return T.sscatter([2], inner, lambda i: [minH(i[0] + i[0], 1)], 2)


def scatter12(T, t):
    return T.sscatter([2, 4], t, lambda i: [minH(i[0] + i[0] + i[0] + i[1], 1), minH(i[0], 3)], 2)


ONES72 = [1.0] * 14
GATHER_GOLDEN = [
    # (name, program, input, expected gradient wrt the input, dt = ones)
    ("gatherNested1", gatherNested1, IN72, [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0]),
    ("gather1", gather1, IN72, [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0]),
    ("gatherNested02", gatherNested02, np.full(4, 0.1), [1, 0, 0, 0]),
    ("gatherNested2", gatherNested2, IN72, [1, 0, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 0, 1]),
    ("gather2", gather2, IN72, [1, 0, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 0, 1]),
    ("gatherNested12", gatherNested12, IN72, [1, 0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 1]),
    ("gather12", gather12, IN72, [1, 0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 1]),
    ("scatterNested1", scatterNested1, IN72, ONES72),
    ("scatter1", scatter1, IN72, ONES72),
    ("scatterNested2", scatterNested2, IN72, ONES72),
    ("scatter2", scatter2, IN72, ONES72),
    ("scatterNested12", scatterNested12, IN72, ONES72),
    ("scatter12", scatter12, IN72, ONES72),
]

# pairs the reference asserts to be @?=-equal (bit-exact) on VALS72
# (TestGatherSimplified.hs:208-235, 302-345): nested == fused-by-hand
GATHER_EQUAL_PAIRS = [(gatherNested1, gather1), (gatherNested2, gather2), (gatherNested12, gather12)]


# ---- barRelu (TestGatherSimplified.hs:1262-1288) -------------------------
def _foo(T, x, y, z):
    w = T.binary("mul", x, T.unary("sin", y))
    return T.binary("add", T.binary("atan2", z, w), T.binary("mul", z, w))


def _bar(T, x, y):
    w = T.binary("mul", _foo(T, x, y, x), T.unary("sin", y))
    return T.binary("add", T.binary("atan2", x, w), T.binary("mul", y, w))


def barRelu(T, x):
    from horde_ad_b200 import nn
    sh = list(T.shape(x))
    t = T.binary("mul", T.srepl(sh, 0.001), x)
    return nn.reluS(T, _bar(T, t, nn.reluS(T, t)))


BARRELU_GOLDEN = np.array([2.885038541771792e-4,2.885145151321922e-4,2.8854294397024206e-4,2.885034988157713e-4,2.885923176600045e-4,2.887454843457817e-4,2.886097295122454e-4,2.8846476339094805e-4,2.885038541771792e-4,2.885145151321922e-4,2.8854294397024206e-4,2.8851415976532735e-4,2.885923176600045e-4,2.887454843457817e-4,2.8849246223035154e-4,2.884182085399516e-4,2.884075468755327e-4,2.8842176240868867e-4,2.8840399312321096e-4,0.0,2.887454843457817e-4,2.886097295122454e-4,2.887454843457817e-4,2.88599069218435e-4,2.887454843457817e-4,2.886097295122454e-4,2.8846476339094805e-4,2.885038541771792e-4,2.885145151321922e-4,2.8854294397024206e-4,2.885145151321922e-4,2.885145151321922e-4,2.8854294397024206e-4,2.8858878438222746e-4,2.885923176600045e-4,0.0,2.884007943794131e-4,0.0,2.884469945274759e-4,2.8843242392031246e-4,2.884288700806792e-4,0.0,2.885034988157713e-4,2.884110805753153e-4,0.0,2.8849283778617973e-4,2.884075468755327e-4,2.884075468755327e-4,2.884075468755327e-4,2.884075468755327e-4,0.0,0.0,0.0,0.0,2.884892851579934e-4,2.884892851579934e-4,2.884892851579934e-4,2.884892851579934e-4,0.0,0.0,0.0,0.0,2.884892851579934e-4,2.884892851579934e-4,2.884892851579934e-4,2.884892851579934e-4,2.8854294397024206e-4,2.884288700806792e-4,2.884395315486472e-4,0.0,2.8849246223035154e-4,2.8850276789489724e-4,0.0,2.8849212704517413e-4,2.8854294397024206e-4,2.884288700806792e-4,2.884395315486472e-4,0.0,2.8849246223035154e-4,2.8850276789489724e-4,0.0,2.8849212704517413e-4,2.8842922547482884e-4,2.885038541771792e-4,2.885145151321922e-4,2.8854294397024206e-4,2.885145151321922e-4,2.8854294397024206e-4,2.894378297730782e-4,2.885923176600045e-4,2.887454843457817e-4,2.88599069218435e-4,2.887454843457817e-4,2.887056688523444e-4,2.887454843457817e-4,2.887056688523444e-4,2.8846476339094805e-4,2.885038541771792e-4,2.885145151321922e-4,2.8854294397024206e-4,2.885145151321922e-4,2.8854294397024206e-4,2.885145151321922e-4,2.8854294397024206e-4,2.884786229769816e-4,2.885923176600045e-4,2.887454843457817e-4,2.886950092188272e-4,2.887454843457817e-4,2.884818011261814e-4,2.887454843457817e-4,2.886097295122454e-4,2.8846476339094805e-4,2.885038541771792e-4,2.885145151321922e-4,2.8854294397024206e-4,2.885145151321922e-4,2.8854294397024206e-4,2.885145151321922e-4,2.8854294397024206e-4,2.887167039107226e-4,2.885923176600045e-4,2.887454843457817e-4,2.8860262265516213e-4,2.887454843457817e-4,2.885884088500461e-4,2.887454843457817e-4,2.88599069218435e-4])

# ---- conv gradients (TestConvSimplified.hs:183-297) -----------------------
# (name, kernel K or None, input A or None, differentiate wrt, expected, eps)
CONV_GOLDEN = [
    ("KonstG0Rev", T16B, np.zeros((2, 2, 2, 2)), "A",
     [18.1,29.1,32.1,40.1,582932.0,582934.99432,582597.1,582625.8943200001,18.1,29.1,32.1,40.1,582932.0,582934.99432,582597.1,582625.8943200001], 1e-4),
    ("KonstG0Tiny1", np.array([-0.2]).reshape(1, 1, 1, 1), np.zeros((1, 1, 1, 1)), "A", [-0.2], 1e-10),
    ("KonstG0TinyS", np.full((1, 1, 1, 1), T16B.sum()), np.zeros((1, 1, 1, 1)), "A", [582665.99432], 1e-10),
    ("KonstG0TinyA", np.array([-0.2, 25.0003]).reshape(1, 2, 1, 1), np.zeros((1, 2, 1, 1)), "A", [-0.2, 25.0003], 1e-10),
    ("KonstG0LittleA", np.array([-0.2, 25.0003]).reshape(1, 2, 1, 1), np.zeros((2, 2, 2, 2)), "A",
     [-0.2,-0.2,-0.2,-0.2,25.0003,25.0003,25.0003,25.0003,-0.2,-0.2,-0.2,-0.2,25.0003,25.0003,25.0003,25.0003], 1e-10),
    ("Konst5LittleB", T16B, np.full((2, 2, 2, 2), 5.0), "A",
     [18.1,29.1,32.1,40.1,582932.0,582934.99432,582597.1,582625.8943200001,18.1,29.1,32.1,40.1,582932.0,582934.99432,582597.1,582625.8943200001], 1e-8),
    ("Konst5LittleC", np.full((2, 2, 2, 2), 5.0), T16B, "K",
     [40.1,8.0,11.0,-3.0,582625.89432,28.79432,-309.09999999999997,25.8,40.1,8.0,11.0,-3.0,582625.89432,28.79432,-309.09999999999997,25.8], 1e-8),
]


def _conv_goldens_from_file():
    """tests/golden/conv_goldens.json (written by tests/golden/make_conv_goldens.py from the literals of
    TestConvSimplified.hs:196-470): conv2dPreserving over t16b / t128b / t128c held constant as the kernel
    (conv2dB*: gradient wrt the input) or as the input (conv2dC* = flip conv2dPreserving: gradient wrt the
    kernel), at a constant point or at [37, 36 .. -10]."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conv_goldens.json")
    fixtures = {"t16b": T16B, "t128b": T128B, "t128c": T128C}
    have = {c[0] for c in CONV_GOLDEN}
    out = []
    for t in json.load(open(path))["tests"]:
        name = t["name"][len("test"):]
        if name in have:
            continue
        pt = t["point"]
        if pt["kind"] == "constant":
            x = np.full(pt["shape"], pt["value"], dtype=np.float64)
        else:
            x = np.arange(37, -11, -1, dtype=np.float64).reshape(pt["shape"])
        fx = fixtures[t["fixture"]]
        K, A = (fx, x) if t["wrt"] == "A" else (x, fx)
        out.append((name, K, A, t["wrt"], t["expected"], t["eps"]))
    return out


CONV_GOLDEN = CONV_GOLDEN + _conv_goldens_from_file()


# ---------------------------------------------------------------------------
# testCNNOPP3 (test/simplified/TestGatherSimplified.hs:1365-1436,1574-1578): the toy "convolution" and
# "max-pool" written with rbuild, rindex and ifH whose constant-folded value the reference prints as
#   sfromListLinear [2,2,2,2] [14,0,14,0,0,...,0]
# on a [3,3,3,3] glyph of sevens.  Every index may run out of bounds (then `!` is 0) and the guard
# `within3` keeps only indices that are 0 or dim - 2.  Non-vectorised code by construction: host loops.
def _within3(sh, ix):
    return all(i == 0 or d - 2 == i for i, d in zip(ix, sh))


def _indexz03(T, d, ix):
    zero = T.srepl([], 0.0, T.dtype(d))
    return nn.scond(T, _within3(T.shape(d), ix), T.sindex(d, list(ix)), zero)


def _slicez4(T, sh_out, d, ix_base):
    return nn.sbuild(T, sh_out, lambda ixr: _indexz03(T, d, [a + b for a, b in zip(ix_base, ixr)]))


def _slicez3(T, sh_out, d, ix_base):
    return nn.sbuild(T, sh_out, lambda _ixr: _indexz03(T, d, [a + a for a in ix_base]))


def conv2dPreserving3(T, arrA):
    def at(ix):
        iImg, _, iBh, iBw = ix
        arrAt = _slicez4(T, [2, 2, 2, 2], arrA, [iImg, 0, iBw, 1])
        return T.binary("add", T.sindex(arrAt, [0, iBw, iImg, iBh]), T.sindex(arrAt, [iImg, 1, iBw + 1, iBh]))
    return nn.sbuild(T, [2, 2, 2, 2], at)


def maxPool2dUnpadded3(T, arr):
    def at(ix):
        aa, bb, iBh, iBw = ix
        arrt = _slicez3(T, [2, 2, 2, 2], arr, [iBh // 2, aa, bb, iBw])
        return T.sindex(arrt, [0, 0, 0, 0])
    return nn.sbuild(T, [2, 2, 2, 2], at)


CNNOPP3_GOLDEN = [14.0, 0.0, 14.0, 0.0] + [0.0] * 12


# testCNNOPP6 / testCNNOPP7 (TestGatherSimplified.hs:1580-1650): the same helpers with every base index
# doubled (slicez3); folded values [7,0,7,0,0,...] and [7,0,0,...] on a [2,2,2,2] glyph of sevens.
def conv2dPreserving3z(T, arrA):
    def at(ix):
        iImg, _, iBh, iBw = ix
        arrAt = _slicez3(T, [2, 2, 2, 2], arrA, [iImg, iImg, iImg, iBw])
        return T.sindex(arrAt, [iBh, iBw, iImg, iBh])
    return nn.sbuild(T, [2, 2, 2, 2], at)


def conv2dPreserving3y(T, arrA):
    def at(ix):
        iImg, _, iBh, iBw = ix
        arrAt = _slicez3(T, [2, 2, 2, 2], arrA, [iImg, iImg, iImg, iBh])
        return T.sindex(arrAt, [iBh, iBw, iImg, iBh])
    return nn.sbuild(T, [2, 2, 2, 2], at)


def maxPool2dUnpadded3y(T, arr):
    def at(ix):
        aa, bb, iBh, iBw = ix
        arrt = _slicez3(T, [2, 2, 2, 2], arr, [iBh, aa, bb, iBw])
        return T.sindex(arrt, [0, 0, 0, 0])
    return nn.sbuild(T, [2, 2, 2, 2], at)


CNNOPP6_GOLDEN = [7.0, 0.0, 7.0, 0.0] + [0.0] * 12
CNNOPP7_GOLDEN = [7.0] + [0.0] * 15


# ---------------------------------------------------------------------------
# vstackABC / vstackBuild (test/simplified/TestAdaptorSimplified.hs:785-813) and their known answers
# (:815-835, :949): a[0]+b[1] | a[1:n-1]+b[2:n]+c[:n-2] | a[n-1]+c[n-2], once with slices and appends,
# once as rbuild1 over host conditionals.
def vstackABC(T, a, b, c):
    n = T.shape(a)[0]
    add = lambda x, y: T.binary("add", x, y)
    head = T.sreplicate(1, add(T.sindex(a, [0]), T.sindex(b, [1])))
    mid = add(T.sslice(1, n - 2, a), add(T.sslice(2, n - 2, b), T.sslice(0, n - 2, c)))
    last = T.sreplicate(1, add(T.sindex(a, [n - 1]), T.sindex(c, [n - 2])))
    return T.sappend(head, T.sappend(mid, last))


def vstackBuild(T, a, b, c):
    n = T.shape(a)[0]
    add = lambda x, y: T.binary("add", x, y)

    def at(i):
        if i == 0:
            return add(T.sindex(a, [0]), T.sindex(b, [1]))
        if i == n - 1:
            return add(T.sindex(a, [n - 1]), T.sindex(c, [n - 2]))
        return add(add(T.sindex(a, [i]), T.sindex(b, [i + 1])), T.sindex(c, [i - 1]))
    return nn.sbuild1(T, n, at)


VSTACK_GOLDEN = [
    ("repl", lambda n: (np.full(n, 1.0), np.full(n, 2.0), np.full(n, 3.0)), [3.0, 6.0, 6.0, 6.0, 6.0, 6.0, 6.0, 6.0, 6.0, 4.0]),
    ("iota", lambda n: (np.arange(n, dtype=float),) * 3, [1.0, 3.0, 6.0, 9.0, 12.0, 15.0, 18.0, 21.0, 24.0, 17.0]),
    ("replIota", lambda n: (1.0 * np.arange(n), 2.0 * np.arange(n), 3.0 * np.arange(n)),
     [2.0, 5.0, 11.0, 17.0, 23.0, 29.0, 35.0, 41.0, 47.0, 33.0]),
]


# ---------------------------------------------------------------------------
# testGatherTranspose33 / testGatherTransposeBuild33 (test/simplified/TestGatherSimplified.hs:564-606):
# 21 transpositions (several with permutations shorter than the rank) and two reshapes of the rank-10
# fixture t128, then a matmul with the fixture t48 (test/tool/CrossTesting.hs:463-464); the 128-element
# gradients are the reference's literals.  Pins the permutation convention of tstranspose.
T48 = np.array([18.1, 29.1, 32.1, 40.1, 52.0, 53.99432, 97.1, 58.8943200001, 18.1, 29.1, 32.1, 40.1, 58.0, 54.99432, 97.1,
                52.8943200001, 5, 2, 6, 1, -2, 0.92, 0.1, -0.2, 13.1, 9, 8, -4, 34, 2.99432, -33, 26, 2, 2, 2, 2, -0.2, -0.2,
                -0.2, -0.2, 25.0003, -0.2, -0.2, -0.2, 25.0003, 25.0003, 25.0003, 25.0003]).reshape(3, 1, 2, 2, 1, 2, 2)


def gatherTranspose33(T, t):
    x = T.stranspose([0, 1, 7, 4, 5, 3, 6, 2, 8], t)
    for perm in ([1, 0], [0, 1], [0], [], [0, 1, 2, 4, 3, 5, 6, 7, 9, 8], [5, 0, 1, 2, 3, 4], [0, 1, 2, 3, 4, 5, 6],
                 [0, 1, 2, 3, 7, 5, 6, 4], [0, 1, 2, 3, 4, 5, 6, 7, 9, 8], [1, 0, 3, 2], [1, 2, 3, 0], [0, 1, 2, 3]):
        x = T.stranspose(perm, x)
    x = T.sreshape([2, 2, 8, 4], x)
    for perm in ([3, 0, 2, 1], [1, 2, 3, 0], [0, 1, 2, 3], [1, 0], [1, 0, 2], [1, 2, 0], [2, 0, 1], [0, 1, 2]):
        x = T.stranspose(perm, x)
    x = T.sreshape([16, 8], x)
    return T.smatmul2(T.sreshape([6, 8], T.sconcrete(T48)), T.stranspose([1, 0], x))


def gatherTransposeBuild33(T, t):
    """rbuild1 4 (\\i -> gatherTranspose33 (t * replicate0N i))"""
    sh = list(T.shape(t))
    return nn.sbuild1(T, 4, lambda i: gatherTranspose33(T, T.binary("mul", t, T.srepl(sh, float(i), T.dtype(t)))))


GT33_GOLDEN = [
    81.3003, 71.0, 81.3003, 71.0, 81.3003, 71.0, 81.3003, 71.0, 80.0, 79.0, 80.0, 79.0, 80.0, 79.0, 80.0, 79.0,
    81.3003, 71.0, 81.3003, 71.0, 81.3003, 71.0, 81.3003, 71.0, 80.0, 79.0, 80.0, 79.0, 80.0, 79.0, 80.0, 79.0,
    81.3003, 71.0, 81.3003, 71.0, 81.3003, 71.0, 81.3003, 71.0, 80.0, 79.0, 80.0, 79.0, 80.0, 79.0, 80.0, 79.0,
    81.3003, 71.0, 81.3003, 71.0, 81.3003, 71.0, 81.3003, 71.0, 80.0, 79.0, 80.0, 79.0, 80.0, 79.0, 80.0, 79.0,
    166.8003, 137.70326, 166.8003, 137.70326, 166.8003, 137.70326, 166.8003, 137.70326, 186.1003, 162.3889400002,
    186.1003, 162.3889400002, 186.1003, 162.3889400002, 186.1003, 162.3889400002, 166.8003, 137.70326, 166.8003,
    137.70326, 166.8003, 137.70326, 166.8003, 137.70326, 186.1003, 162.3889400002, 186.1003, 162.3889400002,
    186.1003, 162.3889400002, 186.1003, 162.3889400002, 166.8003, 137.70326, 166.8003, 137.70326, 166.8003,
    137.70326, 166.8003, 137.70326, 186.1003, 162.3889400002, 186.1003, 162.3889400002, 186.1003, 162.3889400002,
    186.1003, 162.3889400002, 166.8003, 137.70326, 166.8003, 137.70326, 166.8003, 137.70326, 166.8003, 137.70326,
    186.1003, 162.3889400002, 186.1003, 162.3889400002, 186.1003, 162.3889400002, 186.1003, 162.3889400002,
]
GT33_BUILD_GOLDEN = [
    487.80179999999996, 426.0, 487.80179999999996, 426.0, 487.80179999999996, 426.0, 487.80179999999996, 426.0,
    480.0, 474.0, 480.0, 474.0, 480.0, 474.0, 480.0, 474.0, 487.80179999999996, 426.0, 487.80179999999996, 426.0,
    487.80179999999996, 426.0, 487.80179999999996, 426.0, 480.0, 474.0, 480.0, 474.0, 480.0, 474.0, 480.0, 474.0,
    487.80179999999996, 426.0, 487.80179999999996, 426.0, 487.80179999999996, 426.0, 487.80179999999996, 426.0,
    480.0, 474.0, 480.0, 474.0, 480.0, 474.0, 480.0, 474.0, 487.80179999999996, 426.0, 487.80179999999996, 426.0,
    487.80179999999996, 426.0, 487.80179999999996, 426.0, 480.0, 474.0, 480.0, 474.0, 480.0, 474.0, 480.0, 474.0,
    1000.8018, 826.21956, 1000.8018, 826.21956, 1000.8018, 826.21956, 1000.8018, 826.21956, 1116.6018,
    974.3336400012, 1116.6018, 974.3336400012, 1116.6018, 974.3336400012, 1116.6018, 974.3336400012, 1000.8018,
    826.21956, 1000.8018, 826.21956, 1000.8018, 826.21956, 1000.8018, 826.21956, 1116.6018, 974.3336400012,
    1116.6018, 974.3336400012, 1116.6018, 974.3336400012, 1116.6018, 974.3336400012, 1000.8018, 826.21956,
    1000.8018, 826.21956, 1000.8018, 826.21956, 1000.8018, 826.21956, 1116.6018, 974.3336400012, 1116.6018,
    974.3336400012, 1116.6018, 974.3336400012, 1116.6018, 974.3336400012, 1000.8018, 826.21956, 1000.8018,
    826.21956, 1000.8018, 826.21956, 1000.8018, 826.21956, 1116.6018, 974.3336400012, 1116.6018, 974.3336400012,
    1116.6018, 974.3336400012, 1116.6018, 974.3336400012,
]


# ---------------------------------------------------------------------------
# gatherCond / gatherCond2 (test/simplified/TestGatherSimplified.hs:716-770; the tests are parked behind a
# "re-enable once we drop GHC 9.10" comment, their expected gradients are still stated): a conditional
# `ifH (i ==. 3) 0 j` inside the index function of a gather.
def gatherCond(T, u):
    w = T.shape(u)[0]
    v = T.stranspose([2, 0, 1], T.sreplicate(2 * w, u))
    return T.sgather([w, 2], v, lambda ix: [ixm.ifH(ix[0].eq(3), 0, ix[1]), 2 * ix[0], ix[0]], 3)


def gatherCond2(T, u):
    w = T.shape(u)[0]
    v = T.sreplicate(2 * w, u)
    return T.stranspose([1, 0], T.sgather([2, w], v, lambda ix: [2 * ix[1], ix[1], ixm.ifH(ix[1].eq(3), 0, ix[0])], 3))


GATHER_COND_GOLDEN = [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 2.0, 0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]


# gatherCond3 .. gatherCond6 (TestGatherSimplified.hs:795-915): the same with index components that run
# out of bounds (2 * i over a short axis): out-of-bounds elements read 0 and receive no cotangent.
def gatherCond3(T, u):
    w = T.shape(u)[0]
    v = T.stranspose([2, 0, 1], T.sreplicate(2 * w, u))
    return T.sgather([w, 2], v, lambda ix: [2 * ix[0], ix[0], ixm.ifH(ix[0].eq(3), 0, ix[1])], 3)


def gatherCond4(T, u):
    w = T.shape(u)[0]
    v = T.sreplicate(2 * w, u)
    return T.stranspose([1, 0], T.sgather([2, w], v, lambda ix: [ix[1], ixm.ifH(ix[1].eq(3), 0, ix[0]), 2 * ix[1]], 3))


def gatherCond5(T, v):
    return T.sgather([T.shape(v)[0], 2], v, lambda ix: [ixm.ifH(ix[0].eq(1), 0, ix[1]), 2 * ix[0], ix[0]], 3)


def gatherCond6(T, u):
    v = T.stranspose([2, 0, 1], u)
    return T.stranspose([1, 0], T.sgather([2, T.shape(v)[0]], v,
                                          lambda ix: [ix[1], ixm.ifH(ix[1].eq(1), 0, ix[0]), 2 * ix[1]], 3))


GATHER_COND34_GOLDEN = [1.0, 0.0, 1.0, 0.0] + [0.0] * 10                  # input [7, 2]
GATHER_COND56_GOLDEN = [1.0, 0.0, 0.0, 0.0, 0.0, 2.0, 0.0, 0.0, 1.0] + [0.0] * 7   # input [2, 4, 2]


# ---------------------------------------------------------------------------
# The rbuild1 variants of the gather / scatter tests (TestGatherSimplified.hs:161-398, 968-1232):
#   build1:  rbuild1 5 (\i -> ifH (i >. 2) (prog t) (t ! [i]))
#   build2:  rbuild1 4 (\i -> prog (t * replicate0N i))              -> 6 x the plain gradient
#   build12: rbuild1 5 (\i -> ifH (i >. 2) (prog t) (rtr (rreplicate 4 (t ! [i])))) ! [1]
def build1_variant(prog):
    return lambda T, t: nn.sbuild1(T, 5, lambda i: prog(T, t) if i > 2 else T.sindex(t, [i]))


def build2_variant(prog):
    return lambda T, t: nn.sbuild1(T, 4, lambda i: prog(T, T.binary("mul", t, T.srepl(list(T.shape(t)), float(i), T.dtype(t)))))


def build12_variant(prog):
    return lambda T, t: T.sindex(nn.sbuild1(
        T, 5, lambda i: prog(T, t) if i > 2 else T.stranspose([1, 0], T.sreplicate(4, T.sindex(t, [i])))), [1])


_G1B = [3.0, 1.0, 1.0, 1.0, 1.0, 3.0] + [0.0] * 8
_S1B = [3.0] * 6 + [2.0] * 8
_B12 = [0.0, 0.0, 4.0, 4.0] + [0.0] * 10
_G2 = [1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0]
GATHER_BUILD_GOLDEN = [
    ("gatherNestedBuild1", build1_variant(gatherNested1), _G1B), ("gatherBuild1", build1_variant(gather1), _G1B),
    ("scatterNestedBuild1", build1_variant(scatterNested1), _S1B), ("scatterBuild1", build1_variant(scatter1), _S1B),
    ("gatherNestedBuild2", build2_variant(gatherNested2), [6 * x for x in _G2]),
    ("gatherBuild2", build2_variant(gather2), [6 * x for x in _G2]),
    ("scatterNestedBuild2", build2_variant(scatterNested2), [6.0] * 14), ("scatterBuild2", build2_variant(scatter2), [6.0] * 14),
    ("gatherNestedBuild12", build12_variant(gatherNested12), _B12), ("gatherBuild12", build12_variant(gather12), _B12),
    ("scatterNestedBuild12", build12_variant(scatterNested12), _B12), ("scatterBuild12", build12_variant(scatter12), _B12),
]


# ---- relu / reluMax / barRelu known answers (TestAdaptorSimplified.hs:1247-1251, 1300-1345, 1404-1462,
#      1836-1925): relu = oneIfGtZero * v (CommonRankedOps.hs:40-49), reluMax = rmap0N (maxH 0) with
#      maxH u v = ifH (u >=. v) u v (OpsTensor.hs:327-329) as a per-element host function (tsmap0N),
#      bar (x, y) = let w = foo2 (x, y, x) * sin y in atan2H x w + y * w (:1585-1588) ----------------------
RELU_IN = np.array([1.1, -2.2, 0, 4.4, 5.5, 6.6, 7.7, 8.8, -9.9, -10, 11, 12]).reshape(3, 4)
RELU_MASK = np.array([1.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 1.0, 1.0]).reshape(3, 4)
BAR_IN = np.array([1.1, 2, 3, 4.2]).reshape(2, 1, 2)
BAR_GRAD = np.array([4.5309153191767395, 4.5302138998556, -9.39547533946234, 95.29759282497125]).reshape(2, 1, 2)
BAR_TAN = np.array([0.1, 0.2, 0.3, 0.42]).reshape(2, 1, 2)
BAR_FWD = np.array([0.45309153191767404, 0.9060427799711201, -2.8186426018387007, 40.02498898648793]).reshape(2, 1, 2)


def relu_(T, v):
    from horde_ad_b200 import nn
    return nn.reluS(T, v)


def reluMax(T, v):
    from horde_ad_b200 import nn
    zero = lambda: T.sscalar(0.0)
    return nn.smap0N(T, lambda s: zero() if 0.0 >= float(T.kfromS(s)) else s, v)


def reluT2(act):
    """reluT2 (t, r) = act (t * rreplicate 3 (rreplicate 4 r))  (:1300-1306, 1327-1336, 1451-1459)."""
    def f(T, t, r):
        return act(T, T.binary("mul", t, T.sreplicate0N([3, 4], r)))
    return f


def barRelu0(T, x):
    return relu_(T, _bar(T, x, relu_(T, x)))


def barReluMax(T, x):
    return reluMax(T, _bar(T, x, reluMax(T, x)))


def barReluMax0(T, x):
    return reluMax(T, _bar(T, x, x))


def fooBuild1(T, v):
    """fooBuild1 (TestAdaptorSimplified.hs:1690-1700): rbuild1 3 over rsum0, rminimum and rindex v [ix + 1]."""
    from horde_ad_b200 import nn
    r = T.ssum0(v)
    vmin = nn.sminimum(T, v)
    three, five = T.sscalar(3.0), T.sscalar(5.0)

    def body(ix):
        a = T.binary("mul", r, _foo(T, three, T.binary("mul", five, r),
                                    T.binary("mul", T.binary("mul", r, nn.sminimum(T, v)), vmin)))
        return T.binary("add", a, _bar(T, r, T.sindex(v, [ix + 1])))
    return nn.sbuild1(T, 3, body)


def nestedBuildIndex(T, v):
    """nestedBuildIndex (TestAdaptorSimplified.hs:1826-1830): three nested rbuild1 / rindex pairs."""
    from horde_ad_b200 import nn
    inner = lambda: nn.sbuild1(T, 4, lambda i4: T.sindex(v, [i4]))
    mid = lambda: nn.sbuild1(T, 3, lambda i3: T.sindex(inner(), [i3]))
    return nn.sbuild1(T, 2, lambda i2: T.sindex(mid(), [i2]))


FOOBUILD_IN = np.array([1.1, 2.2, 3.3, 4])
# (name, f, inputs, expected gradients) -- gradient of the SUM of the result (rev' / grad (rsum0 . f)), eps 1e-10
ADAPTOR_REV_GOLDEN = [
    ("testReluSimpler", relu_, [RELU_IN], [RELU_MASK]),
    ("testReluSimpler3", reluT2(relu_), [RELU_IN, np.array(7.0)], [7.0 * RELU_MASK, np.array(57.1)]),
    ("testReluMax", reluMax, [RELU_IN], [RELU_MASK]),
    ("testReluMax3", reluT2(reluMax), [RELU_IN, np.array(7.0)], [7.0 * RELU_MASK, np.array(57.1)]),
    ("testBarRelu", barRelu0, [np.array(1.1)], [np.array(4.5309153191767395)]),
    ("testBarRelu3", barRelu0, [BAR_IN], [BAR_GRAD]),
    ("testBarReluMax", barReluMax, [np.array(1.1)], [np.array(4.5309153191767395)]),
    ("testBarReluMax30", barReluMax0, [np.array([1.1])], [np.array([4.5309153191767395])]),
    ("testBarReluMax31", barReluMax, [BAR_IN], [BAR_GRAD]),
    ("testFooBuild", fooBuild1, [FOOBUILD_IN],                                      # :1724-1728
     [np.array([-4521.201512195087, -5568.7163677622175, -5298.386349932494, -4907.349735554627])]),
    ("testNestedBuildIndex", nestedBuildIndex, [np.array([1.1, 2.2, 3.3, 4, -5.22])],   # :1832-1836
     [np.array([1.0, 1, 0, 0, 0])]),
]
# vjp with an incoming cotangent (:1843-1848, 1872-1877): (name, f, input, dt, expected)
ADAPTOR_VJP_GOLDEN = [
    ("testBarReluDt", barRelu0, np.array(1.1), np.array(42.2), np.array(191.20462646925841)),
    ("testBarReluMaxDt", barReluMax, np.array(1.1), np.array(42.2), np.array(191.20462646925841)),
    ("testFooBuildDt", fooBuild1, FOOBUILD_IN, np.full(3, 42.0),                      # :1703-1708 (eps 1e-5)
     np.array([-189890.46351219364, -233886.08744601303, -222532.22669716467, -206108.68889329425])),
]
# forward mode (:1896-1925): (name, f, input, tangent, expected)
ADAPTOR_FWD_GOLDEN = [
    ("testBarReluMax3CFwd", barReluMax, BAR_IN, BAR_TAN, BAR_FWD),
    ("testBarReluMax3FwdS", barReluMax, BAR_IN, BAR_TAN, BAR_FWD),
    ("testFooBuildFwd", fooBuild1, FOOBUILD_IN, np.full(4, 42.0),                     # :1710-1722 (eps 1e-5)
     np.array([-296584.8166864211, -290062.472288043, -265770.1775742018])),
]


# ---- rscan known answers (TestRevFwdFold.hs:934-1000): rev' = gradient of the sum of the stacked accumulators ----
def _rtr(T, t):
    r = len(T.shape(t))
    return T.stranspose([1, 0] + list(range(2, r)), t)


def sin0Scan1(T, x0):      # rscan (\x _a -> sin x) x0 (rrepl [1] 42)
    from horde_ad_b200 import nn
    return nn.sscan(T, lambda x, _a: T.unary("sin", x), x0, T.srepl([1], 42.0))


def sin0Scan2(T, x0):      # ... (rrepl [5] 42)
    from horde_ad_b200 import nn
    return nn.sscan(T, lambda x, _a: T.unary("sin", x), x0, T.srepl([5], 42.0))


def sin0Scan3(T, a0):      # rscan (\_x a -> sin a) (rreplicate0N [1,1,1,1,1] 84) (rreplicate 3 a0)
    from horde_ad_b200 import nn
    return nn.sscan(T, lambda _x, a: T.unary("sin", a), T.srepl([1, 1, 1, 1, 1], 84.0), T.sreplicate(3, a0))


def sin0Scan4(T, a0):      # rscan (\x a -> atan2H (sin x) (sin a)) (2 * a0) (rreplicate 3 a0)
    from horde_ad_b200 import nn
    two = T.srepl([1, 1, 1, 1, 1], 2.0)
    return nn.sscan(T, lambda x, a: T.binary("atan2", T.unary("sin", x), T.unary("sin", a)),
                    T.binary("mul", two, a0), T.sreplicate(3, a0))


def sin0Scan6(T, a0):      # rscan (\x a -> rtr (rtr x + rreplicate 1 (rreplicate 2 a))) (rreplicate 2 (rreplicate 1 a0)) (rreplicate 2 a0)
    from horde_ad_b200 import nn
    step = lambda x, a: _rtr(T, T.binary("add", _rtr(T, x), T.sreplicate(1, T.sreplicate(2, a))))
    return nn.sscan(T, step, T.sreplicate(2, T.sreplicate(1, a0)), T.sreplicate(2, a0))


_X5 = np.full((1, 1, 1, 1, 1), 1.1)
SCAN_GOLDEN = [   # (name, f, input, expected gradient of the sum), eps 1e-10
    ("testSin0Scan1", sin0Scan1, _X5, np.full((1, 1, 1, 1, 1), 1.4535961214255773)),
    ("testSin0Scan2", sin0Scan2, _X5, np.full((1, 1, 1, 1, 1), 2.2207726343670955)),
    ("testSin0Scan3", sin0Scan3, _X5, np.full((1, 1, 1, 1, 1), 1.360788364276732)),
    ("testSin0Scan4", sin0Scan4, _X5, np.full((1, 1, 1, 1, 1), -0.4458209450295252)),
    ("testSin0Scan6", sin0Scan6, np.full((1, 1), 1.1), np.full((1, 1), 12.0)),
]


# ---- overleaf (TestAdaptorSimplified.hs:498-600): rsum (rbuild [50] (\[i] -> rindex v [i `remH` 28])), which the
#      reference prints as  ssum0 (sgather1 @50 v (\i -> [remH i 28]))  with the gradient artifact
#      sscatter1 @50 (sreplicate0N @[50] dret) (\i -> [remH i 28])  -- a gather with wrap-around and a scatter
#      with collisions, in Double and in every integer width ----
def overleaf(T, v):
    n = T.shape(v)[0]
    return T.ssum0(T.sgather([50], v, lambda ix: [ix[0].remH(n)], 1))


OVERLEAF_GRAD = [2.0] * 22 + [1.0] * 6


# ---- high-rank known answers (TestHighRankSimplified.hs:527-560, 696-712) over t16 / t48 (CrossTesting.hs:456-463) ----
T16 = np.array([5, 2, 6, 1, -2, 0.000001, 0.1, -0.2, 13.1, 9, 8, -4, 34, 2.99432, -33, 26]).reshape(2, 2, 1, 2, 2)
LOGISTIC5_GRAD = np.array([6.648056670790033e-3,0.10499358540350662,2.466509291359931e-3,0.19661193324148185,0.1049935854035065,0.2499999999999375,0.24937604019289197,0.24751657271185995,2.0452222584760427e-6,1.2337934976493025e-4,3.3523767075636815e-4,1.7662706213291118e-2,1.7763568394002473e-15,4.540945566439111e-2,4.6588861451033536e-15,5.109024314693943e-12]).reshape(2, 2, 1, 2, 2)
LOGISTIC52_GRAD = np.array([1.3111246391159124e-3,2.1750075272657612e-2,4.8549901151740267e-4,4.312916016242333e-2,2.6155373652699744e-2,5.8750924453083504e-2,5.8238333583278255e-2,5.8847120842749026e-2,4.0211548220008027e-7,2.425923569564766e-5,6.592194028602285e-5,4.415319450597324e-3,3.4925295232121135e-16,9.12284655835676e-3,1.1647215362758384e-15,1.0044951474920845e-12]).reshape(2, 2, 1, 2, 2)
BARRELU_T48_GRAD = np.array([3.513740871835189e-4,3.8830416352632824e-4,3.981974371104471e-4,4.2420226755643853e-4,4.6186212581292275e-4,4.6805323209889415e-4,5.933633926875981e-4,4.8311739820100107e-4,3.513740871835189e-4,3.8830416352632824e-4,3.981974371104471e-4,4.2420226755643853e-4,4.803836032226148e-4,4.7114455958615145e-4,5.933633926875981e-4,4.6464270870595213e-4,3.060675467148428e-4,2.954918864100193e-4,3.095763053673437e-4,2.9195025355591045e-4,0.0,2.9166656928452994e-4,2.887557883241243e-4,0.0,3.342503234557057e-4,3.2005299770444394e-4,3.165690448140097e-4,0.0,4.0442335110155446e-4,2.990052759764126e-4,0.0,3.780004614233832e-4,2.954918864100193e-4,2.954918864100193e-4,2.954918864100193e-4,2.954918864100193e-4,0.0,0.0,0.0,0.0,3.7466025157760897e-4,0.0,0.0,0.0,3.7466025157760897e-4,3.7466025157760897e-4,3.7466025157760897e-4,3.7466025157760897e-4]).reshape(3, 1, 2, 2, 1, 2, 2)
BARRELU_T48_SCALED_GRAD = np.array([2.8846476339094805e-4,2.885038541771792e-4,2.885145151321922e-4,2.8854294397024206e-4,2.885852309100301e-4,2.885923176600045e-4,2.887454843457817e-4,2.886097295122454e-4,2.8846476339094805e-4,2.885038541771792e-4,2.885145151321922e-4,2.8854294397024206e-4,2.8860655161315664e-4,2.88595871110374e-4,2.887454843457817e-4,2.885884088500461e-4,2.884182085399516e-4,2.884075468755327e-4,2.8842176240868867e-4,2.8840399312321096e-4,0.0,2.8840370860416445e-4,2.884007943794131e-4,0.0,2.884469945274759e-4,2.8843242392031246e-4,2.884288700806792e-4,0.0,2.885212670262263e-4,2.884110805753153e-4,0.0,2.8849283778617973e-4,2.884075468755327e-4,2.884075468755327e-4,2.884075468755327e-4,2.884075468755327e-4,0.0,0.0,0.0,0.0,2.884892851579934e-4,0.0,0.0,0.0,2.884892851579934e-4,2.884892851579934e-4,2.884892851579934e-4,2.884892851579934e-4]).reshape(3, 1, 2, 2, 1, 2, 2)
BARRELU_T16_DT_GRAD = np.array([1.2916050471365906e-2,1.2469757606504572e-2,1.3064120086501589e-2,1.2320300700062944e-2,0.0,1.217049789428711e-2,1.2185494267265312e-2,0.0,1.4105363649830907e-2,1.3506236503127638e-2,1.3359213691150671e-2,0.0,1.7066665416485535e-2,1.2618022646204737e-2,0.0,1.595161947206668e-2]).reshape(2, 2, 1, 2, 2)


# ---- rbuild / rindex / rfromList programs of TestHighRankSimplified.hs:557-690 (non-vectorised host loops) ----
def braidedBuilds(T, r):
    r"""rbuild [3, 4] (\[ix1, ix2] -> rfromList [rfromIndex0 ix2, 7, rsum0 (rslice 1 1 r), -0.2] ! [ix1])  (:559-563)."""
    from horde_ad_b200 import nn

    def f(ix):
        ix1, ix2 = ix
        lst = T.sfromVector([T.sscalar(float(ix2)), T.sscalar(7.0), T.ssum0(T.sslice(1, 1, r)), T.sscalar(-0.2)])
        return T.sindex(lst, [ix1])
    return nn.sbuild(T, [3, 4], f)


def concatBuild2(T, r):
    r"""rbuild1 5 (\i -> rbuild1 2 (\j -> rmap0N (* rfromIndex0 (maxH j (i `quotH` (j + 1)))) r))  (:672-676)."""
    from horde_ad_b200 import nn
    sh = list(T.shape(r))
    scale = lambda k: T.binary("mul", r, T.srepl(sh, float(k)))
    return nn.sbuild1(T, 5, lambda i: nn.sbuild1(T, 2, lambda j: scale(max(j, i // (j + 1)))))


def concatBuildm(T, r):
    r"""rbuild1 7 (\i -> rmap0N (* rfromIndex0 (kfloor (rsum0 (r ! [i])))) r)  (:640-645); r ! [i] is all zeros
    when i runs past the leading extent (OpsConcrete.hs:1448-1450)."""
    from horde_ad_b200 import nn
    sh = list(T.shape(r))

    def row(i):
        k = int(np.floor(float(T.kfromS(T.ssum0(T.sindex(r, [i]))))))
        return T.binary("mul", r, T.srepl(sh, float(k)))
    return nn.sbuild1(T, 7, row)


V7 = np.array([0.651, 0.14, 0.3414, -0.14, 0.0014, 0.0020014, 0.9999])
BUILD_GOLDEN = [   # (name, f, input, expected gradient of the sum), eps 1e-10
    ("testBraidedBuilds", braidedBuilds, np.full(4, 3.4), np.array([0.0, 4.0, 0.0, 0.0])),
    ("testBraidedBuilds1", braidedBuilds, T16, np.array([0.0] * 8 + [4.0] * 8).reshape(2, 2, 1, 2, 2)),
    ("testConcatBuild2", concatBuild2, np.array([0.651, 0.14, 0.3414]), np.full(3, 16.0)),
    ("testConcatBuild22", concatBuild2, T48, np.full(T48.shape, 16.0)),
    ("testConcatBuild0m", concatBuildm, V7, np.full(7, -1.0)),
    ("testConcatBuild1m", concatBuildm, T48, np.full(T48.shape, 962.0)),
]


# ---- detSquare (TestGatherSimplified.hs:1725-1790): the determinant as a sum over all permutations,
#      det a = ssum0 (kbuild1 @(ell!) (\i -> (-1) ** q[i] * rfold (*) 1 (rgather1 ell a (\i2 -> [i2, p[i, i2]]))))
#      with p = the permutation table (idx_to_perm: factorial-base digits pick the n-th free spot = lexicographic
#      order) and q = the inversion numbers; a data-dependent gather (table look-up inside the index function), a
#      product fold and a per-element host build in one program ----
def _perm_tables(n):
    import itertools
    perms = np.array(list(itertools.permutations(range(n))), dtype=np.int64)           # lexicographic = idx_to_perm
    inv = np.array([sum(1 for a in range(n) for b in range(a + 1, n) if p[a] > p[b]) for p in perms], dtype=np.int64)
    return perms, inv


def detSquare(T, a):
    from horde_ad_b200 import nn
    ell = T.shape(a)[1]
    perms, inv = _perm_tables(ell)
    p = T.sconcrete(perms) if not hasattr(T, "B") else T.B.sconcrete(perms)     # a plain (non-differentiated) table
    one = T.sscalar(1.0)

    def f(i):
        row = T.sgather([ell], a, lambda ix: [ix[0], T.ix_tbl(p, [i, ix[0]])], 2)
        prod = nn.sfold(T, lambda acc, e: T.binary("mul", acc, e), one, row)
        return T.binary("mul", T.sscalar(float((-1.0) ** int(inv[i]))), prod)
    return T.ssum0(T.sfromVector([f(i) for i in range(len(perms))]))


DET6_IN = 0.01 * np.array([40.1, 32.1, 40.1, 32.1, 40.1, 22.0,
                           9.71, 58.8943200001, 18.1, 29.1, 32.1, 40.1,
                           55.8943200001, 1, -2, 0.97, 58.200001, 1,
                           40.1, 29.0, 53.99432, 7.1, 58.8943200001, 18.1,
                           32.1, 40.1, 29.0, 53.99432, 2.971, 58.8943200001,
                           40.1, 32.0, 53.99432, 0.97, 25.8943200001, 5]).reshape(6, 6)
DET6_GRAD = np.array([-4.869053606042778e-3,5.097805950791719e-4,2.8429105389292052e-3,2.7764540017998364e-2,4.728231081593177e-3,-2.4786176657837146e-2,-2.3469141678115176e-3,1.1152144353350756e-2,-5.653171988563392e-3,-2.9237121329780706e-3,1.9330615364900276e-3,-9.47499409844526e-4,8.025631739575951e-3,1.532840697726606e-3,-7.047121923697899e-3,-4.305940089911722e-3,1.4834586948324607e-4,1.9922535230070225e-3,-8.022565662650867e-3,-1.2603575569596842e-2,8.144948602197404e-3,-9.343418236792108e-3,8.06305133890595e-3,1.7102844846324558e-2,4.888906896881243e-3,-4.884300437195866e-3,4.144990592410334e-4,-4.765756648788264e-3,-4.741976230085896e-3,1.3056977293001447e-2,1.0096879396193574e-2,1.1166653675142208e-2,-1.279846251614077e-4,-7.896243314270949e-3,-9.67019219688691e-3,-5.316279354872351e-3]).reshape(6, 6)


def fooNoGo(T, v):
    """fooNoGo (TestAdaptorSimplified.hs:1730-1745): rbuild1 with an index clamped by minH / maxH after a floor of
    the sum, ifH on a tensor element, rmap0N with a data-dependent choice, rslice, elementwise / and *."""
    from horde_ad_b200 import nn
    r = T.ssum0(v)
    rv = float(T.kfromS(r))

    def body(ix):
        j = max(min(ix + int(np.floor(rv)), 2), 0)
        inner = _bar(T, T.sscalar(3.14), _bar(T, T.sscalar(3.14), T.sindex(v, [j])))
        cond = float(T.kfromS(T.sindex(v, [ix * 2]))) <= 0 and 6 > abs(ix)
        return T.binary("add", inner, r if cond else T.binary("mul", T.sscalar(5.0), r))
    built = nn.sbuild1(T, 3, body)
    mapped = nn.smap0N(T, lambda x: r if float(T.kfromS(x)) > rv else x, v)
    ones = nn.sbuild1(T, 3, lambda _i: T.sscalar(1.0))
    return T.binary("mul", T.binary("divide", built, T.sslice(1, 3, mapped)), ones)


FOONOGO_IN = np.array([1.1, 2.2, 3.3, 4, 5])
FOONOGO_GRAD = np.array([5.037878787878788, -14.394255484765257, 43.23648655081373, -0.8403418295960368, 5.037878787878788])


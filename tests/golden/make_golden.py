"""Writes tests/golden/reference_goldens.json: the literal known-answer vectors
of horde-ad's own test-suite for the hot path, transcribed in tests/programs.py
(each entry cites the reference file:line it was copied from).  Run from the
repository root:  python tests/golden/make_golden.py

The reference is Haskell and cannot be executed in this environment (no GHC),
so these literals -- not outputs of a run -- are what pins the oracle.
"""
import json

import numpy as np
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import programs as P  # noqa: E402

out = {
    "source": "horde-ad v0.4.0.0 test-suite literals",
    "fixtures": {
        "t16b": {"cite": "test/tool/CrossTesting.hs:460-462", "shape": list(P.T16B.shape), "data": P.T16B.reshape(-1).tolist()},
        "t128": {"cite": "test/tool/CrossTesting.hs:463-473", "shape": list(P.T128.shape), "data": P.T128.reshape(-1).tolist()},
        "t48": {"cite": "test/tool/CrossTesting.hs:463-464", "shape": list(P.T48.shape), "data": P.T48.reshape(-1).tolist()},
        "vals72": {"cite": "test/simplified/TestGatherSimplified.hs:212", "shape": [7, 2], "data": P.VALS72.reshape(-1).tolist()},
    },
    "gather_scatter_gradients": [
        {"name": n, "cite": "test/simplified/TestGatherSimplified.hs:154-398,961-1232", "input": inp.reshape(-1).tolist(),
         "input_shape": list(inp.shape), "expected_gradient": [float(x) for x in exp], "eps": 1e-10}
        for n, _prog, inp, exp in P.GATHER_GOLDEN],
    "conv_gradients": [
        {"name": n, "cite": "test/simplified/TestConvSimplified.hs:196-228,287-297", "wrt": wrt,
         "K": K.reshape(-1).tolist(), "K_shape": list(K.shape), "A": A.reshape(-1).tolist(), "A_shape": list(A.shape),
         "expected_gradient": [float(x) for x in exp], "eps": eps}
        for n, K, A, wrt, exp, eps in P.CONV_GOLDEN],
    "bar_relu_gradient": {"cite": "test/simplified/TestGatherSimplified.hs:1283-1288", "input": "0.001 * t128",
                          "expected_gradient": P.BARRELU_GOLDEN.reshape(-1).tolist(), "eps": 1e-10},
    "scalars": [
        {"name": "README foo value", "cite": "README.md:104-105", "value": 4.242393641025528},
        {"name": "README foo gradient", "cite": "README.md:75-76,89-90", "value": [2.4396285219055063, -1.953374825727421, 0.9654825811012627]},
        {"name": "testListProd", "cite": "test/simplified/TestAdaptorSimplified.hs:1043-1053", "input": [1, 2, 3, 4], "value": [24, 12, 8, 6]},
        {"name": "testSin0Fold2", "cite": "test/simplified/TestRevFwdFold.hs:545-550", "value": 0.12389721944941383},
        {"name": "testCNNOPP3 folded value", "cite": "test/simplified/TestGatherSimplified.hs:1375", "shape": [2, 2, 2, 2], "value": P.CNNOPP3_GOLDEN},
        {"name": "testCNNOPP6 folded value", "cite": "test/simplified/TestGatherSimplified.hs:1590", "shape": [2, 2, 2, 2], "value": P.CNNOPP6_GOLDEN},
        {"name": "testCNNOPP7 folded value", "cite": "test/simplified/TestGatherSimplified.hs:1628", "shape": [2, 2, 2, 2], "value": P.CNNOPP7_GOLDEN},
        *[{"name": "vstackABC " + n, "cite": "test/simplified/TestAdaptorSimplified.hs:815-835,949", "value": e} for n, _mk, e in P.VSTACK_GOLDEN],
        {"name": "testGatherTranspose33", "cite": "test/simplified/TestGatherSimplified.hs:564-600", "value": P.GT33_GOLDEN},
        {"name": "testGatherTransposeBuild33", "cite": "test/simplified/TestGatherSimplified.hs:600-606", "value": P.GT33_BUILD_GOLDEN},
        {"name": "testGatherCond / testGatherCond2", "cite": "test/simplified/TestGatherSimplified.hs:716-770", "value": P.GATHER_COND_GOLDEN},
        {"name": "testGatherCond3 / testGatherCond4", "cite": "test/simplified/TestGatherSimplified.hs:795-835", "value": P.GATHER_COND34_GOLDEN},
        {"name": "testGatherCond5 / testGatherCond6", "cite": "test/simplified/TestGatherSimplified.hs:870-915", "value": P.GATHER_COND56_GOLDEN},
        *[{"name": "test" + n[0].upper() + n[1:], "cite": "test/simplified/TestGatherSimplified.hs:161-398,968-1232", "value": e}
          for n, _f, e in P.GATHER_BUILD_GOLDEN],
        {"name": "testSin0FoldNestedS1", "cite": "test/simplified/TestRevFwdFold.hs:2316-2326", "value": 2.0504979297616553e-43},
        {"name": "testSin0Fold3", "cite": "test/simplified/TestRevFwdFold.hs:558-563", "value": 0.4535961214255773},
        {"name": "testSin0Fold4", "cite": "test/simplified/TestRevFwdFold.hs:565-570", "value": -0.7053476446727861},
        {"name": "testSin0Fold6", "cite": "test/simplified/TestRevFwdFold.hs:583-590", "value": 6},
        {"name": "testSin0Fold7", "cite": "test/simplified/TestRevFwdFold.hs:592-598", "value": 250},
        *[{"name": n, "cite": "test/simplified/TestRevFwdFold.hs:934-1000", "input": x.reshape(-1).tolist(),
           "expected_gradient": e.reshape(-1).tolist(), "eps": 1e-10} for n, _f, x, e in P.SCAN_GOLDEN],
        *[{"name": n, "cite": "test/simplified/TestAdaptorSimplified.hs:1247-1251,1300-1345,1404-1462,1850-1895",
           "inputs": [x.reshape(-1).tolist() for x in ins], "expected_gradients": [e.reshape(-1).tolist() for e in exp], "eps": 1e-10}
          for n, _f, ins, exp in P.ADAPTOR_REV_GOLDEN],
        *[{"name": n, "cite": "test/simplified/TestAdaptorSimplified.hs:1843-1848,1872-1877", "input": np.asarray(x).reshape(-1).tolist(),
           "dt": np.asarray(dt).reshape(-1).tolist(), "value": np.asarray(e).reshape(-1).tolist()}
          for n, _f, x, dt, e in P.ADAPTOR_VJP_GOLDEN],
        *[{"name": n, "cite": "test/simplified/TestAdaptorSimplified.hs:1896-1925", "input": x.reshape(-1).tolist(),
           "tangent": dx.reshape(-1).tolist(), "value": e.reshape(-1).tolist()} for n, _f, x, dx, e in P.ADAPTOR_FWD_GOLDEN],
    ],
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote reference_goldens.json")

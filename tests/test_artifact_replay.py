"""The reference's PRINTED gradient artifacts (testCNNOPP2S, test/simplified/TestMnistPP.hs:728-735: the complete
MnistCnnShaped2 forward and gradient programs, plain and simplified) as golden PROGRAMS: parsed from the committed
fixture tests/golden/cnn_artifacts.json and replayed call by call against a carrier (tests/artifact_replay.py).

CPU: on the oracle all four artifacts must agree with this repo's own AD (ad.py) of this repo's own restatement of
the model (mnist_cnn.py) -- which pins the oracle's op semantics, the transposition rules of ad.py and the model
against the op sequence horde-ad itself emits.  GPU (`-m gpu`): the same artifacts through the C ABI, eagerly and
recorded + rewritten by the contraction pass."""
import json
import os

import numpy as np
import pytest

from horde_ad_b200 import ad, mnist_cnn
from tests.artifact_replay import run_artifact

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cnn_artifacts.json")))
NAMES = ["ker1", "bias1", "ker2", "bias2", "weightsDense", "biasesDense", "weightsReadout", "biasesReadout"]


def _inputs(seed=44):
    rng = np.random.Generator(np.random.Philox(seed))
    P = [0.8 * (rng.random(G["shapes"][n]) - 0.5) for n in NAMES]       # randomValue 0.4 (Adaptor.hs:166-179)
    dret = rng.uniform(-1, 1, (10, 7))
    return P, dret


def _pairs(P):
    return (((P[0], P[1]), (P[2], P[3])), ((P[4], P[5]), (P[6], P[7])))


def _flat(t):
    return [t[0][0][0], t[0][0][1], t[0][1][0], t[0][1][1], t[1][0][0], t[1][0][1], t[1][1][0], t[1][1][1]]


GLYPH = np.full(G["shapes"]["glyph"], 7.0)          # blackGlyph: treplicate ... (sscalar 7)


def _reference_values(oracle, P, dret):
    fwd = np.asarray(mnist_cnn.convMnistTwoS(oracle, GLYPH, P, fused=False))
    _, grads = ad.grad(oracle, lambda T, *ps: mnist_cnn.convMnistTwoS(T, GLYPH, list(ps), fused=False), P, dt=dret)
    return fwd, [np.asarray(g) for g in grads]


@pytest.mark.parametrize("which", ["primal_simplified", "primal"])
def test_printed_primal_artifacts_equal_the_model_on_the_oracle(oracle, which):
    P, dret = _inputs()
    out, ev = run_artifact(oracle, G["artifacts"][which], {"u1": _pairs(P)})
    fwd, _ = _reference_values(oracle, P, dret)
    assert ev.n_calls > 100
    np.testing.assert_allclose(np.asarray(out), fwd, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("which", ["rev_simplified", "rev"])
def test_printed_gradient_artifacts_equal_our_ad_on_the_oracle(oracle, which):
    P, dret = _inputs()
    out, ev = run_artifact(oracle, G["artifacts"][which], {"u1": _pairs(P), "dret": dret})
    _, grads = _reference_values(oracle, P, dret)
    assert ev.n_calls > 180
    for name, a, b in zip(NAMES, _flat(out), grads):
        a = np.asarray(a)
        assert a.shape == b.shape, (name, a.shape, b.shape)
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12 * np.max(np.abs(b)), err_msg=name)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["rev_simplified", "rev", "primal_simplified"])
def test_printed_artifacts_through_the_c_abi(dev, oracle, which):
    """Every call of the printed program goes through the C ABI (gathers with table / kargMax / ifH index programs,
    sdot1In over broadcast views, scatters); data movement is exact, so the result must match the oracle's replay of
    the same text to rounding."""
    P, dret = _inputs()
    dP = [dev.sconcrete(p) for p in P]
    ins = {"u1": _pairs(dP), "dret": dev.sconcrete(dret)}
    oins = {"u1": _pairs(P), "dret": dret}
    if which.startswith("primal"):
        ins.pop("dret"), oins.pop("dret")
    out, _ = run_artifact(dev, G["artifacts"][which], ins)
    ref, _ = run_artifact(oracle, G["artifacts"][which], oins)
    got = _flat(out) if isinstance(out, tuple) else [out]
    want = _flat(ref) if isinstance(ref, tuple) else [ref]
    for a, b in zip(got, want):
        b = np.asarray(b)
        np.testing.assert_allclose(dev.to_numpy(a), b, rtol=1e-12, atol=1e-12 * np.max(np.abs(b)))


@pytest.mark.gpu
def test_printed_gradient_artifact_recorded_and_rewritten(dev, oracle):
    """The simplified gradient artifact recorded as a lazy program, run through the contraction pass and replayed."""
    P, dret = _inputs()
    dP = [dev.sconcrete(p) for p in P]
    ddret = dev.sconcrete(dret)
    prog = dev.record(lambda: _flat(run_artifact(dev, G["artifacts"]["rev_simplified"], {"u1": _pairs(dP), "dret": ddret})[0]))
    log = prog.dump().split("--")[1]
    assert "relu_pool_fwd" in log and log.count("-> relu") >= 3, log
    prog.run()
    _, grads = _reference_values(oracle, P, dret)
    for a, b in zip(prog.outs, grads):
        np.testing.assert_allclose(dev.to_numpy(a), b, rtol=1e-12, atol=1e-12 * np.max(np.abs(b)))

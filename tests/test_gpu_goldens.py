"""The reference's known-answer tests that round 1 only ran on the oracle, replayed ON THE DEVICE through the C ABI
(`-m gpu`): the test bodies are the backend-generic ones of tests/test_oracle_golden.py, handed the device carrier
instead of the CPU oracle, so the literals of the reference's own test-suite pin the CUDA path directly."""
import numpy as np
import pytest

from tests import programs as P
from tests import test_oracle_golden as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.int64, np.int32, np.int16, np.int8])
def test_overleaf_gather_with_wraparound_on_the_device(dev, dtype):
    G.test_overleaf_gather_with_wraparound(dev, dtype)


def test_high_rank_logistic_and_bar_relu_known_answers_on_the_device(dev):
    G.test_high_rank_logistic_and_bar_relu_known_answers(dev)


def test_map_accum_empty_and_single_step_known_answers_on_the_device(dev):
    G.test_map_accum_empty_and_single_step_known_answers(dev)


@pytest.mark.parametrize("name,f,x,expected", P.BUILD_GOLDEN, ids=[g[0] for g in P.BUILD_GOLDEN])
def test_high_rank_build_known_answers_on_the_device(dev, name, f, x, expected):
    G.test_high_rank_build_known_answers(dev, name, f, x, expected)


def test_det_square_known_answer_on_the_device(dev):
    G.test_det_square_known_answer(dev)


def test_fold_goldens_on_the_device(dev):
    G.test_fold_goldens(dev)
    G.test_fold_goldens_with_tensor_steps(dev)


def test_foo_no_go_known_answer_on_the_device(dev):
    G.test_foo_no_go_known_answer(dev)


def test_foo_float_rank5_known_answer_on_the_device(dev):
    G.test_foo_float_rank5_known_answer(dev)


def test_readme_and_list_prod_known_answers_on_the_device(dev):
    G.test_readme_foo_known_answers(dev)
    G.test_list_prod_gradient(dev)

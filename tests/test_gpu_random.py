"""The randomised configurations of tests/test_model_random.py on the shipped CUDA library."""
import pytest

from oracle import pyoracle as po
from test_model_random import draw_lexfree, draw_lexicon, draw_widened, run_random

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from flt_backend import FltBackend

    return FltBackend("cuda")


@pytest.fixture(scope="module")
def A():
    return po.Oracle("ora")


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_lexfree_random(A, G, seed):
    run_random(A, G, draw_lexfree, seed, 40, 1e-4)


@pytest.mark.parametrize("seed", [10, 11, 12, 13])
def test_lexicon_random(A, G, seed):
    run_random(A, G, draw_lexicon, seed, 30, 1e-4)


@pytest.mark.parametrize("seed", [20, 21, 22, 23])
def test_widened_random(A, G, seed):
    run_random(A, G, draw_widened, seed, 40, 1e-4)

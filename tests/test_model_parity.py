"""Kernel LOGIC vs the oracle, without a GPU: the product's kernel sources compiled by g++ as a
one-thread-per-CTA sequential program (tests/model, -DFLT_HOST_MODEL) must reproduce the oracle's
n-best lists on every parity case. This catches algorithmic errors (candidate pruning bounds,
merge keys, arithmetic order) in the build container; races and CUDA-specific behaviour are covered
by the same cases run on the B200 in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import parity_cases
from cases import Built, assert_same_nbest, has_ties
from flt_backend import FltBackend
from oracle import pyoracle as po


@pytest.fixture(scope="module")
def M():
    return FltBackend("model")


@pytest.fixture(scope="module")
def A():
    return po.Oracle("ora")


def run_case(A, M, spec, em):
    ba, bm = Built(A, spec), Built(M, spec)
    K = spec["opt"].beamSize
    got = bm.O.decode_batch(bm.dec, em, K)
    checked = 0
    for b, e in enumerate(em):
        ra = ba.decode(e)
        if has_ties(ra):
            continue
        assert_same_nbest(ra, got[b], 1e-9, what=f"utt {b}")
        checked += 1
    ba.close(), bm.close()
    assert checked, "all utterances had score ties: vacuous"


@pytest.mark.parametrize("name,spec,em", parity_cases.lexfree_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_lexfree(A, M, name, spec, em):
    run_case(A, M, spec, em)


@pytest.mark.parametrize("name,spec,em", parity_cases.lexicon_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_lexicon(A, M, name, spec, em):
    run_case(A, M, spec, em)

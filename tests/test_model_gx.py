"""The experimental two-pass frame step (text_b200/csrc/beam_gx.h, opt-in with FLT_GX=1) on the
GPU-less logic harness: same parity cases, same oracle, bit-equal n-best. FLT_DBG=2-free: the step has no
guessing left; FLT_TEST_CAP=1 starts from a tiny candidate capacity so that the zoom / capacity-retry paths
run too."""
import os

import pytest

import parity_cases
from cases import Built, assert_same_nbest, has_ties
from oracle import pyoracle as po


@pytest.fixture(scope="module", params=["plain", "tiny_capacity"])
def M(request):
    from flt_backend import FltBackend

    old = {k: os.environ.get(k) for k in ("FLT_GX", "FLT_TEST_CAP")}
    os.environ["FLT_GX"] = "1"
    if request.param == "tiny_capacity":
        os.environ["FLT_TEST_CAP"] = "1"
    yield FltBackend("model")
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.fixture(scope="module")
def A():
    return po.Oracle("ora")


CASES = [c for c in parity_cases.lexfree_cases() + parity_cases.lexicon_cases()
         if c[0] in ("cfg1_shape", "bst_small", "thr_silneg", "sil_positive_bst", "asg", "cfg2_scaled_bstN",
                     "chunk_path_bst300", "beam_gt_cands", "zero_ctc", "zero_ctc_bst_thr", "zero_ctc_silpos",
                     "zero_ctc_words1", "cfg3_scaled_bstN", "cfg3_mid", "arpa3_ctc", "arpa3_ctc_bst", "cfg4_scaled")]


@pytest.mark.parametrize("name,spec,em", CASES, ids=[c[0] for c in CASES])
def test_gx_step(A, M, name, spec, em):
    ba, bm = Built(A, spec), Built(M, spec)
    K = spec["opt"].beamSize
    got = bm.O.decode_batch(bm.dec, em, K)
    checked = 0
    for b, e in enumerate(em):
        ra = ba.decode(e)
        if has_ties(ra):
            continue
        assert_same_nbest(ra, got[b], 1e-9, what=f"{name} utt {b}")
        checked += 1
    ba.close(), bm.close()
    assert checked or name.startswith("zero_"), "all utterances had score ties: vacuous"

"""Validate the CPU restatement (ora_*) against the unmodified reference compiled in place
(ref_*, oracle/_ref/libflref.so) on seeded synthetic inputs: n-best token/word strings bit-equal,
scores equal (same arithmetic order, so atol is 1e-9, far below the 1e-4 contract)."""
import os

import numpy as np
import pytest

from cases import Built, spec_lexfree, spec_lexicon, assert_same_nbest, has_ties
from oracle import pyoracle as po
from text_b200 import synth

pytestmark = pytest.mark.skipif(not po.available("ref"), reason="compiled reference not present")


@pytest.fixture(scope="module")
def R():
    return po.Oracle("ref")


@pytest.fixture(scope="module")
def A():
    return po.Oracle("ora")


def _both(R, A, spec, em, max_hyp=None):
    br, ba = Built(R, spec), Built(A, spec)
    out = []
    for e in em:
        rr, ra = br.decode(e, max_hyp), ba.decode(e, max_hyp)
        if has_ties(rr):
            continue  # implementation-defined order
        assert_same_nbest(rr, ra, 1e-9, what=str(spec["opt"].beamSize))
        out.append(rr)
    br.close(), ba.close()
    assert out, "every utterance had score ties: test is vacuous"
    return out


LEXFREE = [
    # N, T, beam, bst, thr, criterion, log_add, sil_score
    (29, 200, 10, 29, 1e9, po.CTC, False, 0.0),   # BASELINE config 1 shape
    (29, 80, 10, 5, 1e9, po.CTC, False, 0.0),
    (29, 80, 10, 29, 8.0, po.CTC, False, -0.5),
    (64, 60, 25, 64, 1e9, po.CTC, True, 0.0),
    (64, 60, 25, 7, 12.0, po.CTC, True, 0.3),
    (40, 60, 12, 40, 1e9, po.ASG, False, 0.0),
    (40, 60, 12, 6, 20.0, po.ASG, True, -1.0),
    (500, 40, 50, 500, 1e9, po.CTC, False, 0.0),  # config-2 shape, scaled down, bst = N
    (500, 40, 50, 50, 1e9, po.CTC, False, 0.0),   # bst = beam
    (3, 30, 4, 3, 1e9, po.CTC, False, 0.0),       # tiny vocabulary
    (10, 1, 5, 10, 1e9, po.CTC, False, 0.0),      # single frame
]


@pytest.mark.parametrize("N,T,beam,bst,thr,crit,log_add,sil_score", LEXFREE)
def test_lexfree_zero_lm(R, A, N, T, beam, bst, thr, crit, log_add, sil_score):
    em = synth.emissions(3, T, N, seed=100 + N + T + beam, sigma=1.0)
    tr = None
    if crit == po.ASG:
        tr = np.random.default_rng(5).random(N * N, dtype=np.float32)
    spec = spec_lexfree(N, beam, bst, thr, sil=0, blank=(N - 1 if crit == po.CTC else -1),
                        criterion=crit, log_add=log_add, sil_score=sil_score, transitions=tr)
    _both(R, A, spec, em)


LEXICON = [
    # N, T, W, beam, bst, thr, criterion, log_add, sil_score, word_score, unk_score
    (30, 60, 200, 20, 30, 1e9, po.CTC, False, 0.0, 0.0, float("-inf")),
    (30, 60, 200, 20, 8, 15.0, po.CTC, False, -0.2, 1.0, float("-inf")),
    (30, 60, 200, 20, 30, 1e9, po.CTC, True, 0.0, 0.5, float("-inf")),
    (30, 60, 200, 20, 30, 1e9, po.CTC, False, 0.0, 0.0, -2.0),       # unk branch on
    (30, 50, 200, 20, 30, 1e9, po.ASG, False, -0.3, 0.7, float("-inf")),
    (200, 40, 2000, 50, 200, 1e9, po.CTC, False, 0.0, 0.0, float("-inf")),  # config-3 shape, scaled
    (200, 40, 2000, 50, 50, 25.0, po.CTC, False, 0.0, 0.0, float("-inf")),
]


@pytest.mark.parametrize("N,T,W,beam,bst,thr,crit,log_add,sil_score,word_score,unk_score", LEXICON)
def test_lexicon_zero_lm(R, A, N, T, W, beam, bst, thr, crit, log_add, sil_score, word_score,
                         unk_score):
    blank = N - 1 if crit == po.CTC else -1
    sp = synth.lexicon(W, N, 2, 4, seed=7, exclude=(0, N - 1))
    em = synth.emissions(3, T, N, seed=300 + N + beam, sigma=2.0)
    tr = np.random.default_rng(5).random(N * N, dtype=np.float32) if crit == po.ASG else None
    spec = spec_lexicon(N, beam, bst, sp, thr, sil=0, blank=blank, criterion=crit, log_add=log_add,
                        sil_score=sil_score, word_score=word_score, unk_score=unk_score,
                        transitions=tr)
    _both(R, A, spec, em)


@pytest.fixture(scope="module")
def small_arpa():
    W = 300
    path = os.path.join(synth.cache_dir(), "t_small4.arpa")
    synth.write_arpa(path, W, order=4, counts=[0, 3000, 3000, 2000], seed=3)
    return W, path, synth.word_names(W) + ["<unk>"]


@pytest.mark.parametrize("smear", [po.SMEAR_MAX, po.SMEAR_NONE])
@pytest.mark.parametrize("log_add", [False, True])
def test_lexicon_arpa_lm(R, A, small_arpa, smear, log_add):
    W, path, words = small_arpa
    N, T = 40, 80
    sp = synth.lexicon(W, N, 1, 3, seed=9, exclude=(0, N - 1))
    em = synth.emissions(3, T, N, seed=77, sigma=2.0)
    spec = spec_lexicon(N, 40, N, sp, 30.0, lm_weight=1.5, word_score=0.5, lm=("arpa", path, words),
                        smear=smear, log_add=log_add, unk=W)
    _both(R, A, spec, em)


def test_lexfree_token_arpa_lm(R, A, small_arpa):
    """LexiconFreeDecoder with a token-level n-gram LM (tokens are the LM's words)."""
    W, path, words = small_arpa
    N, T = 50, 60
    em = synth.emissions(2, T, N, seed=78, sigma=2.0)
    spec = spec_lexfree(N, 15, N, 1e9, lm_weight=0.8, lm=("arpa", path, words[:N]))
    _both(R, A, spec, em)


def test_lexicon_token_lm(R, A, small_arpa):
    W, path, words = small_arpa
    N, T = 30, 50
    sp = synth.lexicon(150, N, 1, 4, seed=10, exclude=(0, N - 1))
    em = synth.emissions(2, T, N, seed=79, sigma=2.0)
    # word_score != 0: with a token LM the node and word candidates of one token otherwise tie
    spec = spec_lexicon(N, 20, N, sp, 1e9, lm_weight=0.7, word_score=0.37,
                        lm=("arpa", path, words[:N]), is_lm_token=True)
    _both(R, A, spec, em)


def test_invalid_lm_index_fails_like_kenlm(R, A, small_arpa):
    """lm/KenLM.cpp:66-69 throws on an out-of-range user index; both sides must fail."""
    W, path, words = small_arpa
    for O in (R, A):
        lm = O.lm_arpa(path, words[:5])
        with pytest.raises(RuntimeError):
            O.lm_score_seq(lm, [7])
        O.lm_destroy(lm)


def test_trie_bounds_and_label_cap(R, A):
    for O in (R, A):
        t = O.trie_create(5, 0)
        with pytest.raises(IndexError):
            O.trie_insert(t, [1, 7], 0, 0.0)  # Trie.cpp:31-34
        for lab in range(8):                  # 7th+ label dropped, Trie.cpp:40-46
            O.trie_insert(t, [1, 2], lab, -float(lab))
        O.trie_smear(t, po.SMEAR_LOGADD)
        got = O.trie_search(t, [1, 2])
        assert got["labels"].tolist() == [0, 1, 2, 3, 4, 5]
        assert O.trie_search(t, [3]) is None
        O.trie_destroy(t)
    # smear values agree (single child chain => order-independent)
    vals = []
    for O in (R, A):
        t = O.trie_create(5, 0)
        O.trie_insert(t, [1, 2], 0, -1.5)
        O.trie_insert(t, [1, 2], 1, -0.5)
        O.trie_insert(t, [1, 3], 2, -2.5)
        O.trie_smear(t, po.SMEAR_LOGADD)
        vals.append((O.trie_search(t, [1])["maxScore"], O.trie_search(t, [1, 2])["maxScore"]))
        O.trie_destroy(t)
    assert vals[0] == vals[1]


@pytest.mark.parametrize("lexicon", [False, True])
def test_streaming_prune_api(R, A, lexicon):
    """decodeBegin / chunked decodeStep / getBestHypothesis(lookBack) / prune / decodeEnd
    (Decoder.h:18-35, Utils.h:268-342)."""
    N, T, chunk = 30, 90, 15
    em = synth.emissions(1, T, N, seed=91, sigma=2.0)[0]
    if lexicon:
        sp = synth.lexicon(200, N, 2, 4, seed=7, exclude=(0, N - 1))
        spec = spec_lexicon(N, 20, N, sp, 1e9, word_score=0.3)
    else:
        spec = spec_lexfree(N, 12, N, 1e9)
    br, ba = Built(R, spec), Built(A, spec)
    for O, b in ((R, br), (A, ba)):
        O.decode_begin(b.dec)
    for c in range(0, T, chunk):
        outs = []
        for O, b in ((R, br), (A, ba)):
            O.decode_step(b.dec, em[c:c + chunk])
            best = O.best(b.dec, 5, T + 2)
            nh, nf = O.n_hypothesis(b.dec), O.n_frames_in_buffer(b.dec)
            O.prune(b.dec, 5)
            outs.append((best, nh, nf, O.n_frames_in_buffer(b.dec)))
        (b0, *r0), (b1, *r1) = outs
        assert r0 == r1
        np.testing.assert_array_equal(b0["tokens"], b1["tokens"])
        np.testing.assert_array_equal(b0["words"], b1["words"])
        np.testing.assert_allclose(b0["scores"], b1["scores"], atol=1e-9)
    fin = []
    for O, b in ((R, br), (A, ba)):
        O.decode_end(b.dec)
        fin.append(O.all_final(b.dec, 64, T + 2))
    assert fin[0]["n"] == fin[1]["n"] and fin[0]["n"] > 0
    np.testing.assert_array_equal(fin[0]["lens"], fin[1]["lens"])
    np.testing.assert_array_equal(fin[0]["tokens"], fin[1]["tokens"])
    np.testing.assert_allclose(fin[0]["scores"], fin[1]["scores"], atol=1e-9)
    br.close(), ba.close()

"""The reference-facing Python surface (pybind11 module `flashlight_lib_text_decoder` over the C++
mirror, text_b200/csrc/host/flashlight_text.h), without a GPU: names, kwargs, pickling, Trie / LM
host semantics against the oracle, error types, and that a decoder cannot be built without CUDA.
Mirrors what bindings/python/test/test_decoder.py:385-470 (pickling) and
flashlight/lib/text/test/decoder/DecoderTest.cpp:107-155 (LM / Trie known answers) check."""
import math
import os
import pickle
import sys

import numpy as np
import pytest

from oracle import pyoracle as po
from text_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text_b200", "compat"))


@pytest.fixture(scope="module")
def D():
    import flashlight.lib.text.decoder as d

    return d


def test_names_match_reference_package(D):
    for name in ("CriterionType", "DecodeResult", "LexiconDecoder", "LexiconDecoderOptions",
                 "LexiconFreeDecoder", "LexiconFreeDecoderOptions", "LM", "LMState", "SmearingMode", "Trie",
                 "TrieNode", "ZeroLM", "KenLM"):
        assert hasattr(D, name), name
    from flashlight.lib.text.decoder.kenlm import KenLM  # noqa: F401
    from flashlight.lib.text.dictionary import Dictionary  # noqa: F401
    assert {m for m in ("decode_begin", "decode_step", "decode_end", "decode", "prune", "get_best_hypothesis",
                        "get_all_final_hypothesis", "decode_batch")} <= set(dir(D.LexiconDecoder))
    assert [e for e in ("ASG", "CTC", "S2S")] == [n for n in D.CriterionType.__members__]
    assert [e for e in ("NONE", "MAX", "LOGADD")] == [n for n in D.SmearingMode.__members__]


def test_options_kwargs_and_pickle(D):
    o = D.LexiconDecoderOptions(beam_size=2500, beam_size_token=25000, beam_threshold=100.0, lm_weight=2.0,
                                word_score=2.0, unk_score=-math.inf, sil_score=-1, log_add=False,
                                criterion_type=D.CriterionType.ASG)
    o2 = pickle.loads(pickle.dumps(o))
    for f in ("beam_size", "beam_size_token", "beam_threshold", "lm_weight", "word_score", "unk_score",
              "sil_score", "log_add", "criterion_type"):
        assert getattr(o, f) == getattr(o2, f), f
    f = D.LexiconFreeDecoderOptions(beam_size=10, beam_size_token=29, beam_threshold=1e9, lm_weight=0.0,
                                    sil_score=0.5, log_add=False, criterion_type=D.CriterionType.CTC)
    f2 = pickle.loads(pickle.dumps(f))
    assert (f2.beam_size, f2.sil_score, f2.criterion_type) == (10, 0.5, D.CriterionType.CTC)
    r = D.DecodeResult(5)
    assert r.words == [-1] * 5 and r.tokens == [-1] * 5


def test_trie_matches_oracle(D):
    A = po.Oracle("ora")
    N = 30
    sp = synth.lexicon(300, N, 1, 4, seed=3, exclude=(0, N - 1))
    rng = np.random.default_rng(0)
    scores = rng.uniform(-5, 0, size=len(sp)).astype(np.float32)
    t, ta = D.Trie(N, 0), A.trie_create(N, 0)
    for w, (s, sc) in enumerate(zip(sp, scores)):
        node = t.insert([int(x) for x in s], w, float(sc))
        A.trie_insert(ta, s, w, float(sc))
        assert w in node.labels
    t.smear(D.SmearingMode.MAX)
    A.trie_smear(ta, po.SMEAR_MAX)
    for s in sp[:50]:
        got, want = t.search([int(x) for x in s]), A.trie_search(ta, s)
        assert got is not None and want is not None
        assert got.max_score == pytest.approx(want["maxScore"], abs=0)
        assert list(got.labels) == list(want["labels"])
    assert t.search([N - 1, N - 1, N - 1, N - 1, 1]) is None
    with pytest.raises(IndexError):  # std::out_of_range, Trie.cpp:31-34
        t.insert([N + 3], 0, 0.0)
    A.trie_destroy(ta)


def test_zero_lm_state_identity(D):
    lm = D.ZeroLM()
    s0 = lm.start(False)
    a, sc = lm.score(s0, 7)
    b, _ = lm.score(s0, 7)
    c, _ = lm.score(s0, 8)
    assert sc == 0.0 and a.compare(b) == 0 and a.compare(c) != 0  # lm/ZeroLM.cpp:18-22, lm/LM.h:37-49
    f, fs = lm.finish(a)
    assert fs == 0.0 and f.compare(a) == 0  # lm/ZeroLM.cpp:24-26


def test_kenlm_arpa_scores_match_oracle(D, tmp_path):
    from flashlight.lib.text.dictionary import Dictionary

    W = 40
    path = str(tmp_path / "t.arpa")
    synth.write_arpa(path, W, order=3, counts=[0, 200, 150], seed=5)
    words = synth.word_names(W) + ["<unk>", "notinlm"]
    lm = D.KenLM(path, Dictionary(words))
    A = po.Oracle("ora")
    la = A.lm_arpa(path, words)
    seq = [3, 17, 17, W + 1, 5, 0]
    want = A.lm_score_seq(la, seq, True)
    st, got = lm.start(False), []
    for w in seq:
        st, s = lm.score(st, w)
        got.append(s)
    _, fs = lm.finish(st)
    np.testing.assert_array_equal(np.array(got + [fs], np.float32), want)
    with pytest.raises(RuntimeError):  # lm/KenLM.cpp:66-69
        lm.score(lm.start(False), len(words) + 5)
    A.lm_destroy(la)


def test_decoder_needs_cuda_and_device_lm(D):
    import torch

    opt = D.LexiconFreeDecoderOptions(10, 29, 1e9, 0.0, 0.0, False, D.CriterionType.CTC)

    class MyLM(D.LM):
        def start(self, start_with_nothing):
            return D.LMState()

        def score(self, state, idx):
            return state.child(idx), 0.0

        def finish(self, state):
            return state, 0.0

    with pytest.raises(ValueError):  # std::invalid_argument: no device model behind a Python LM
        D.LexiconFreeDecoder(opt, MyLM(), 0, 28, [])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            D.LexiconFreeDecoder(opt, D.ZeroLM(), 0, 28, [])


def test_trie_children_can_be_walked(D):
    """TrieNode::children (decoder/Trie.h:39-54): the mirror keeps the reference's node tree, so code that
    walks it from get_root() sees every inserted spelling, with labels, scores and smeared scores."""
    trie = D.Trie(10, 0)
    words = {(1, 2): (7, -1.0), (1, 2, 3): (8, -2.0), (1, 4): (9, -0.5), (5,): (3, -3.0)}
    for sp, (label, score) in words.items():
        trie.insert(list(sp), label, score)
    trie.smear(D.SmearingMode.MAX)
    root = trie.get_root()
    assert sorted(root.children.keys()) == [1, 5]
    found = {}

    def walk(node, path):
        for lab, sc in zip(node.labels, node.scores):
            found[tuple(path)] = (lab, sc)
        for tok, child in node.children.items():
            assert child.idx == tok
            walk(child, path + [tok])

    walk(root, [])
    assert found == {k: (v[0], v[1]) for k, v in words.items()}
    n1 = root.children[1]
    assert abs(n1.max_score - (-0.5)) < 1e-6 and abs(root.max_score - (-0.5)) < 1e-6
    assert abs(n1.children[2].max_score - (-1.0)) < 1e-6
    assert trie.search([1, 2, 3]).labels == [8] and trie.search([2]) is None
    with pytest.raises(IndexError):
        trie.insert([1, 99], 1, 0.0)
    assert sorted(root.children[1].children.keys()) == [2, 4]  # nothing was added by the failed insert's tail

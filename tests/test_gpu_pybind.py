"""The reference's Python call sequence (bindings/python/test/test_decoder.py:92-252: build a
Dictionary-backed KenLM, score words into a Trie, smear, construct the decoder with keyword options,
`decode(emissions.ctypes.data, T, N)`) run against the CUDA path through the pybind11 module, and
compared with the CPU oracle on the same inputs. Also the C++ mirror through a compiled test
program (tests/cpp/mirror_test.cpp)."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import Built, assert_same_nbest, has_ties, spec_lexfree, spec_lexicon
from oracle import pyoracle as po
from text_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text_b200", "compat"))


def _as_res(results, T):
    n = len(results)
    return dict(n=n, scores=np.array([[r.score, r.emittingModelScore, r.lmScore] for r in results], np.float64).reshape(n, 3),
                tokens=np.array([r.tokens for r in results], np.int32).reshape(n, T + 2),
                words=np.array([r.words for r in results], np.int32).reshape(n, T + 2))


def test_lexicon_kenlm_like_reference_python_test(tmp_path):
    from flashlight.lib.text.decoder import (CriterionType, LexiconDecoder, LexiconDecoderOptions, SmearingMode,
                                             Trie)
    from flashlight.lib.text.decoder.kenlm import KenLM
    from flashlight.lib.text.dictionary import Dictionary

    N, T, W = 40, 70, 300
    path = str(tmp_path / "lm.arpa")
    synth.write_arpa(path, W, order=3, counts=[0, 3000, 2000], seed=9)
    words = synth.word_names(W) + ["<unk>"]
    spell = synth.lexicon(W, N, 1, 3, seed=7, exclude=(0, N - 1))
    em = synth.emissions(3, T, N, seed=42, sigma=2.0)
    # --- the reference's setup sequence
    word_dict = Dictionary(words)
    lm = KenLM(path, word_dict)
    trie = Trie(N, 0)
    start = lm.start(False)
    for w, sp in enumerate(spell):
        _, score = lm.score(start, w)
        trie.insert([int(x) for x in sp], w, score)
    trie.smear(SmearingMode.MAX)
    opts = LexiconDecoderOptions(beam_size=40, beam_size_token=N, beam_threshold=30.0, lm_weight=1.5, word_score=0.5,
                                 unk_score=-math.inf, sil_score=-0.1, log_add=False,
                                 criterion_type=CriterionType.CTC)
    dec = LexiconDecoder(opts, trie, lm, 0, N - 1, W, [], False)
    # --- oracle on the same inputs
    A = po.Oracle("ora")
    spec = spec_lexicon(N, 40, N, spell, 30.0, sil=0, blank=N - 1, lm_weight=1.5, word_score=0.5, sil_score=-0.1,
                        lm=("arpa", path, words), unk=W)
    ba = Built(A, spec)
    checked = 0
    for b in range(len(em)):
        e = np.ascontiguousarray(em[b])
        results = dec.decode(e.ctypes.data, T, N)
        ra = ba.decode(e)
        if has_ties(ra):
            continue
        assert_same_nbest(ra, _as_res(results, T), 1e-4, what=f"utt {b}")
        checked += 1
    assert checked
    # batch entry, streaming calls (buffered), best hypothesis
    allb = dec.decode_batch(np.ascontiguousarray(em).ctypes.data, len(em), T, N, nbest=5)
    assert [len(x) for x in allb] == [5] * len(em)
    e = np.ascontiguousarray(em[0])
    dec.decode_begin()
    dec.decode_step(e[:30].ctypes.data, 30, N)
    dec.decode_step(np.ascontiguousarray(e[30:]).ctypes.data, T - 30, N)
    dec.decode_end()
    res = dec.get_all_final_hypothesis()
    assert res[0].tokens == allb[0][0].tokens and res[0].score == allb[0][0].score
    assert dec.get_best_hypothesis().tokens == res[0].tokens
    ba.close()


def test_lexfree_zero_lm_device_pointer():
    import torch
    from flashlight.lib.text.decoder import CriterionType, LexiconFreeDecoder, LexiconFreeDecoderOptions, ZeroLM

    N, T = 64, 50
    em = synth.emissions(2, T, N, seed=5)
    opts = LexiconFreeDecoderOptions(beam_size=16, beam_size_token=N, beam_threshold=1e9, lm_weight=0.0,
                                     sil_score=0.0, log_add=False, criterion_type=CriterionType.CTC)
    dec = LexiconFreeDecoder(opts, ZeroLM(), 0, N - 1, [])
    A = po.Oracle("ora")
    ba = Built(A, spec_lexfree(N, 16, N, 1e9))
    d_em = torch.from_numpy(em).cuda()
    for b in range(2):
        results = dec.decode(d_em[b].data_ptr(), T, N)  # device address, like an acoustic model's output
        assert_same_nbest(ba.decode(em[b]), _as_res(results, T), 1e-4, what=f"utt {b}")
    ba.close()


def test_log_add_and_token_lm_through_python_names(tmp_path):
    """log_add=True (lexicon-free) and is_token_lm=True (lexicon + KenLM over tokens) with the
    reference's constructor signatures; both run as a full expansion on the device."""
    from cases import assert_close_nbest
    from flashlight.lib.text.decoder import (CriterionType, LexiconDecoder, LexiconDecoderOptions,
                                             LexiconFreeDecoder, LexiconFreeDecoderOptions, SmearingMode, Trie,
                                             ZeroLM)
    from flashlight.lib.text.decoder.kenlm import KenLM
    from flashlight.lib.text.dictionary import Dictionary

    A = po.Oracle("ora")
    N, T = 40, 50
    em = synth.emissions(3, T, N, seed=17, sigma=2.0)
    # lexicon-free, logAdd
    opts = LexiconFreeDecoderOptions(beam_size=12, beam_size_token=10, beam_threshold=20.0, lm_weight=0.0,
                                     sil_score=-0.2, log_add=True, criterion_type=CriterionType.CTC)
    dec = LexiconFreeDecoder(opts, ZeroLM(), 0, N - 1, [])
    ba = Built(A, spec_lexfree(N, 12, 10, 20.0, sil_score=-0.2, log_add=True))
    checked = 0
    for b in range(len(em)):
        e = np.ascontiguousarray(em[b])
        ra = ba.decode(e)
        if has_ties(ra) or A.tie_events(ba.dec):
            continue
        assert_close_nbest(ra, _as_res(dec.decode(e.ctypes.data, T, N), T), 1e-4, what=f"logAdd utt {b}")
        checked += 1
    ba.close()
    # lexicon decoder with a token-level n-gram LM
    W = 60
    path = str(tmp_path / "tok.arpa")
    synth.write_arpa(path, N, order=3, counts=[0, 600, 400], seed=5)
    toks = synth.word_names(N)
    spell = synth.lexicon(W, N, 2, 4, seed=7, exclude=(0, N - 1))
    lm = KenLM(path, Dictionary(toks))
    trie = Trie(N, 0)
    for w, sp in enumerate(spell):
        trie.insert([int(x) for x in sp], w, 0.0)
    trie.smear(SmearingMode.MAX)
    lopts = LexiconDecoderOptions(beam_size=20, beam_size_token=N, beam_threshold=1e9, lm_weight=0.7,
                                  word_score=0.4, unk_score=-math.inf, sil_score=0.0, log_add=False,
                                  criterion_type=CriterionType.CTC)
    ldec = LexiconDecoder(lopts, trie, lm, 0, N - 1, W, [], True)
    bl = Built(A, spec_lexicon(N, 20, N, spell, 1e9, lm_weight=0.7, word_score=0.4, lm=("arpa", path, toks),
                               unk=W, is_lm_token=True))
    for b in range(len(em)):
        e = np.ascontiguousarray(em[b])
        ra = bl.decode(e)
        if has_ties(ra) or A.tie_events(bl.dec):
            continue
        assert_same_nbest(ra, _as_res(ldec.decode(e.ctypes.data, T, N), T), 1e-4, what=f"token LM utt {b}")
        checked += 1
    bl.close()
    assert checked >= 2


def test_cpp_mirror_program(tmp_path):
    """tests/cpp/mirror_test.cpp is written against fl::lib::text like the reference's DecoderTest."""
    exe = str(tmp_path / "mirror_test")
    lib = os.path.join(ROOT, "text_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"), "-o", exe,
                    f"-L{lib}", "-lflt_decoder", f"-Wl,-rpath,{lib}"], check=True)
    N, T, W = 30, 40, 200
    spell = synth.lexicon(W, N, 2, 4, seed=7, exclude=(0, N - 1))
    em = synth.emissions(1, T, N, seed=11, sigma=2.0)[0]
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        np.array([N, T, W], np.int32).tofile(f)
        for sp in spell:
            np.array([len(sp)] + [int(x) for x in sp], np.int32).tofile(f)
        em.astype(np.float32).tofile(f)
    out = subprocess.run([exe, str(inp)], check=True, capture_output=True, text=True).stdout
    got = json.loads(out)
    A = po.Oracle("ora")
    ba = Built(A, spec_lexicon(N, 20, N, spell, 1e9, word_score=0.3))
    ra = ba.decode(em)
    res = dict(n=len(got), scores=np.array([[g["score"], g["am"], g["lm"]] for g in got]).reshape(len(got), 3),
               tokens=np.array([g["tokens"] for g in got], np.int32).reshape(len(got), T + 2),
               words=np.array([g["words"] for g in got], np.int32).reshape(len(got), T + 2))
    assert_same_nbest(ra, res, 1e-4, what="C++ mirror")
    ba.close()


def test_lexfree_decoder_pickle_and_getters():
    """LexiconFreeDecoder pickling (bindings/python/flashlight/lib/text/_decoder.cpp:409-441): a decoder
    without state and with a ZeroLM round-trips through pickle and decodes the same; one that has decoded,
    or that carries another LM, refuses with the reference's messages."""
    import pickle

    from flashlight.lib.text.decoder import CriterionType, LexiconFreeDecoder, LexiconFreeDecoderOptions, ZeroLM

    N, T = 32, 25
    opts = LexiconFreeDecoderOptions(beam_size=8, beam_size_token=N, beam_threshold=1e9, lm_weight=0.0,
                                     sil_score=-0.2, log_add=False, criterion_type=CriterionType.CTC)
    dec = LexiconFreeDecoder(opts, ZeroLM(), 0, N - 1, [])
    assert dec.get_sil_idx() == 0 and dec.get_blank_idx() == N - 1 and dec.get_transitions() == []
    assert dec.get_options().beam_size == 8 and dec.get_options().sil_score == -0.2
    twin = pickle.loads(pickle.dumps(dec))
    assert twin.get_blank_idx() == N - 1 and twin.get_options().beam_size_token == N
    em = np.ascontiguousarray(synth.emissions(1, T, N, seed=3)[0])
    a, b = dec.decode(em.ctypes.data, T, N), twin.decode(em.ctypes.data, T, N)
    assert len(a) == len(b) > 0
    for x, y in zip(a, b):
        assert x.tokens == y.tokens and x.score == y.score
    with pytest.raises(RuntimeError, match="has state"):
        pickle.dumps(dec)

"""Table files (text_b200/csrc/table_io.h): a built Trie and the hashed n-gram tables written once and loaded
back without repeating the reference's per-process setup (test/decoder/DecoderTest.cpp:126-146: one
Trie::insert per word + smear; lm/KenLM.cpp:32-47: parse the LM file). A loaded object must be
indistinguishable from the one that was saved: same nodes / scores / labels, same LM scores, same decode."""
import os
import sys
import time

import numpy as np
import pytest

from cases import Built, assert_same_nbest, spec_lexicon
from flt_backend import FltBackend
from oracle import pyoracle as po
from text_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    return capi.Api()


def _trie(api, N=40, words=800, seed=5, smear=capi.SMEAR_MAX):
    sp = synth.lexicon(words, N, 1, 5, seed=seed, exclude=(0, N - 1))
    rng = np.random.default_rng(seed)
    t = api.trie_create(N, 0)
    for w, s in enumerate(sp):
        api.trie_insert(t, s, w % 700, float(rng.uniform(-6, 0)))  # some words share a label id
    api.trie_smear(t, smear)
    return t, sp


def test_trie_round_trip(api, tmp_path):
    t, sp = _trie(api)
    path = str(tmp_path / "lexicon.flt")
    api.trie_save(t, path)
    t2 = api.trie_load(path)
    a, b = api.trie_export(t), api.trie_export(t2)
    assert a["maxChildren"] == b["maxChildren"] and a["rootIdx"] == b["rootIdx"]
    for k in ("childOff", "childTok", "childNode", "labelOff", "labels", "scores", "maxScore"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    for s in sp[:100]:
        x, y = api.trie_search(t, s), api.trie_search(t2, s)
        assert x["maxScore"] == y["maxScore"] and list(x["labels"]) == list(y["labels"])
        np.testing.assert_array_equal(x["scores"], y["scores"])
    # a loaded Trie is an ordinary Trie: it can still be extended and smeared before a decoder uses it
    api.trie_insert(t2, [1, 2, 3, 4, 5, 6, 7], 9, -1.0)
    api.trie_smear(t2, capi.SMEAR_MAX)
    assert api.trie_num_nodes(t2) > api.trie_num_nodes(t)
    api.trie_destroy(t), api.trie_destroy(t2)


def test_trie_file_errors(api, tmp_path):
    bad = tmp_path / "bad.flt"
    bad.write_bytes(b"not a table file at all")
    with pytest.raises(capi.FltError) as e:
        api.trie_load(str(bad))
    assert e.value.code == capi.ERR_RUNTIME and "not a flt table file" in e.value.msg
    with pytest.raises(capi.FltError):
        api.trie_load(str(tmp_path / "missing.flt"))
    # a truncated file is rejected, not half-loaded
    t, _ = _trie(api, words=200)
    path = str(tmp_path / "t.flt")
    api.trie_save(t, path)
    data = open(path, "rb").read()
    open(path, "wb").write(data[: len(data) // 2])
    with pytest.raises(capi.FltError):
        api.trie_load(path)
    api.trie_destroy(t)


def _arpa(tmp_path, vocab=300, counts=(0, 2000, 1500, 800)):
    path = str(tmp_path / "lm.arpa")
    synth.write_arpa(path, vocab, order=len(counts), counts=list(counts), seed=3)
    return path, synth.word_names(vocab) + ["<unk>"]


def test_lm_round_trip(api, tmp_path):
    path, words = _arpa(tmp_path)
    lm = api.lm_arpa(path, words)
    tbl = str(tmp_path / "lm.flt")
    api.lm_save(lm, tbl)
    # the ARPA entry point recognises the table file by its magic (as KenLM's constructor does with its binaries)
    lm2 = api.lm_arpa(tbl, words)
    rng = np.random.default_rng(0)
    for _ in range(50):
        seq = rng.integers(0, len(words), size=int(rng.integers(1, 9)))
        np.testing.assert_array_equal(api.lm_score_seq(lm, seq, True), api.lm_score_seq(lm2, seq, True))
    # another user vocabulary (different order, an OOV word) maps through the stored LM vocabulary
    words3 = list(reversed(words[:50])) + ["never-seen-word"]
    lm3, lm4 = api.lm_arpa(path, words3), api.lm_arpa(tbl, words3)
    seq = np.arange(len(words3))
    np.testing.assert_array_equal(api.lm_score_seq(lm3, seq, True), api.lm_score_seq(lm4, seq, True))
    for h in (lm, lm2, lm3, lm4):
        api.lm_destroy(h)
    z = api.lm_zero()
    with pytest.raises(capi.FltError):
        api.lm_save(z, str(tmp_path / "zero.flt"))
    api.lm_destroy(z)
    # a Trie file is not an LM
    t, _ = _trie(api, words=50)
    tp = str(tmp_path / "t.flt")
    api.trie_save(t, tp)
    with pytest.raises(capi.FltError):
        api.lm_arpa(tp, words)
    api.trie_destroy(t)


def test_decode_from_loaded_tables_equals_built(tmp_path):
    """Logic harness (tests/model, the kernel sources compiled for the CPU): decoding with a Trie and an LM that
    came from table files gives the bits of decoding with the freshly built ones, which equal the oracle's."""
    G, A = FltBackend("model"), po.Oracle("ora")
    N, W = 24, 150
    path, words = _arpa(tmp_path, vocab=W, counts=(0, 900, 500))
    sp = synth.lexicon(W, N, 1, 4, seed=9, exclude=(0, N - 1))
    spec = spec_lexicon(N, 12, N, sp, 50.0, lm_weight=1.5, word_score=0.4, lm=("arpa", path, words), unk=W)
    em = synth.emissions(3, 25, N, seed=4)
    ba, bg = Built(A, spec), Built(G, spec)
    want = [ba.decode(em[b]) for b in range(len(em))]
    got = bg.O.decode_batch(bg.dec, em, 12)
    lp, tp = str(tmp_path / "lm.flt"), str(tmp_path / "trie.flt")
    G.api.lm_save(bg.lm, lp), G.api.trie_save(bg.trie, tp)
    lm2, trie2 = G.api.lm_arpa(lp, words), G.api.trie_load(tp)
    dec2 = G.decoder_lexicon(spec["opt"], trie2, lm2, spec["sil"], spec["blank"], spec["unk"])
    got2 = G.decode_batch(dec2, em, 12)
    for b in range(len(em)):
        assert_same_nbest(want[b], got[b], 1e-4, what=f"built {b}")
        assert_same_nbest(got[b], got2[b], 0.0, what=f"loaded {b}")
    G.decoder_destroy(dec2), G.api.trie_destroy(trie2), G.api.lm_destroy(lm2)
    ba.close(), bg.close()


def test_mirror_and_pybind_surface(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "text_b200", "compat"))
    import flashlight.lib.text.decoder as D
    from flashlight.lib.text.dictionary import Dictionary

    N = 30
    sp = synth.lexicon(200, N, 1, 4, seed=3, exclude=(0, N - 1))
    t = D.Trie(N, 0)
    for w, s in enumerate(sp):
        t.insert([int(x) for x in s], w, -0.01 * w)
    t.smear(D.SmearingMode.MAX)
    p = str(tmp_path / "t.flt")
    t.save(p)
    t2 = D.Trie.load(p)

    def walk(a, b):  # the host node tree of a loaded Trie can be walked like the built one's
        assert a.idx == b.idx and a.max_score == b.max_score and list(a.labels) == list(b.labels)
        assert list(a.scores) == list(b.scores) and sorted(a.children) == sorted(b.children)
        return 1 + sum(walk(a.children[k], b.children[k]) for k in a.children)

    assert walk(t.get_root(), t2.get_root()) > 200
    assert t2.search([int(x) for x in sp[7]]).labels == t.search([int(x) for x in sp[7]]).labels
    path, words = _arpa(tmp_path, vocab=60, counts=(0, 300))
    d = Dictionary()
    for w in words:
        d.add_entry(w)
    lm = D.KenLM(path, d)
    lp = str(tmp_path / "lm.flt")
    lm.save(lp)
    lm2 = D.KenLM(lp, d)
    s1, s2 = lm.start(False), lm2.start(False)
    for w in (3, 17, 5, 59):
        s1, a = lm.score(s1, w)
        s2, b = lm2.score(s2, w)
        assert a == b
    assert lm.finish(s1)[1] == lm2.finish(s2)[1]


def test_load_is_faster_than_rebuilding(api, tmp_path):
    """what the file is for: a 20 k-word lexicon loads faster than it inserts (host-only, no GPU)"""
    N, W = 200, 20000
    sp = synth.lexicon(W, N, 2, 6, seed=1, exclude=(0, N - 1))
    t0 = time.perf_counter()
    t = api.trie_create(N, 0)
    for w, s in enumerate(sp):
        api.trie_insert(t, s, w, 0.0)
    api.trie_smear(t, capi.SMEAR_MAX)
    build = time.perf_counter() - t0
    p = str(tmp_path / "big.flt")
    api.trie_save(t, p)
    t0 = time.perf_counter()
    t2 = api.trie_load(p)
    load = time.perf_counter() - t0
    assert api.trie_num_nodes(t2) == api.trie_num_nodes(t)
    assert load < build, (load, build)
    api.trie_destroy(t), api.trie_destroy(t2)


@pytest.mark.gpu
def test_decode_from_loaded_tables_cuda(tmp_path):
    """the same on the device: tables loaded from their files are flattened / uploaded like built ones"""
    G, A = FltBackend("cuda"), po.Oracle("ora")
    N, W = 64, 400
    path, words = _arpa(tmp_path, vocab=W, counts=(0, 2500, 1200))
    sp = synth.lexicon(W, N, 1, 4, seed=9, exclude=(0, N - 1))
    spec = spec_lexicon(N, 20, N, sp, 50.0, lm_weight=1.5, word_score=0.4, lm=("arpa", path, words), unk=W)
    em = synth.emissions(4, 40, N, seed=4)
    ba, bg = Built(A, spec), Built(G, spec)
    want = [ba.decode(em[b]) for b in range(len(em))]
    got = bg.O.decode_batch(bg.dec, em, 20)
    lp, tp = str(tmp_path / "lm.flt"), str(tmp_path / "trie.flt")
    G.api.lm_save(bg.lm, lp), G.api.trie_save(bg.trie, tp)
    lm2, trie2 = G.api.lm_arpa(lp, words), G.api.trie_load(tp)
    dec2 = G.decoder_lexicon(spec["opt"], trie2, lm2, spec["sil"], spec["blank"], spec["unk"])
    got2 = G.decode_batch(dec2, em, 20)
    for b in range(len(em)):
        assert_same_nbest(want[b], got[b], 1e-4, what=f"built {b}")
        assert_same_nbest(got[b], got2[b], 0.0, what=f"loaded {b}")
    G.decoder_destroy(dec2), G.api.trie_destroy(trie2), G.api.lm_destroy(lm2)
    ba.close(), bg.close()

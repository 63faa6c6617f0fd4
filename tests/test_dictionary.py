"""Host setup path (text_b200/csrc/host/flashlight_dictionary.h through the pybind11 module): the
known answers of the reference's own dictionary tests
(flashlight/lib/text/test/dictionary/DictionaryTest.cpp:18-175 — basic / file / replabel / UTF-8
vectors, restated here as literals) and the lexicon -> Trie one-shot builder against the
entry-by-entry setup of DecoderTest.cpp:126-146."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text_b200", "compat"))


@pytest.fixture(scope="module")
def D():
    import flashlight.lib.text.dictionary as d

    return d


def test_basic(D):  # DictionaryTest.cpp:18-43
    d = D.Dictionary()
    d.add_entry("1", 1)
    d.add_entry("2", 2)
    d.add_entry("3", 3)
    d.add_entry("4", 3)
    assert d.get_entry(1) == "1" and d.get_entry(3) == "3"
    assert d.get_index("2") == 2 and d.get_index("4") == 3
    assert (d.entry_size(), d.index_size()) == (4, 3)
    d.add_entry("5")
    assert d.get_index("5") == 4 and d.entry_size() == 5
    d.add_entry("6")
    assert d.get_index("6") == 5 and d.index_size() == 5
    with pytest.raises(ValueError):
        d.add_entry("6")
    with pytest.raises(ValueError):
        d.get_index("nope")
    d.set_default_index(2)
    assert d.get_index("nope") == 2


def test_from_file(D, tmp_path):  # DictionaryTest.cpp:45-55: entries on one line share an index
    with pytest.raises(RuntimeError):
        D.Dictionary("not_a_real_file")
    p = tmp_path / "dict.txt"
    p.write_text("a\nb\nc x y\nd\ne z\nf\ng\n")
    d = D.Dictionary(str(p))
    assert (d.entry_size(), d.index_size()) == (10, 7)
    assert d.contains("a") and not d.contains("q")
    assert d.get_entry(1) == "b" and d.get_index("e") == 4 and d.get_index("z") == 4
    assert d.is_contiguous()
    assert d.map_entries_to_indices(["a", "x", "g"]) == [0, 2, 6]
    assert d.map_indices_to_entries([0, 2, 6]) == ["a", "c", "g"]


def test_pack_unpack_replabels(D):  # DictionaryTest.cpp:82-101
    d = D.Dictionary()
    for i in (1, 2, 3):
        d.add_entry(f"<{i}>", i)
    labels = [5, 6, 6, 6, 10, 8, 8, 10, 10, 10, 10, 10]
    want = [labels, [5, 6, 1, 6, 10, 8, 1, 10, 1, 10, 1, 10], [5, 6, 2, 10, 8, 1, 10, 2, 10, 1],
            [5, 6, 2, 10, 8, 1, 10, 3, 10]]
    for i in range(4):
        packed = D.pack_replabels(labels, d, i)
        assert packed == want[i]
        assert D.unpack_replabels(packed, d, i) == labels


def test_unpack_replabels_vectors(D):  # DictionaryTest.cpp:103-152
    d = D.Dictionary()
    for i in (1, 2, 3):
        d.add_entry(f"<{i}>", i)
    for i in (1, 2, 3):
        d.add_entry(str(i), 3 + i)
    labels = [6, 3, 7, 2, 8, 0, 1]
    assert D.unpack_replabels(labels, d, 1) == [6, 3, 7, 2, 8, 0, 0]
    assert D.unpack_replabels(labels, d, 2) == [6, 3, 7, 7, 7, 8, 0, 0]
    assert D.unpack_replabels(labels, d, 3) == [6, 6, 6, 6, 7, 7, 7, 8, 0, 0]
    d2 = D.Dictionary()
    d2.add_entry("<1>", 1)
    d2.add_entry("<2>", 2)
    d2.add_entry("1", 3)
    d2.add_entry("2", 4)
    assert D.unpack_replabels([1, 5, 1, 6], d2, 2) == [5, 5, 6]
    assert D.unpack_replabels([1, 5, 1, 2, 6], d2, 2) == [5, 5, 6]
    assert D.unpack_replabels([1, 5, 1, 2, 6], d2, 1) == [5, 5, 2, 6]
    assert D.unpack_replabels([5, 1, 2, 1, 2, 6], d2, 2) == [5, 5, 6]


def test_utf8_split(D):  # DictionaryTest.cpp:154-175
    assert D.split_wrd("Vendetta") == list("Vendetta")
    assert D.split_wrd("Beyoncé") == ["B", "e", "y", "o", "n", "c", "é"]
    assert D.split_wrd("Beyoncé") == ["B", "e", "y", "o", "n", "c", "e", "́"]


def test_load_words_and_build_trie(D, tmp_path):
    import flashlight.lib.text.decoder as dec

    tokens = ["|", "a", "b", "c", "<1>"]
    token_dict = D.Dictionary(tokens)
    lex = tmp_path / "lexicon.txt"
    lex.write_text("ab a b |\nabb a b b |\ncab c a b |\nab a b\n")
    lexicon = D.load_words(str(lex))
    assert set(lexicon) == {"ab", "abb", "cab", "<unk>"} and len(lexicon["ab"]) == 2
    word_dict = D.create_word_dict(lexicon)
    assert word_dict.get_index("never-seen") == word_dict.get_index("<unk>")
    assert D.tkn_to_idx(["a", "b", "b", "|"], token_dict, 1) == [1, 2, 4, 0]
    with pytest.raises(RuntimeError):
        bad = tmp_path / "bad.txt"
        bad.write_text("lonely\n")
        D.load_words(str(bad))
    trie = D.build_trie(lexicon, token_dict, word_dict, dec.ZeroLM(), 0, 1, dec.SmearingMode.MAX)
    node = trie.search([1, 2, 4, 0])  # "abb" with the repeat packed
    assert node is not None and list(node.labels) == [word_dict.get_index("abb")]
    assert trie.search([1, 2]) is not None and list(trie.search([1, 2]).labels) == [word_dict.get_index("ab")]
    assert trie.search([3, 3]) is None

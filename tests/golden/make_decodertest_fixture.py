#!/usr/bin/env python
"""Generate tests/golden/decodertest_fixture.npz: the inputs of the reference's only decoder golden
test (flashlight/lib/text/test/decoder/DecoderTest.cpp:57-195 — emission / transition matrices,
token list, lexicon, 3-gram ARPA), in a compact derived form (lexicon spellings already mapped to
token indices with one replabel, dictionary/Utils.cpp:90-121; the ARPA text zlib-compressed), plus
the n-best list the UNMODIFIED compiled reference (oracle/_ref) produces for DecoderTest's decoder
configuration. Run in the build container only (reads /root/reference/flashlight/lib/text/test/
decoder/data); the .npz lets the CUDA path run that test where the reference tree is absent."""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reffix  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    fx = reffix.load()
    words, spell_flat, spell_off, spell_word = fx["words"], [], [0], []
    for w, sps in fx["lexicon"].items():
        for sp in sps:
            idx = reffix.tkn2idx(sp, fx["tok2idx"], 1)
            spell_flat += idx
            spell_off.append(len(spell_flat))
            spell_word.append(fx["word2idx"][w])
    arpa = open(fx["arpa"], "rb").read()
    po.build("ref")
    R = po.Oracle("ref")
    lm = R.lm_arpa(fx["arpa"], words)
    sil, unk = fx["tok2idx"]["|"], fx["word2idx"]["<unk>"]
    trie = R.trie_create(len(fx["tokens"]), sil)
    for k, wi in enumerate(spell_word):
        sc = float(R.lm_score_seq(lm, [wi])[0])
        R.trie_insert(trie, spell_flat[spell_off[k]:spell_off[k + 1]], wi, sc)
    R.trie_smear(trie, po.SMEAR_MAX)
    opt = po.make_options(2500, 25000, 100.0, 2.0, 2.0, float("-inf"), -1.0, False, po.ASG)
    dec = R.decoder_lexicon(opt, trie, lm, sil, -1, unk, fx["transitions"], False)
    r = R.decode(dec, fx["emissions"], 2500)
    assert r["n"] == 16, r["n"]  # DecoderTest.cpp:184
    out = dict(emissions=fx["emissions"], transitions=fx["transitions"], tokens=np.array(fx["tokens"]),
               words=np.array(words), spell_flat=np.array(spell_flat, np.int32),
               spell_off=np.array(spell_off, np.int32), spell_word=np.array(spell_word, np.int32),
               arpa_z=np.frombuffer(zlib.compress(arpa, 9), np.uint8), sil=np.int32(sil), unk=np.int32(unk),
               ref_scores=r["scores"], ref_tokens=r["tokens"].astype(np.int32), ref_words=r["words"].astype(np.int32))
    path = os.path.join(HERE, "decodertest_fixture.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB; top-5 {r['scores'][:5, 0]}")


if __name__ == "__main__":
    main()

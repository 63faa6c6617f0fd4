"""BASELINE.json's shapes (N = 10000 tokens) for the committed full-size golden vectors
(tests/golden/fullsize_nbest.npz, made by tests/golden/make_fullsize.py from the compiled reference /
the oracle port) and the tests that compare the CUDA path with them (tests/test_gpu_golden_fullsize.py).

Inputs are regenerated from seeds at test time: emissions by `synth.emissions_exact` (integer hashing
and IEEE add / multiply only, so the fp32 bits are the same on every machine; a CRC in the golden file
guards that), the 200k-word lexicon and the synthetic 4-gram ARPA by `synth.lexicon` / `synth.write_arpa`.

`bst = beam` rows run the UNMODIFIED reference over the full length; `bst = N` rows (token pruning off,
the reference allocates ~66 MB of LMState per frame there) run the oracle port — which is itself pinned
bit-equal to the reference — over a prefix."""
import os

from cases import spec_lexfree, spec_lexicon
from text_b200 import synth

N = 10000
W = 200000
NGRAMS = [0, 500000, 500000, 250000]

# name, kind, beam, bst, T, utterances, beamThreshold, oracle, emission seed
ROWS = [
    ("cfg2_bstK", "lexfree", 50, 50, 1000, 6, 1e9, "ref", 101),
    ("cfg2_bstN", "lexfree", 50, N, 128, 4, 1e9, "ora", 102),
    ("cfg3_bstK", "lexicon", 100, 100, 1000, 6, 1e9, "ref", 103),
    ("cfg3_bstN", "lexicon", 100, N, 128, 4, 1e9, "ora", 104),
    ("cfg4_bstK", "lexicon_lm", 200, 200, 1500, 4, 25.0, "ref", 105),
    ("cfg4_bstN", "lexicon_lm", 200, N, 64, 4, 25.0, "ora", 106),
    ("cfg5_bstK", "lexicon_lm", 500, 500, 300, 2, 25.0, "ref", 107),
    ("cfg5_bstN", "lexicon_lm", 500, N, 40, 2, 25.0, "ora", 108),
]

_cache = {}


def lexicon_200k():
    if "sp" not in _cache:
        _cache["sp"] = synth.lexicon(W, N, 2, 5, seed=7, exclude=(0, N - 1))
    return _cache["sp"]


def arpa_4gram():
    path = os.path.join(synth.cache_dir(), f"golden4_{W}_{'_'.join(map(str, NGRAMS))}.arpa")
    if not os.path.exists(path):
        tmp = path + f".tmp{os.getpid()}"
        synth.write_arpa(tmp, W, order=4, counts=NGRAMS, seed=11)
        os.replace(tmp, path)
    return path


def build(row):
    """(spec, emissions [B,T,N]) of one row."""
    name, kind, beam, bst, T, B, thr, _, seed = row
    if kind == "lexfree":
        spec = spec_lexfree(N, beam, bst, thr, sil=0, blank=N - 1)
    elif kind == "lexicon":
        spec = spec_lexicon(N, beam, bst, lexicon_200k(), thr, sil=0, blank=N - 1, unk=W)
    else:
        spec = spec_lexicon(N, beam, bst, lexicon_200k(), thr, sil=0, blank=N - 1, unk=W, lm_weight=2.0,
                            lm=("arpa", arpa_4gram(), synth.word_names(W) + ["<unk>"]))
    return spec, synth.emissions_exact(B, T, N, seed=seed)

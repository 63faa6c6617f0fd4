"""Kernel LOGIC vs the oracle, without a GPU: the product's kernel sources compiled by g++ as a
one-thread-per-CTA sequential program (tests/model, -DFLT_HOST_MODEL) must reproduce the oracle's
n-best lists on every parity case. This catches algorithmic errors (candidate pruning bounds,
merge keys, arithmetic order) in the build container; races and CUDA-specific behaviour are covered
by the same cases run on the B200 in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import parity_cases
from cases import Built, assert_close_nbest, assert_same_nbest, has_ties
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


def long_ragged_case():
    """Utterances spanning several backtrace checkpoints (every 32 history rows), lengths on and
    around the checkpoint boundaries."""
    from cases import spec_lexfree, spec_lexicon
    from text_b200 import synth

    N, T = 24, 200
    lens = np.array([200, 33, 64, 97, 31, 32, 0, 1], np.int32)
    em = synth.emissions(len(lens), T, N, seed=10, sigma=2.0)
    specs = [spec_lexfree(N, 8, N, 1e9),
             spec_lexicon(N, 12, N, synth.lexicon(60, N, 2, 4, seed=2, exclude=(0, N - 1)), 1e9, word_score=0.3)]
    return specs, em, lens


def run_long_ragged(A, G, tol):
    specs, em, lens = long_ragged_case()
    checked = 0
    for spec in specs:
        ba, bg = Built(A, spec), Built(G, spec)
        got = bg.O.decode_batch(bg.dec, em, spec["opt"].beamSize, lens)
        for b in range(len(lens)):
            ra = ba.decode(em[b][: lens[b]])
            if has_ties(ra) or A.tie_events(ba.dec):
                continue
            assert_same_nbest(ra, got[b], tol, what=f"utt {b} len {lens[b]}")
            checked += 1
        ba.close(), bg.close()
    assert checked >= 8


def test_long_ragged_checkpoints(A, M):
    run_long_ragged(A, M, 1e-9)


def masked_cases():
    """Emission rows with -inf entries (masked tokens): fewer finite values than list entries wanted
    (short token lists), rows with a handful of finite values, and heavy masking with a lexicon."""
    from cases import spec_lexfree, spec_lexicon
    from text_b200 import synth

    out = []
    rng = np.random.default_rng(4)
    for N, K, frac in ((64, 50, 0.4), (64, 10, 0.9), (200, 50, 0.5), (40, 16, 0.2)):
        em = synth.emissions(2, 30, N, seed=N + K, sigma=2.0)
        mask = rng.random((2, 30, N)) < frac
        mask[:, :, 0] = False      # sil stays finite
        mask[:, :, N - 1] = False  # blank stays finite
        em = em.copy()
        em[mask] = -np.inf
        out.append((spec_lexfree(N, K, N, 1e9), em))
    N = 40
    em = synth.emissions(2, 30, N, seed=77, sigma=2.0).copy()
    mask = rng.random((2, 30, N)) < 0.3
    mask[:, :, 0] = False
    mask[:, :, N - 1] = False
    em[mask] = -np.inf
    out.append((spec_lexicon(N, 20, N, synth.lexicon(150, N, 2, 4, seed=3, exclude=(0, N - 1)), 1e9, word_score=0.2), em))
    return out


def run_masked(A, G, tol):
    checked = 0
    for spec, em in masked_cases():
        ba, bg = Built(A, spec), Built(G, spec)
        got = bg.O.decode_batch(bg.dec, em, spec["opt"].beamSize)
        for b, e in enumerate(em):
            ra = ba.decode(e)
            if has_ties(ra) or A.tie_events(ba.dec):
                continue
            assert_same_nbest(ra, got[b], tol, what=f"masked utt {b}")
            checked += 1
        ba.close(), bg.close()
    assert checked >= 4


def test_masked_emissions(A, M):
    run_masked(A, M, 1e-9)


def run_widened(A, G, spec, em, exact, tol):
    """logAdd / token-LM modes (parity_cases.widened_cases). `exact`: strings and scores as strict as
    everywhere else; otherwise (logAdd) near-equal neighbours may swap and scores carry `tol`."""
    ba, bg = Built(A, spec), Built(G, spec)
    got = bg.O.decode_batch(bg.dec, em, spec["opt"].beamSize)
    checked = 0
    for b, e in enumerate(em):
        ra = ba.decode(e)
        if A.tie_events(ba.dec) or has_ties(ra):
            continue
        if exact:
            assert_same_nbest(ra, got[b], tol, what=f"utt {b}")
        else:
            assert_close_nbest(ra, got[b], max(tol, 1e-9), what=f"utt {b}")
        checked += 1
    ba.close(), bg.close()
    assert checked, "all utterances had tie events: vacuous"


@pytest.mark.parametrize("name,spec,em,exact", parity_cases.widened_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_widened_modes(A, M, name, spec, em, exact):
    # same libm on both sides here: logAdd scores are bit-equal too
    run_widened(A, M, spec, em, True, 1e-9)


def test_lexicon_split_workspace(A, M, monkeypatch):
    """The generic step with its capacity-sized arrays (candidate records, merge table, representatives) in a
    region of their own, as the device places them in a global slab when a wide beam outgrows shared memory
    (FLT_TEST_HYBRID makes the logic harness do the same): results unchanged. Includes a word-LM case, whose
    n-gram probes are bounded before they are made, and a histogram-ranked select (groups > K)."""
    monkeypatch.setenv("FLT_TEST_HYBRID", "1")
    cases = parity_cases.lexicon_cases()
    ran = 0
    for name, spec, em in cases[::3] + [c for c in cases if c[1]["lm"][0] == "arpa"][:3]:
        run_case(A, M, spec, em)
        ran += 1
    assert ran >= 4

"""Parity tests proper: the shipped CUDA library (text_b200/lib/libflt_decoder.so, through the
C-ABI) against the CPU oracle on the same seeded inputs. Token / word strings bit-equal, the three
scores within 1e-4 (BASELINE.json north_star); in practice the scores are bit-equal too because the
kernels keep the reference's FP64/FP32 evaluation order and are built with -fmad=false."""
import numpy as np
import pytest

import parity_cases
from cases import Built, assert_same_nbest, has_ties
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from flt_backend import FltBackend

    return FltBackend("cuda")


@pytest.fixture(scope="module")
def A():
    return po.Oracle("ora")


def run_case(A, G, spec, em, lengths=None):
    ba, bg = Built(A, spec), Built(G, spec)
    K = spec["opt"].beamSize
    got = bg.O.decode_batch(bg.dec, em, K, lengths)
    checked = 0
    for b, e in enumerate(em):
        ra = ba.decode(e if lengths is None else e[: lengths[b]])
        if has_ties(ra):
            continue
        assert_same_nbest(ra, got[b], 1e-4, what=f"utt {b}")
        assert np.array_equal(ra["scores"], got[b]["scores"]), "scores not bit-equal"
        checked += 1
    ba.close(), bg.close()
    assert checked, "all utterances had score ties: vacuous"


@pytest.mark.parametrize("name,spec,em", parity_cases.lexfree_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_lexfree(A, G, name, spec, em):
    run_case(A, G, spec, em)


@pytest.mark.parametrize("name,spec,em", parity_cases.lexicon_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_lexicon(A, G, name, spec, em):
    run_case(A, G, spec, em)


def test_ragged_lengths_and_many_utterances(A, G):
    """More utterances than resident CTAs, ragged lengths (incl. 0 frames)."""
    from cases import spec_lexfree
    from text_b200 import synth

    N, T, B = 64, 50, 700
    em = synth.emissions(B, T, N, seed=77)
    lengths = np.random.default_rng(3).integers(0, T + 1, size=B).astype(np.int32)
    lengths[:4] = [0, 1, T, T - 1]
    spec = spec_lexfree(N, 16, N, 1e9)
    run_case(A, G, spec, em, lengths)


def test_long_ragged_checkpoints(A, G):
    from test_model_parity import run_long_ragged

    run_long_ragged(A, G, 1e-4)


def test_masked_emissions(A, G):
    from test_model_parity import run_masked

    run_masked(A, G, 1e-4)


@pytest.mark.parametrize("name,spec,em,exact", parity_cases.widened_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_widened_modes(A, G, name, spec, em, exact):
    """logAdd merging, token-level n-gram LM (lexicon-free), isLmToken (lexicon): full expansion on the
    device. Without logAdd the scores are bit-equal; with it exp/log1p differ from libm in the last
    bit, so scores carry the north star's 1e-4 and near-equal neighbours may swap."""
    from test_model_parity import run_widened

    run_widened(A, G, spec, em, exact, 1e-4)
    if exact:
        ba, bg = Built(A, spec), Built(G, spec)
        ra, rg = ba.decode(em[0]), bg.decode(em[0], spec["opt"].beamSize)
        if not (A.tie_events(ba.dec) or has_ties(ra)):
            assert np.array_equal(ra["scores"], rg["scores"]), "scores not bit-equal"
        ba.close(), bg.close()


def test_lexicon_split_workspace(A, G, monkeypatch):
    """The generic step with its capacity-sized arrays in the CTA's global slab and the small region in shared
    memory (what a wide beam or a grown candidate capacity gets on the device), forced here on small cases by
    lowering the shared-memory budget of the plan (FLT_SMEM_KB): same bits as everything in shared memory."""
    monkeypatch.setenv("FLT_SMEM_KB", "16")
    ran = 0
    for name, spec, em in parity_cases.lexicon_cases():
        if name in ("cfg3_scaled_bstN", "arpa3_ctc", "zero_unk", "arpa3_ctc_bst", "arpa_asg", "cfg4_scaled"):
            run_case(A, G, spec, em)
            ran += 1
    assert ran >= 3


def test_very_wide_beam(A, G, tmp_path):
    """beam 500 (BASELINE configs[4]): the small workspace region leaves one CTA per SM, which then runs the
    step with 1024 threads (flt_k_decode1024), the capacity-sized arrays in its global slab, the long-list
    select (M = 505) and the histogram-rank select; lexicon + 3-gram LM and lexicon + ZeroLM."""
    from cases import spec_lexicon
    from text_b200 import synth

    N, W, T = 640, 3000, 24
    sp = synth.lexicon(W, N, 2, 4, seed=21, exclude=(0, N - 1))
    path = str(tmp_path / "lm3.arpa")
    synth.write_arpa(path, W, order=3, counts=[0, 6000, 3000], seed=5)
    em = synth.emissions(3, T, N, seed=31, sigma=1.5)
    for lm, lw in ((("arpa", path, synth.word_names(W) + ["<unk>"]), 1.5), (("zero",), 0.0)):
        spec = spec_lexicon(N, 500, N, sp, 40.0, lm_weight=lw, word_score=0.2, lm=lm, unk=W)
        run_case(A, G, spec, em)

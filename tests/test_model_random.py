"""Randomised kernel-logic parity (GPU-less harness, tests/model): the product's kernel sources as a
sequential program against the oracle over randomly drawn decoder configurations — vocabulary and
beam sizes, token beams, thresholds, CTC / ASG, sil / word / unk scores, ZeroLM / ARPA LM, ragged
lengths. Exercises the exactness of every pruning device (corner bound, column budgets, in-place
merge resolution of the lexicon-free step, two-pass histogram pruning of the lexicon step, the
fused kernel's host flow). Utterances where the oracle reports a tie event are skipped
(SURVEY.md §0.4). The same configurations run on the GPU in tests/test_gpu_random.py."""
import os

import numpy as np
import pytest

from cases import Built, assert_close_nbest, assert_same_nbest, has_ties, spec_lexfree, spec_lexicon
from oracle import pyoracle as po
from text_b200 import synth


def draw_lexfree(rng):
    N = int(rng.integers(1, 60)) * 4 if rng.random() < 0.5 else int(rng.integers(3, 200))
    T = int(rng.integers(1, 40))
    K = int(rng.choice([1, 2, 3, 5, 10, 17, 50, 64, 65, 100, 128, 200, 256]))
    bst = int(rng.choice([N, max(1, N // 2), min(N, K), min(N, 5), min(N, K + 3)]))
    thr = float(rng.choice([1e9, 25.0, 8.0, 2.0]))
    crit = int(rng.choice([po.CTC, po.CTC, po.ASG]))
    tr = rng.random(N * N, dtype=np.float32) if crit == po.ASG else None
    spec = spec_lexfree(N, K, bst, thr, sil=0, blank=(N - 1 if crit == po.CTC else -1), criterion=crit,
                        sil_score=float(rng.choice([0.0, 0.0, -0.5, 0.7])), transitions=tr)
    em = synth.emissions(3, T, N, seed=int(rng.integers(1 << 30)), sigma=float(rng.choice([0.3, 1.0, 2.0, 4.0, 8.0])))
    lengths = rng.integers(0, T + 1, size=3).astype(np.int32) if rng.random() < 0.3 else None
    return spec, em, lengths


_ARPA = {}


def arpa():
    if not _ARPA:
        path = os.path.join(synth.cache_dir(), "rand3.arpa")
        if not os.path.exists(path):
            synth.write_arpa(path, 300, order=3, counts=[0, 3000, 2000], seed=3)
        _ARPA["v"] = (path, synth.word_names(300) + ["<unk>"])
    return _ARPA["v"]


def draw_lexicon(rng):
    N = int(rng.integers(8, 120))
    T = int(rng.integers(1, 40))
    K = int(rng.choice([1, 3, 10, 30, 100, 200]))
    mn, mx = int(rng.choice([1, 2])), int(rng.choice([2, 3, 5]))
    cap = sum((N - 2) ** l for l in range(mn, mx + 1))
    W = min(int(rng.choice([20, 100, 300])), max(1, cap // 2))
    bst = int(rng.choice([N, max(1, N // 2), min(N, K), min(N, 6)]))
    crit = int(rng.choice([po.CTC, po.CTC, po.ASG]))
    use_arpa = rng.random() < 0.5
    path, words = arpa()
    sp = synth.lexicon(W, N, mn, mx, seed=int(rng.integers(1000)), exclude=(0, N - 1))
    tr = rng.random(N * N, dtype=np.float32) if crit == po.ASG else None
    spec = spec_lexicon(N, K, bst, sp, float(rng.choice([1e9, 25.0, 6.0])), sil=0,
                        blank=(N - 1 if crit == po.CTC else -1), criterion=crit,
                        sil_score=float(rng.choice([0.0, -0.5, 0.4])),
                        lm_weight=float(rng.choice([0.5, 2.0])) if use_arpa else 0.0,
                        word_score=float(rng.choice([0.0, 0.5, -1.0])),
                        unk_score=float(rng.choice([float("-inf"), -2.0])),
                        lm=("arpa", path, words) if use_arpa else ("zero",), transitions=tr,
                        unk=300 if use_arpa else W)
    em = synth.emissions(2, T, N, seed=int(rng.integers(1 << 30)), sigma=float(rng.choice([1.0, 2.0, 4.0])))
    return spec, em, None


def draw_widened(rng):
    """logAdd merging and token-level LMs (full expansion on the device), both decoders."""
    log_add = bool(rng.random() < 0.6)
    path, words = arpa()
    crit = int(rng.choice([po.CTC, po.CTC, po.ASG]))
    T = int(rng.integers(1, 30))
    thr = float(rng.choice([1e9, 25.0, 8.0]))
    if rng.random() < 0.5:
        N = int(rng.integers(3, 120))
        K = int(rng.choice([1, 3, 10, 30, 64, 100]))
        token_lm = bool(rng.random() < 0.5) or not log_add
        bst = int(rng.choice([N, max(1, N // 2), min(N, K), min(N, 5)]))
        tr = rng.random(N * N, dtype=np.float32) if crit == po.ASG else None
        spec = spec_lexfree(N, K, bst, thr, sil=0, blank=(N - 1 if crit == po.CTC else -1), criterion=crit,
                            log_add=log_add, sil_score=float(rng.choice([0.0, -0.5, 0.7])),
                            lm_weight=float(rng.choice([0.5, 1.5])) if token_lm else 0.0,
                            lm=("arpa", path, words[:N]) if token_lm else ("zero",), transitions=tr)
    else:
        N = int(rng.integers(8, 80))
        K = int(rng.choice([1, 3, 10, 30, 100]))
        mn, mx = int(rng.choice([1, 2])), int(rng.choice([3, 5]))
        W = int(rng.choice([20, 60, 200]))
        W = min(W, max(1, sum((N - 2) ** l for l in range(mn, mx + 1)) // 2))
        bst = int(rng.choice([N, max(1, N // 2), min(N, 6)]))
        mode = rng.choice(["zero", "word", "token", "token_zero"]) if log_add else rng.choice(["token", "token_zero"])
        token_lm = mode in ("token", "token_zero")
        lm = ("zero",) if mode in ("zero", "token_zero") else ("arpa", path, words[:N] if token_lm else words)
        sp = synth.lexicon(W, N, mn, mx, seed=int(rng.integers(1000)), exclude=(0, N - 1))
        tr = rng.random(N * N, dtype=np.float32) if crit == po.ASG else None
        spec = spec_lexicon(N, K, bst, sp, thr, sil=0, blank=(N - 1 if crit == po.CTC else -1),
                            criterion=crit, log_add=log_add, sil_score=float(rng.choice([0.0, -0.5, 0.4])),
                            lm_weight=float(rng.choice([0.5, 2.0])) if lm[0] == "arpa" else 0.0,
                            word_score=float(rng.choice([0.0, 0.53, -1.1])),
                            unk_score=float(rng.choice([float("-inf"), -2.3])), lm=lm, transitions=tr,
                            unk=300 if mode == "word" else W, is_lm_token=token_lm)
    em = synth.emissions(3, T, N, seed=int(rng.integers(1 << 30)), sigma=float(rng.choice([1.0, 2.0, 4.0])))
    lengths = rng.integers(0, T + 1, size=3).astype(np.int32) if rng.random() < 0.3 else None
    return spec, em, lengths


def run_random(A, G, draw, seed, rounds, tol):
    rng = np.random.default_rng(seed)
    checked = 0
    for _ in range(rounds):
        spec, em, lengths = draw(rng)
        ba, bg = Built(A, spec), Built(G, spec)
        got = bg.O.decode_batch(bg.dec, em, spec["opt"].beamSize, lengths)
        for b, e in enumerate(em):
            ra = ba.decode(e if lengths is None else e[: lengths[b]])
            if has_ties(ra) or A.tie_events(ba.dec):
                continue
            what = f"seed {seed} spec {spec['opt'].beamSize}/{spec['N']}"
            if spec["opt"].logAdd and tol > 1e-9:  # device exp/log1p vs libm: last-bit differences
                assert_close_nbest(ra, got[b], tol, what=what)
            else:
                assert_same_nbest(ra, got[b], tol, what=what)
            checked += 1
        ba.close(), bg.close()
    assert checked > rounds // 2


@pytest.fixture(scope="module")
def M():
    from flt_backend import FltBackend

    return FltBackend("model")


@pytest.fixture(scope="module")
def A():
    return po.Oracle("ora")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_lexfree_random(A, M, seed):
    run_random(A, M, draw_lexfree, seed, 40, 1e-9)


@pytest.mark.parametrize("seed", [10, 11])
def test_lexicon_random(A, M, seed):
    run_random(A, M, draw_lexicon, seed, 30, 1e-9)


@pytest.mark.parametrize("seed", [20, 21])
def test_widened_random(A, M, seed):
    run_random(A, M, draw_widened, seed, 40, 1e-9)


@pytest.mark.parametrize("seed", [12, 22])
def test_random_split_workspace(A, M, seed, monkeypatch):
    """the same draws with the capacity-sized arrays in a region of their own (what the device does when a wide
    beam outgrows shared memory; FLT_TEST_HYBRID makes the logic harness do it for every configuration):
    lexicon decoder incl. n-gram LMs, and the full-expansion modes"""
    monkeypatch.setenv("FLT_TEST_HYBRID", "1")
    run_random(A, M, draw_lexicon if seed < 20 else draw_widened, seed, 25, 1e-9)

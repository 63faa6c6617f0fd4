"""The parity case list shared by the GPU tests (product CUDA library vs oracle) and the CPU
logic-harness tests (tests/model build of the same kernels vs oracle)."""
import os

import numpy as np

from cases import spec_lexfree, spec_lexicon
from oracle import pyoracle as po
from text_b200 import synth

NEG_INF = float("-inf")


def _arpa(name, W, order, counts, seed):
    path = os.path.join(synth.cache_dir(), name)
    if not os.path.exists(path):
        synth.write_arpa(path, W, order=order, counts=counts, seed=seed)
    return path, synth.word_names(W) + ["<unk>"]


def lexfree_cases():
    # name, N, T, B, beam, bst, thr, criterion, sil_score, sigma
    rows = [
        ("cfg1_shape", 29, 200, 2, 10, 29, 1e9, po.CTC, 0.0, 1.0),
        ("bst_small", 29, 80, 2, 10, 5, 1e9, po.CTC, 0.0, 1.0),
        ("thr_silneg", 29, 80, 2, 10, 29, 8.0, po.CTC, -0.5, 1.0),
        ("sil_positive", 40, 60, 2, 12, 40, 1e9, po.CTC, 0.7, 1.0),
        ("sil_positive_bst", 40, 60, 2, 12, 9, 20.0, po.CTC, 0.7, 2.0),
        ("asg", 40, 60, 2, 12, 40, 1e9, po.ASG, 0.0, 1.0),
        ("asg_bst_thr", 40, 60, 2, 12, 6, 20.0, po.ASG, -1.0, 1.0),
        ("cfg2_scaled_bstN", 500, 40, 2, 50, 500, 1e9, po.CTC, 0.0, 1.0),
        ("cfg2_scaled_bstK", 500, 40, 2, 50, 50, 1e9, po.CTC, 0.0, 1.0),
        ("chunk_path_bstN", 5000, 12, 2, 50, 5000, 1e9, po.CTC, 0.0, 1.0),
        ("chunk_path_bst300", 5000, 12, 2, 20, 300, 1e9, po.CTC, 0.0, 1.0),
        ("peaky", 300, 60, 2, 30, 300, 25.0, po.CTC, 0.0, 4.0),
        ("tiny_vocab", 3, 30, 2, 4, 3, 1e9, po.CTC, 0.0, 1.0),
        ("one_frame", 10, 1, 2, 5, 10, 1e9, po.CTC, 0.0, 1.0),
        ("beam1", 20, 30, 2, 1, 20, 1e9, po.CTC, 0.0, 1.0),
        ("beam_gt_cands", 6, 10, 2, 200, 6, 1e9, po.CTC, 0.0, 1.0),
    ]
    out = []
    for name, N, T, B, beam, bst, thr, crit, sil_score, sigma in rows:
        tr = np.random.default_rng(5).random(N * N, dtype=np.float32) if crit == po.ASG else None
        spec = spec_lexfree(N, beam, bst, thr, sil=0, blank=(N - 1 if crit == po.CTC else -1),
                            criterion=crit, sil_score=sil_score, transitions=tr)
        em = synth.emissions(B, T, N, seed=1000 + len(out), sigma=sigma)
        out.append((name, spec, em))
    return out


def lexicon_cases():
    # name, N, T, B, W, (minlen,maxlen), beam, bst, thr, criterion, sil_score, word_score, unk_score, lm
    rows = [
        # (ZeroLM cases: spellings of one length, or a vocabulary large enough, so that two word
        #  segmentations of one token string — equal scores, libstdc++-internal order in the reference —
        #  do not make every utterance a tie)
        ("zero_ctc", 30, 60, 2, 200, (3, 3), 20, 30, 1e9, po.CTC, 0.0, 0.0, NEG_INF, "zero"),
        ("zero_ctc_bst_thr", 30, 60, 2, 200, (2, 4), 20, 8, 15.0, po.CTC, -0.2, 1.0, NEG_INF, "zero"),
        ("zero_ctc_silpos", 30, 60, 2, 200, (2, 4), 20, 30, 1e9, po.CTC, 0.4, 0.3, NEG_INF, "zero"),
        ("zero_ctc_words1", 120, 60, 2, 300, (1, 3), 20, 120, 1e9, po.CTC, 0.0, 0.37, NEG_INF, "zero"),
        ("zero_unk", 60, 50, 2, 150, (3, 3), 10, 60, 1e9, po.CTC, 0.0, 0.0, -4.0, "zero"),
        ("zero_asg", 120, 50, 3, 300, (3, 3), 20, 120, 1e9, po.ASG, -0.3, 0.7, NEG_INF, "zero"),
        ("cfg3_scaled_bstN", 200, 40, 2, 2000, (2, 4), 50, 200, 1e9, po.CTC, 0.0, 0.0, NEG_INF, "zero"),
        ("cfg3_scaled_bstK", 200, 40, 2, 2000, (2, 4), 50, 50, 25.0, po.CTC, 0.0, 0.0, NEG_INF, "zero"),
        ("cfg3_mid", 2000, 30, 2, 20000, (2, 5), 100, 2000, 1e9, po.CTC, 0.0, 0.0, NEG_INF, "zero"),
        ("arpa3_ctc", 40, 80, 2, 300, (1, 3), 40, 40, 30.0, po.CTC, 0.0, 0.5, NEG_INF, "arpa4"),
        ("arpa3_ctc_bst", 40, 80, 2, 300, (1, 3), 40, 12, 30.0, po.CTC, -0.1, 0.5, NEG_INF, "arpa4"),
        ("arpa_asg", 40, 60, 2, 300, (1, 3), 40, 40, 30.0, po.ASG, 0.0, 0.5, NEG_INF, "arpa4"),
        ("arpa_unk", 40, 60, 2, 300, (2, 3), 30, 40, 30.0, po.CTC, 0.0, 0.5, -3.0, "arpa4"),
        ("cfg4_scaled", 300, 60, 2, 3000, (2, 5), 200, 300, 25.0, po.CTC, 0.0, 0.0, NEG_INF, "arpa4big"),
    ]
    out = []
    for (name, N, T, B, W, (mn, mx), beam, bst, thr, crit, sil_score, word_score, unk_score,
         lm) in rows:
        sp = synth.lexicon(W, N, mn, mx, seed=7, exclude=(0, N - 1))
        if lm == "zero":
            lmspec, lmw = ("zero",), 0.0
        elif lm == "arpa4":
            path, words = _arpa("p_small4.arpa", 300, 4, [0, 3000, 3000, 2000], 3)
            lmspec, lmw = ("arpa", path, words), 1.5
        else:
            path, words = _arpa("p_mid4.arpa", 3000, 4, [0, 30000, 30000, 20000], 4)
            lmspec, lmw = ("arpa", path, words), 2.0
        tr = np.random.default_rng(5).random(N * N, dtype=np.float32) if crit == po.ASG else None
        spec = spec_lexicon(N, beam, bst, sp, thr, sil=0, blank=(N - 1 if crit == po.CTC else -1),
                            criterion=crit, sil_score=sil_score, lm_weight=lmw,
                            word_score=word_score, unk_score=unk_score, lm=lmspec, transitions=tr,
                            unk=W)
        em = synth.emissions(B, T, N, seed=2000 + len(out), sigma=2.0)
        out.append((name, spec, em))
    return out


def widened_cases():
    """Modes without rank dominance (full expansion on the device): logAdd merging, the token-level
    n-gram LM of the lexicon-free decoder, isLmToken in the lexicon decoder. `exact` = scores are
    expected bit-equal on the GPU too (no exp/log1p involved)."""
    path, words = _arpa("p_small4.arpa", 300, 4, [0, 3000, 3000, 2000], 3)
    out = []

    def lexfree(name, N, T, beam, bst, thr, crit, log_add, sil_score, lm, lmw, sigma):
        tr = np.random.default_rng(5).random(N * N, dtype=np.float32) if crit == po.ASG else None
        lmspec = ("zero",) if lm == "zero" else ("arpa", path, words[:N])
        spec = spec_lexfree(N, beam, bst, thr, sil=0, blank=(N - 1 if crit == po.CTC else -1),
                            criterion=crit, log_add=log_add, sil_score=sil_score, lm_weight=lmw,
                            lm=lmspec, transitions=tr)
        em = synth.emissions(3, T, N, seed=3000 + len(out), sigma=sigma)
        out.append((name, spec, em, not log_add))

    lexfree("lf_logadd_ctc", 29, 80, 10, 29, 1e9, po.CTC, True, 0.0, "zero", 0.0, 1.0)
    lexfree("lf_logadd_bst_thr", 40, 60, 12, 9, 12.0, po.CTC, True, -0.3, "zero", 0.0, 2.0)
    lexfree("lf_logadd_asg", 40, 60, 12, 40, 1e9, po.ASG, True, 0.0, "zero", 0.0, 1.0)
    lexfree("lf_logadd_wide", 500, 30, 50, 500, 25.0, po.CTC, True, 0.0, "zero", 0.0, 3.0)
    lexfree("lf_tokenlm_ctc", 40, 60, 15, 40, 1e9, po.CTC, False, 0.0, "arpa", 0.8, 1.0)
    lexfree("lf_tokenlm_bst_thr", 40, 60, 15, 10, 20.0, po.CTC, False, 0.2, "arpa", 1.3, 2.0)
    lexfree("lf_tokenlm_asg", 40, 60, 15, 40, 1e9, po.ASG, False, 0.0, "arpa", 0.8, 1.0)
    lexfree("lf_tokenlm_logadd", 40, 60, 15, 40, 30.0, po.CTC, True, 0.0, "arpa", 0.8, 2.0)

    def lexicon(name, N, T, W, lens, beam, bst, thr, crit, log_add, sil_score, word_score, unk_score,
                lm, lmw, token_lm):
        sp = synth.lexicon(W, N, lens[0], lens[1], seed=7, exclude=(0, N - 1))
        tr = np.random.default_rng(5).random(N * N, dtype=np.float32) if crit == po.ASG else None
        if lm == "zero":
            lmspec = ("zero",)
        else:
            lmspec = ("arpa", path, words[:N] if token_lm else words)
        spec = spec_lexicon(N, beam, bst, sp, thr, sil=0, blank=(N - 1 if crit == po.CTC else -1),
                            criterion=crit, log_add=log_add, sil_score=sil_score, lm_weight=lmw,
                            word_score=word_score, unk_score=unk_score, lm=lmspec, transitions=tr,
                            unk=W, is_lm_token=token_lm)
        em = synth.emissions(3, T, N, seed=4000 + len(out), sigma=2.0)
        out.append((name, spec, em, not log_add))

    lexicon("lex_logadd_zero", 30, 60, 200, (2, 4), 20, 30, 1e9, po.CTC, True, 0.0, 0.3, NEG_INF, "zero", 0.0, False)
    lexicon("lex_logadd_bst_thr", 30, 60, 200, (1, 4), 20, 8, 15.0, po.CTC, True, -0.2, 1.0, NEG_INF, "zero", 0.0, False)
    lexicon("lex_logadd_arpa", 40, 60, 300, (1, 3), 40, 40, 30.0, po.CTC, True, 0.0, 0.5, NEG_INF, "arpa", 1.5, False)
    lexicon("lex_logadd_asg_unk", 40, 50, 300, (2, 3), 30, 40, 30.0, po.ASG, True, 0.0, 0.5, -3.0, "arpa", 1.5, False)
    lexicon("lex_tokenlm_arpa", 30, 60, 200, (2, 4), 20, 30, 1e9, po.CTC, False, 0.0, 0.4, NEG_INF, "arpa", 0.7, True)
    lexicon("lex_tokenlm_bst_unk", 30, 40, 100, (2, 4), 20, 10, 25.0, po.CTC, False, -0.1, 0.43, -3.7, "arpa", 0.7, True)
    lexicon("lex_tokenlm_zero", 30, 60, 200, (2, 4), 20, 30, 1e9, po.CTC, False, 0.0, 0.4, NEG_INF, "zero", 0.0, True)
    lexicon("lex_tokenlm_logadd", 30, 40, 60, (2, 4), 20, 30, 40.0, po.ASG, True, 0.0, 0.4, NEG_INF, "arpa", 0.7, True)
    return out

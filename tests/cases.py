"""Shared builders for parity tests: one `spec` dict describes a decoder configuration and is
instantiated on any backend exposing the oracle_api surface (ref_*, ora_*) — and, in the GPU
tests, on the product's C-ABI — so every side sees identical inputs."""
import numpy as np

from oracle import pyoracle as po


def spec_lexfree(N, beam, bst, thr=1e9, sil=0, blank=None, criterion=po.CTC, log_add=False,
                 sil_score=0.0, lm_weight=0.0, lm=("zero",), transitions=None):
    return dict(kind="lexfree", N=N, sil=sil, blank=(N - 1 if blank is None else blank), unk=-1,
                opt=po.make_options(beam, bst, thr, lm_weight, 0.0, float("-inf"), sil_score,
                                    log_add, criterion),
                lm=lm, transitions=transitions)


def spec_lexicon(N, beam, bst, spellings, thr=1e9, sil=0, blank=None, criterion=po.CTC,
                 log_add=False, sil_score=0.0, lm_weight=0.0, word_score=0.0,
                 unk_score=float("-inf"), lm=("zero",), transitions=None, smear=po.SMEAR_MAX,
                 word_ids=None, unk=None, is_lm_token=False):
    W = len(spellings)
    return dict(kind="lexicon", N=N, sil=sil, blank=(N - 1 if blank is None else blank),
                unk=(W if unk is None else unk), spellings=spellings,
                word_ids=(list(range(W)) if word_ids is None else word_ids),
                opt=po.make_options(beam, bst, thr, lm_weight, word_score, unk_score, sil_score,
                                    log_add, criterion),
                lm=lm, transitions=transitions, smear=smear, is_lm_token=is_lm_token)


class Built:
    def __init__(self, O, spec):
        self.O, self.spec = O, spec
        lm = spec["lm"]
        if lm[0] == "zero":
            self.lm = O.lm_zero()
        else:  # ("arpa", path, usr_words)
            self.lm = O.lm_arpa(lm[1], lm[2])
        self.trie = None
        if spec["kind"] == "lexicon":
            self.trie = O.trie_create(spec["N"], spec["sil"])
            for sp, wid in zip(spec["spellings"], spec["word_ids"]):
                word_lm = lm[0] != "zero" and not spec["is_lm_token"]
                sc = float(O.lm_score_seq(self.lm, [wid])[0]) if word_lm else 0.0
                O.trie_insert(self.trie, sp, wid, sc)
            O.trie_smear(self.trie, spec["smear"])
            self.dec = O.decoder_lexicon(spec["opt"], self.trie, self.lm, spec["sil"],
                                         spec["blank"], spec["unk"], spec["transitions"],
                                         spec["is_lm_token"])
        else:
            self.dec = O.decoder_lexfree(spec["opt"], self.lm, spec["sil"], spec["blank"],
                                         spec["transitions"])

    def decode(self, emissions, max_hyp=None):
        return self.O.decode(self.dec, emissions, max_hyp or self.spec["opt"].beamSize)

    def close(self):
        self.O.decoder_destroy(self.dec)
        if self.trie:
            self.O.trie_destroy(self.trie)
        self.O.lm_destroy(self.lm)


def has_ties(res, tol=0.0):
    """Equal adjacent final scores => the n-best order/cut is implementation-defined."""
    s = res["scores"][:, 0]
    return bool(len(s) > 1 and (np.abs(np.diff(s)) <= tol).any())


def assert_same_nbest(a, b, score_tol=1e-4, what=""):
    assert a["n"] == b["n"], f"{what}: n-best size {a['n']} vs {b['n']}"
    np.testing.assert_array_equal(a["tokens"], b["tokens"], err_msg=f"{what}: tokens")
    np.testing.assert_array_equal(a["words"], b["words"], err_msg=f"{what}: words")
    np.testing.assert_allclose(a["scores"], b["scores"], rtol=0, atol=score_tol,
                               err_msg=f"{what}: scores")


def assert_close_nbest(a, b, score_tol=1e-4, near=1e-9, what=""):
    """n-best comparison for logAdd decoding, whose exp/log1p may round differently on the two
    sides: scores within `score_tol`; token / word strings equal position by position, except that
    neighbours whose oracle scores lie within `near` of each other may be swapped."""
    assert a["n"] == b["n"], f"{what}: n-best size {a['n']} vs {b['n']}"
    sa = a["scores"][:, 0]
    for i in range(a["n"]):
        cand = [j for j in range(a["n"]) if abs(sa[j] - sa[i]) <= near * max(1.0, abs(sa[i]))]
        ok = any(np.array_equal(a["tokens"][j], b["tokens"][i]) and
                 np.array_equal(a["words"][j], b["words"][i]) and
                 np.allclose(a["scores"][j], b["scores"][i], rtol=0, atol=score_tol) for j in cand)
        assert ok, f"{what}: rank {i} differs (tokens / words / scores)"

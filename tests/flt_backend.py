"""Adapter giving the product's C-ABI (text_b200.capi.Api) the same surface as oracle.pyoracle.Oracle
so `cases.Built` can instantiate one spec on the oracle and on the product.

`kind="cuda"` loads the shipped CUDA library (GPU tests). `kind="model"` loads tests/model's g++
build of the same kernel sources (a GPU-less logic harness, NOT a product path)."""
import os
import subprocess

import numpy as np

from text_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
MODEL_LIB = os.path.join(_HERE, "model", "libflt_model.so")


def build_model():
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "model")], check=True)


class FltBackend:
    def __init__(self, kind="cuda"):
        if kind == "model":
            build_model()
            self.api = capi.Api(MODEL_LIB)
        else:
            self.api = capi.Api()
        self.kind = kind
        self.table_cache = False  # bench.py: keep the hashed n-gram tables next to the ARPA file (table_io.h)
        self.setup_times = {}
        a = self.api
        for name in ("trie_create", "trie_insert", "trie_smear", "trie_search", "trie_destroy",
                     "lm_zero", "lm_score_seq", "lm_destroy", "decoder_destroy"):
            setattr(self, name, getattr(a, name))

    def lm_arpa(self, path, words):
        import time

        tbl = path + ".flt"
        t0 = time.perf_counter()
        if self.table_cache and os.path.exists(tbl):
            h = self.api.lm_arpa(tbl, words)
            self.setup_times["lm_table_file_load_s"] = time.perf_counter() - t0
            return h
        h = self.api.lm_arpa(path, words)
        self.setup_times["lm_arpa_parse_s"] = time.perf_counter() - t0
        if self.table_cache:
            tmp = f"{tbl}.tmp{os.getpid()}"
            self.api.lm_save(h, tmp)
            os.replace(tmp, tbl)
        return h

    @staticmethod
    def _opt(o):
        return capi.Options(o.beamSize, o.beamSizeToken, o.beamThreshold, o.lmWeight, o.wordScore,
                            o.unkScore, o.silScore, o.logAdd, o.criterion)

    def decoder_lexfree(self, opt, lm, sil, blank, transitions=None):
        return self.api.decoder_lexfree(self._opt(opt), lm, sil, blank, transitions)

    def decoder_lexicon(self, opt, trie, lm, sil, blank, unk, transitions=None, is_lm_token=False):
        return self.api.decoder_lexicon(self._opt(opt), trie, lm, sil, blank, unk, transitions,
                                        is_lm_token)

    def decode(self, dec, emissions, max_hyp):
        """single utterance [T,N] -> same dict as pyoracle.Oracle.decode"""
        return self.decode_batch(dec, np.asarray(emissions)[None], max_hyp)[0]

    def decode_batch(self, dec, emissions, max_hyp, lengths=None):
        B, T, N = self.api.decode_batch(dec, emissions, lengths)
        r = self.api.nbest(dec, B, T, max_hyp)
        out = []
        for b in range(B):
            n = min(int(r["counts"][b]), max_hyp)
            L = T + 2 if lengths is None else int(lengths[b]) + 2
            out.append(dict(n=n, scores=r["scores"][b, :n], tokens=r["tokens"][b, :n, :L],
                            words=r["words"][b, :n, :L], total=int(r["counts"][b])))
        return out

    # ---- online decoding, same surface as pyoracle.Oracle
    def decode_begin(self, dec):
        self._stream_N = None
        self._stream_dec = dec

    def decode_step(self, dec, emissions):
        e = np.ascontiguousarray(emissions, np.float32)
        if self._stream_N is None:
            self._stream_N = e.shape[1]
            self.api.stream_begin(dec, e.shape[1])
        self.api.stream_step(dec, e)

    def decode_end(self, dec):
        self.api.stream_end(dec)

    def prune(self, dec, look_back=0):
        self.api.stream_prune(dec, look_back)

    def n_hypothesis(self, dec):
        return self.api.stream_n_hypothesis(dec)

    def n_frames_in_buffer(self, dec):
        return self.api.stream_frames_in_buffer(dec)

    def best(self, dec, look_back, max_len):
        return self.api.stream_best(dec, look_back, max_len)

    def all_final(self, dec, max_hyp, max_len):
        return self.api.stream_all_final(dec, max_hyp, max_len)

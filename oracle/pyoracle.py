"""ORACLE (test infrastructure, never shipped): ctypes front-end to oracle/oracle_api.h.

`Oracle("ref")` binds oracle/_ref/libflref.so (the unmodified reference, compiled in place);
`Oracle("ora")` binds oracle/liboracle.so (our CPU restatement). Same methods on both, so a
test can run the two side by side. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ASG, CTC = 0, 1
SMEAR_NONE, SMEAR_MAX, SMEAR_LOGADD = 0, 1, 2


class Options(C.Structure):
    _fields_ = [
        ("beamSize", C.c_int32),
        ("beamSizeToken", C.c_int32),
        ("beamThreshold", C.c_double),
        ("lmWeight", C.c_double),
        ("wordScore", C.c_double),
        ("unkScore", C.c_double),
        ("silScore", C.c_double),
        ("logAdd", C.c_int32),
        ("criterion", C.c_int32),
    ]


def make_options(beam_size, beam_size_token, beam_threshold, lm_weight=0.0, word_score=0.0,
                 unk_score=float("-inf"), sil_score=0.0, log_add=False, criterion=CTC):
    return Options(beam_size, beam_size_token, beam_threshold, lm_weight, word_score, unk_score,
                   sil_score, int(bool(log_add)), int(criterion))


def lib_path(kind):
    return os.path.join(_HERE, "_ref", "libflref.so") if kind == "ref" else os.path.join(
        _HERE, "liboracle.so")


def available(kind):
    return os.path.exists(lib_path(kind))


def build(kind=None):
    """Compile the oracle libraries (building the checker is not using it)."""
    targets = ["oracle", "ref"] if kind is None else [kind if kind != "ora" else "oracle"]
    for t in targets:
        subprocess.run(["make", "-s", "-C", _HERE, t], check=True)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    def __init__(self, kind="ora"):
        assert kind in ("ref", "ora")
        path = lib_path(kind)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle`")
        self.kind = kind
        self.lib = C.CDLL(path)
        p = kind
        L = self.lib
        vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)

        def fn(name, res, args):
            f = getattr(L, f"{p}_{name}")
            f.restype = res
            f.argtypes = args
            setattr(self, "_" + name, f)

        fn("trie_create", vp, [C.c_int, C.c_int])
        fn("trie_insert", C.c_int, [vp, ip, C.c_int, C.c_int, C.c_float])
        fn("trie_smear", None, [vp, C.c_int])
        fn("trie_search", C.c_int, [vp, ip, C.c_int, fp, ip, ip, fp])
        fn("trie_destroy", None, [vp])
        fn("lm_zero", vp, [])
        fn("lm_arpa", vp, [C.c_char_p, C.POINTER(C.c_char_p), C.c_int])
        fn("lm_score_seq", C.c_int, [vp, ip, C.c_int, C.c_int, fp])
        fn("lm_destroy", None, [vp])
        fn("decoder_lexfree", vp, [C.POINTER(Options), vp, C.c_int, C.c_int, fp, C.c_int])
        fn("decoder_lexicon", vp,
           [C.POINTER(Options), vp, vp, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int])
        fn("decoder_destroy", None, [vp])
        fn("decode", C.c_int, [vp, fp, C.c_int, C.c_int, C.c_int, dp, ip, ip])
        fn("decode_begin", None, [vp])
        fn("decode_step", None, [vp, fp, C.c_int, C.c_int])
        fn("decode_end", None, [vp])
        fn("prune", None, [vp, C.c_int])
        fn("n_hypothesis", C.c_int, [vp])
        fn("n_frames_in_buffer", C.c_int, [vp])
        fn("best", C.c_int, [vp, C.c_int, C.c_int, dp, ip, ip])
        fn("all_final", C.c_int, [vp, C.c_int, C.c_int, dp, ip, ip, ip])
        fn("bench_mt", C.c_double,
           [C.c_int, C.POINTER(Options), vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int,
            C.c_int, C.c_int, C.c_int, C.c_int])
        fn("last_error", C.c_char_p, [])
        if kind == "ora":
            fn("tie_events", C.c_long, [vp])

    def _err(self):
        return self._last_error().decode()

    # ---- Trie (decoder/Trie.h:66-86)
    def trie_create(self, max_children, root_idx):
        return self._trie_create(max_children, root_idx)

    def trie_insert(self, trie, indices, label, score):
        a = _i32(indices)
        if self._trie_insert(trie, _ptr(a, C.c_int32), len(a), label, score) != 0:
            raise IndexError(self._err())

    def trie_smear(self, trie, mode):
        self._trie_smear(trie, mode)

    def trie_search(self, trie, indices):
        a = _i32(indices)
        ms, nl = C.c_float(), C.c_int32()
        labels = np.zeros(6, np.int32)
        scores = np.zeros(6, np.float32)
        r = self._trie_search(trie, _ptr(a, C.c_int32), len(a), C.byref(ms), C.byref(nl),
                              _ptr(labels, C.c_int32), _ptr(scores, C.c_float))
        if r < 0:
            raise IndexError(self._err())
        if r == 0:
            return None
        return dict(maxScore=ms.value, labels=labels[:nl.value].copy(), scores=scores[:nl.value].copy())

    def trie_destroy(self, trie):
        self._trie_destroy(trie)

    # ---- LM (decoder/lm/LM.h:52-85)
    def lm_zero(self):
        return self._lm_zero()

    def lm_arpa(self, path, words):
        arr = (C.c_char_p * len(words))(*[w.encode() for w in words])
        h = self._lm_arpa(path.encode(), arr, len(words))
        if not h:
            raise RuntimeError(self._err())
        return h

    def lm_score_seq(self, lm, usr_idx, with_finish=False):
        a = _i32(usr_idx)
        out = np.zeros(len(a) + 1, np.float32)
        if self._lm_score_seq(lm, _ptr(a, C.c_int32), len(a), int(with_finish), _ptr(out, C.c_float)) != 0:
            raise RuntimeError(self._err())
        return out if with_finish else out[:-1]

    def lm_destroy(self, lm):
        self._lm_destroy(lm)

    # ---- decoders
    def decoder_lexfree(self, opt, lm, sil, blank, transitions=None):
        tr = np.ascontiguousarray(transitions if transitions is not None else [], np.float32)
        return self._decoder_lexfree(C.byref(opt), lm, sil, blank, _ptr(tr, C.c_float), tr.size)

    def decoder_lexicon(self, opt, trie, lm, sil, blank, unk, transitions=None, is_lm_token=False):
        tr = np.ascontiguousarray(transitions if transitions is not None else [], np.float32)
        return self._decoder_lexicon(C.byref(opt), trie, lm, sil, blank, unk, _ptr(tr, C.c_float),
                                     tr.size, int(is_lm_token))

    def decoder_destroy(self, dec):
        self._decoder_destroy(dec)

    def decode(self, dec, emissions, max_hyp):
        """emissions [T, N] fp32 -> list of dict(score, amScore, lmScore, tokens, words)."""
        e = np.ascontiguousarray(emissions, np.float32)
        T, N = e.shape
        scores = np.zeros((max_hyp, 3), np.float64)
        tokens = np.full((max_hyp, T + 2), -1, np.int32)
        words = np.full((max_hyp, T + 2), -1, np.int32)
        n = self._decode(dec, _ptr(e, C.c_float), T, N, max_hyp, _ptr(scores, C.c_double),
                         _ptr(tokens, C.c_int32), _ptr(words, C.c_int32))
        if n < 0:
            raise RuntimeError(self._err())
        n = min(n, max_hyp)
        return dict(n=n, scores=scores[:n], tokens=tokens[:n], words=words[:n])

    def tie_events(self, dec):
        """restatement only: implementation-defined tie events during the last decode"""
        return int(self._tie_events(dec))

    def decode_begin(self, dec):
        self._decode_begin(dec)

    def decode_step(self, dec, emissions):
        e = np.ascontiguousarray(emissions, np.float32)
        T, N = e.shape
        self._decode_step(dec, _ptr(e, C.c_float), T, N)

    def decode_end(self, dec):
        self._decode_end(dec)

    def prune(self, dec, look_back=0):
        self._prune(dec, look_back)

    def n_hypothesis(self, dec):
        return self._n_hypothesis(dec)

    def n_frames_in_buffer(self, dec):
        return self._n_frames_in_buffer(dec)

    def best(self, dec, look_back, max_len):
        scores = np.zeros(3, np.float64)
        tokens = np.full(max_len, -1, np.int32)
        words = np.full(max_len, -1, np.int32)
        n = self._best(dec, look_back, max_len, _ptr(scores, C.c_double), _ptr(tokens, C.c_int32),
                       _ptr(words, C.c_int32))
        return dict(scores=scores, tokens=tokens[:n], words=words[:n])

    def all_final(self, dec, max_hyp, max_len):
        scores = np.zeros((max_hyp, 3), np.float64)
        tokens = np.full((max_hyp, max_len), -1, np.int32)
        words = np.full((max_hyp, max_len), -1, np.int32)
        lens = np.zeros(max_hyp, np.int32)
        n = self._all_final(dec, max_hyp, max_len, _ptr(scores, C.c_double), _ptr(tokens, C.c_int32),
                            _ptr(words, C.c_int32), _ptr(lens, C.c_int32))
        n = min(n, max_hyp)
        return dict(n=n, scores=scores[:n], tokens=tokens[:n], words=words[:n], lens=lens[:n])

    def bench_mt(self, lexicon, opt, trie, lm, sil, blank, unk, emissions, n_threads, warmup=1,
                 is_lm_token=False):
        e = np.ascontiguousarray(emissions, np.float32)
        B, T, N = e.shape
        return self._bench_mt(int(lexicon), C.byref(opt), trie, lm, sil, blank, unk,
                              int(is_lm_token), _ptr(e, C.c_float), B, T, N, n_threads, warmup)

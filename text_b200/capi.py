"""ctypes binding of the C-ABI in include/flt_decoder.h (text_b200/lib/libflt_decoder.so).

This is the only native entry into the product. There is no CPU implementation behind it: if the
CUDA library is missing or no device is present, calls fail loudly (FltError / OSError).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLT_LIB") or os.path.join(_HERE, "lib", "libflt_decoder.so")

OK, ERR_INVALID, ERR_OUT_OF_RANGE, ERR_RUNTIME, ERR_CUDA, ERR_UNSUPPORTED = range(6)
CRITERION_ASG, CRITERION_CTC = 0, 1
SMEAR_NONE, SMEAR_MAX, SMEAR_LOGADD = 0, 1, 2


class Options(C.Structure):
    """flt_options == LexiconDecoderOptions (decoder/LexiconDecoder.h:21-31)."""
    _fields_ = [
        ("beamSize", C.c_int32),
        ("beamSizeToken", C.c_int32),
        ("beamThreshold", C.c_double),
        ("lmWeight", C.c_double),
        ("wordScore", C.c_double),
        ("unkScore", C.c_double),
        ("silScore", C.c_double),
        ("logAdd", C.c_int32),
        ("criterionType", C.c_int32),
    ]


class FltError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[flt:{code}] {msg}")
        self.code = code
        self.msg = msg


SYMBOLS = {
    # name: (restype, argtypes)
    "flt_last_error": (C.c_char_p, []),
    "flt_trie_create": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "flt_trie_insert": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_float]),
    "flt_trie_smear": (C.c_int, [C.c_void_p, C.c_int32]),
    "flt_trie_search": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32),
                                  C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_float)]),
    "flt_trie_num_nodes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "flt_trie_max_scores": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int64]),
    "flt_trie_destroy": (None, [C.c_void_p]),
    "flt_trie_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "flt_trie_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "flt_trie_export": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "flt_lm_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "flt_lm_zero_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "flt_lm_ngram_load_arpa": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), C.c_int32,
                                         C.POINTER(C.c_void_p)]),
    "flt_lm_score_seq": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                   C.POINTER(C.c_float)]),
    "flt_lm_destroy": (None, [C.c_void_p]),
    "flt_decoder_create_lexfree": (C.c_int, [C.POINTER(Options), C.c_void_p, C.c_int32, C.c_int32,
                                             C.POINTER(C.c_float), C.c_int64, C.c_int32,
                                             C.POINTER(C.c_void_p)]),
    "flt_decoder_create_lexicon": (C.c_int, [C.POINTER(Options), C.c_void_p, C.c_void_p, C.c_int32,
                                             C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_int64,
                                             C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "flt_decoder_destroy": (None, [C.c_void_p]),
    "flt_decoder_set_nbest": (C.c_int, [C.c_void_p, C.c_int32]),
    "flt_decode_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_int32)]),
    "flt_decode_batch_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p]),
    "flt_decoder_synchronize": (C.c_int, [C.c_void_p]),
    "flt_decoder_stream": (C.c_void_p, [C.c_void_p]),
    "flt_nbest_copy": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "flt_stream_begin": (C.c_int, [C.c_void_p, C.c_int32]),
    "flt_stream_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "flt_stream_end": (C.c_int, [C.c_void_p]),
    "flt_stream_prune": (C.c_int, [C.c_void_p, C.c_int32]),
    "flt_stream_frames_in_buffer": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "flt_stream_n_hypothesis": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "flt_stream_best": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "flt_stream_all_final": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32)]),
    "flt_nbest_device_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "flt_decoder_last_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "flt_decoder_workspace_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "flt_decoder_set_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "flt_decoder_last_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "flt_decoder_last_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "flt_topm_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "flt_topm_rows_bias": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
}

_libs = {}


def load(path=None):
    """dlopen the C-ABI library and type its entry points. `path` other than the product library is
    used only by tests (the GPU-less logic harness, tests/model)."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise OSError(f"{path} not found: build it with `make -C text_b200/csrc` "
                      "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(lib, name)  # AttributeError if the library does not export the symbol
        f.restype = res
        f.argtypes = args
    _libs[path] = lib
    return lib


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Api:
    """Object wrapper over the C entry points; raises FltError on non-zero status."""

    def __init__(self, path=None, device=0):
        self.lib = load(path)
        self.device = device

    def _ck(self, code):
        if code != OK:
            raise FltError(code, (self.lib.flt_last_error() or b"").decode())

    # ---- Trie
    def trie_create(self, max_children, root_idx):
        h = C.c_void_p()
        self._ck(self.lib.flt_trie_create(max_children, root_idx, C.byref(h)))
        return h

    def trie_insert(self, trie, indices, label, score):
        a = np.ascontiguousarray(indices, np.int32)
        self._ck(self.lib.flt_trie_insert(trie, _i32p(a), len(a), label, score))

    def trie_smear(self, trie, mode):
        self._ck(self.lib.flt_trie_smear(trie, mode))

    def trie_search(self, trie, indices):
        a = np.ascontiguousarray(indices, np.int32)
        found, ms, nl = C.c_int32(), C.c_float(), C.c_int32()
        labels = np.zeros(6, np.int32)
        scores = np.zeros(6, np.float32)
        self._ck(self.lib.flt_trie_search(trie, _i32p(a), len(a), C.byref(found), C.byref(ms),
                                          C.byref(nl), _i32p(labels), _f32p(scores)))
        if not found.value:
            return None
        return dict(maxScore=ms.value, labels=labels[:nl.value].copy(), scores=scores[:nl.value].copy())

    def trie_num_nodes(self, trie):
        n = C.c_int64()
        self._ck(self.lib.flt_trie_num_nodes(trie, C.byref(n)))
        return n.value

    def trie_max_scores(self, trie):
        """smeared score of every node, in creation order (node 0 = root)"""
        n = self.trie_num_nodes(trie)
        out = np.zeros(n, np.float32)
        self._ck(self.lib.flt_trie_max_scores(trie, out.ctypes.data_as(C.POINTER(C.c_float)), n))
        return out

    def trie_destroy(self, trie):
        self.lib.flt_trie_destroy(trie)

    def trie_save(self, trie, path):
        self._ck(self.lib.flt_trie_save(trie, path.encode()))

    def trie_load(self, path):
        h = C.c_void_p()
        self._ck(self.lib.flt_trie_load(path.encode(), C.byref(h)))
        return h

    def trie_export(self, trie):
        """CSR arrays of the node tree (creation order; edges ascending by token)"""
        meta = np.zeros(5, np.int32)
        nul_i, nul_f = C.POINTER(C.c_int32)(), C.POINTER(C.c_float)()
        self._ck(self.lib.flt_trie_export(trie, _i32p(meta), nul_i, nul_i, nul_i, nul_i, nul_i, nul_f, nul_f))
        nn, ne, nl = int(meta[2]), int(meta[3]), int(meta[4])
        out = dict(childOff=np.zeros(nn + 1, np.int32), childTok=np.zeros(max(ne, 1), np.int32),
                   childNode=np.zeros(max(ne, 1), np.int32), labelOff=np.zeros(nn + 1, np.int32),
                   labels=np.zeros(max(nl, 1), np.int32), scores=np.zeros(max(nl, 1), np.float32),
                   maxScore=np.zeros(nn, np.float32))
        self._ck(self.lib.flt_trie_export(trie, _i32p(meta), _i32p(out["childOff"]), _i32p(out["childTok"]),
                                          _i32p(out["childNode"]), _i32p(out["labelOff"]), _i32p(out["labels"]),
                                          _f32p(out["scores"]), _f32p(out["maxScore"])))
        out["childTok"], out["childNode"] = out["childTok"][:ne], out["childNode"][:ne]
        out["labels"], out["scores"] = out["labels"][:nl], out["scores"][:nl]
        out.update(maxChildren=int(meta[0]), rootIdx=int(meta[1]))
        return out

    # ---- LM
    def lm_zero(self):
        h = C.c_void_p()
        self._ck(self.lib.flt_lm_zero_create(C.byref(h)))
        return h

    def lm_arpa(self, path, words):
        arr = (C.c_char_p * len(words))(*[w.encode() for w in words])
        h = C.c_void_p()
        self._ck(self.lib.flt_lm_ngram_load_arpa(path.encode(), arr, len(words), C.byref(h)))
        return h

    def lm_save(self, lm, path):
        self._ck(self.lib.flt_lm_save(lm, path.encode()))

    def lm_score_seq(self, lm, usr_idx, with_finish=False):
        a = np.ascontiguousarray(usr_idx, np.int32)
        out = np.zeros(len(a) + 1, np.float32)
        self._ck(self.lib.flt_lm_score_seq(lm, _i32p(a), len(a), int(with_finish), _f32p(out)))
        return out if with_finish else out[:-1]

    def lm_destroy(self, lm):
        self.lib.flt_lm_destroy(lm)

    # ---- decoders
    def decoder_lexfree(self, opt, lm, sil, blank, transitions=None):
        tr = np.ascontiguousarray(transitions if transitions is not None else [], np.float32)
        h = C.c_void_p()
        self._ck(self.lib.flt_decoder_create_lexfree(C.byref(opt), lm, sil, blank, _f32p(tr), tr.size,
                                                     self.device, C.byref(h)))
        return h

    def decoder_lexicon(self, opt, trie, lm, sil, blank, unk, transitions=None, is_lm_token=False):
        tr = np.ascontiguousarray(transitions if transitions is not None else [], np.float32)
        h = C.c_void_p()
        self._ck(self.lib.flt_decoder_create_lexicon(C.byref(opt), trie, lm, sil, blank, unk,
                                                     _f32p(tr), tr.size, int(is_lm_token),
                                                     self.device, C.byref(h)))
        return h

    def decoder_destroy(self, dec):
        self.lib.flt_decoder_destroy(dec)

    def set_nbest(self, dec, nbest):
        self._ck(self.lib.flt_decoder_set_nbest(dec, nbest))

    def decode_batch(self, dec, emissions, lengths=None):
        """emissions: C-contiguous fp32 numpy array [B,T,N] (host)."""
        e = np.ascontiguousarray(emissions, np.float32)
        B, T, N = e.shape
        ln = None if lengths is None else np.ascontiguousarray(lengths, np.int32)
        self._ck(self.lib.flt_decode_batch(dec, e.ctypes.data, B, T, N,
                                           None if ln is None else _i32p(ln)))
        return B, T, N

    def decode_batch_ptr(self, dec, ptr, B, T, N, lengths=None):
        """Raw pointer (host or device address) form, as the reference's Python decode takes."""
        ln = None if lengths is None else np.ascontiguousarray(lengths, np.int32)
        self._ck(self.lib.flt_decode_batch(dec, ptr, B, T, N, None if ln is None else _i32p(ln)))

    def decode_batch_async(self, dec, dev_ptr, B, T, N, dev_lengths_ptr=None):
        self._ck(self.lib.flt_decode_batch_async(dec, dev_ptr, B, T, N, dev_lengths_ptr))

    def synchronize(self, dec):
        self._ck(self.lib.flt_decoder_synchronize(dec))

    def stream(self, dec):
        return self.lib.flt_decoder_stream(dec)

    def _pinned_out(self, B, T, nbest):
        """page-locked result buffers, allocated once per shape and reused (numpy views of pinned torch tensors)"""
        import torch

        key = (B, T, nbest)
        cache = self.__dict__.setdefault("_pinned", {})
        if key not in cache:
            cache.clear()  # one shape at a time: these are hundreds of MB
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
            cache[key] = (mk((B, nbest, T + 2), torch.int32), mk((B, nbest, T + 2), torch.int32),
                          mk((B, nbest, 3), torch.float64), mk((B,), torch.int32))
        return tuple(t.numpy() for t in cache[key])

    def nbest(self, dec, B, T, nbest, pinned=False):
        """n-best of the last batch. pinned=True: the arrays are page-locked buffers owned by this Api object and
        overwritten by the next pinned call of the same shape (the device->host copy of a few hundred MB then runs
        at the link rate instead of the pageable-memory rate)."""
        if pinned:
            tokens, words, scores, counts = self._pinned_out(B, T, nbest)
        else:
            tokens = np.empty((B, nbest, T + 2), np.int32)
            words = np.empty((B, nbest, T + 2), np.int32)
            scores = np.empty((B, nbest, 3), np.float64)
            counts = np.empty(B, np.int32)
        self._ck(self.lib.flt_nbest_copy(dec, nbest, _i32p(tokens), _i32p(words),
                                         scores.ctypes.data_as(C.POINTER(C.c_double)), _i32p(counts)))
        return dict(tokens=tokens, words=words, scores=scores, counts=counts)

    # ---- online decoding (one utterance)
    def stream_begin(self, dec, N):
        self._ck(self.lib.flt_stream_begin(dec, N))

    def stream_step(self, dec, emissions):
        e = np.ascontiguousarray(emissions, np.float32)
        T, N = e.shape
        self._ck(self.lib.flt_stream_step(dec, e.ctypes.data, T, N))

    def stream_end(self, dec):
        self._ck(self.lib.flt_stream_end(dec))

    def stream_prune(self, dec, look_back=0):
        self._ck(self.lib.flt_stream_prune(dec, look_back))

    def stream_frames_in_buffer(self, dec):
        n = C.c_int32()
        self._ck(self.lib.flt_stream_frames_in_buffer(dec, C.byref(n)))
        return n.value

    def stream_n_hypothesis(self, dec):
        n = C.c_int32()
        self._ck(self.lib.flt_stream_n_hypothesis(dec, C.byref(n)))
        return n.value

    def stream_best(self, dec, look_back, max_len):
        scores = np.zeros(3, np.float64)
        tokens = np.full(max_len, -1, np.int32)
        words = np.full(max_len, -1, np.int32)
        n = C.c_int32()
        self._ck(self.lib.flt_stream_best(dec, look_back, max_len, _i32p(tokens), _i32p(words),
                                          scores.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return dict(scores=scores, tokens=tokens[:n.value], words=words[:n.value])

    def stream_all_final(self, dec, max_hyp, max_len):
        scores = np.zeros((max_hyp, 3), np.float64)
        tokens = np.full((max_hyp, max_len), -1, np.int32)
        words = np.full((max_hyp, max_len), -1, np.int32)
        lens = np.zeros(max_hyp, np.int32)
        n = C.c_int32()
        self._ck(self.lib.flt_stream_all_final(dec, max_hyp, max_len, _i32p(tokens), _i32p(words),
                                               scores.ctypes.data_as(C.POINTER(C.c_double)), _i32p(lens),
                                               C.byref(n)))
        k = min(n.value, max_hyp)
        return dict(n=k, scores=scores[:k], tokens=tokens[:k], words=words[:k], lens=lens[:k])

    def nbest_device(self, dec, B, T, nbest, K):
        """The last batch's n-best buffers as torch CUDA tensors that alias the decoder's memory
        (no copy): tokens / words [B,nbest,T+2] int32, scores [B,K,3] f64, counts [B] int32."""
        import torch

        p = [C.c_void_p() for _ in range(4)]
        self._ck(self.lib.flt_nbest_device_ptrs(dec, *[C.byref(x) for x in p]))

        class _Cai:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(ptr, False), version=3)

        dev = torch.device("cuda", self.device)
        mk = lambda ptr, shape, ts: torch.as_tensor(_Cai(ptr.value, shape, ts), device=dev)
        return dict(tokens=mk(p[0], (B, nbest, T + 2), "<i4"), words=mk(p[1], (B, nbest, T + 2), "<i4"),
                    scores=mk(p[2], (B, K, 3), "<f8"), counts=mk(p[3], (B,), "<i4"))

    def last_launches(self, dec):
        n = C.c_int32()
        self._ck(self.lib.flt_decoder_last_launches(dec, C.byref(n)))
        return n.value

    def set_timing(self, dec, on=1):
        """0 off, 1 CUDA events per launch, 2 also in-kernel work / phase counters."""
        self._ck(self.lib.flt_decoder_set_timing(dec, int(on)))

    def last_kernel_ms(self, dec):
        ms = (C.c_float * 4)()
        n = (C.c_int32 * 4)()
        self._ck(self.lib.flt_decoder_last_kernel_ms(dec, ms, n))
        names = ("token_select", "beam_step", "backtrace", "fused_select_step")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(names)}

    def last_stats(self, dec):
        v = (C.c_uint64 * 32)()
        self._ck(self.lib.flt_decoder_last_stats(dec, v))
        frames = max(int(v[0]), 1)
        out = dict(frames=int(v[0]), candidates_per_frame=v[1] / frames, groups_per_frame=v[2] / frames,
                   survivors_per_frame=v[3] / frames)
        if v[30]:  # single-pass step with a guessed cut (beam_gx.h)
            names = ("expand", "merge", "compact", "rank_new_beam")
            out["phase_cycles_per_frame"] = {n: round(v[4 + i] / frames, 1) for i, n in enumerate(names)}
            out["phase_cycles_per_frame"]["wait_list"] = round(v[9] / frames, 1)
            out["guess_redo_frames"] = int(v[12])
            out["redo_events"] = dict(overflow=int(v[13]), miss=int(v[14]), gave_up=int(v[15]))
            if v[16]:
                import struct

                out["first_give_up"] = dict(
                    attempt=int(v[16]) - 1000, want=int(v[17]), n_cand=int(v[18]), n_rep=int(v[19]), flags=int(v[20]),
                    cut_bin=int(v[21]), kept=int(v[22]), span=struct.unpack("f", struct.pack("I", int(v[23])))[0],
                    row=int(v[24]), n_hyp=int(v[25]),
                    beam_spread=struct.unpack("d", struct.pack("Q", int(v[26])))[0],
                    g_minus_best=struct.unpack("d", struct.pack("Q", int(v[27])))[0])
            out["select_guess_misses"] = int(v[11])
        elif v[31]:  # generic step (beam_core.h frameStep): SM cycles of thread 0 per phase
            names = ("rows", "degrees_scan", "pass1_histogram", "cut_bin", "pass2_emit", "merge", "select",
                     "new_beam")
            out["phase_cycles_per_frame"] = {n: round(v[4 + i] / frames, 1) for i, n in enumerate(names)}
            # frames decided in one pass against a guessed cut / frames whose guess missed (redone in two passes)
            out["guessed_cut_frames"], out["guessed_cut_misses"] = int(v[12]), int(v[13])
        elif any(v[4:11]):
            names = ("insert", "emit", "scan", "rank", "new_beam", "wait_list", "handover_gather")
            out["phase_cycles_per_frame"] = {n: round(v[4 + i] / frames, 1) for i, n in enumerate(names)}
            out["select_guess_misses"] = int(v[11])
            # frames expanded against the guessed pruning bound / frames whose guess cut too deep (expanded twice)
            out["guessed_bound_frames"], out["guessed_bound_misses"] = int(v[12]), int(v[13])
            out["emit_cycles_per_warp"] = [round(v[16 + i] / frames, 1) for i in range(12) if v[16 + i]]
        return out

    def workspace_bytes(self, dec):
        n = C.c_int64()
        self._ck(self.lib.flt_decoder_workspace_bytes(dec, C.byref(n)))
        return n.value

    def topm_rows(self, dev_emis_ptr, rows, N, M, dev_tok_ptr, dev_val_ptr, stream=None, dev_bias_ptr=None):
        if dev_bias_ptr is None:
            self._ck(self.lib.flt_topm_rows(dev_emis_ptr, rows, N, M, dev_tok_ptr, dev_val_ptr, stream))
        else:
            self._ck(self.lib.flt_topm_rows_bias(dev_emis_ptr, rows, N, M, dev_bias_ptr, dev_tok_ptr, dev_val_ptr,
                                                 stream))

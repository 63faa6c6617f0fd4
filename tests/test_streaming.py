"""Online decoding through the C-ABI (flt_stream_*: decodeBegin / chunked decodeStep /
getBestHypothesis(lookBack) / prune / decodeEnd, decoder/Decoder.h:18-35, decoder/Utils.h:268-342)
against the oracle, which tests/test_oracle_vs_ref.py pins on the compiled reference for the same
call sequence. CPU logic harness here; the CUDA library in the gpu-marked test."""
import numpy as np
import pytest

from cases import Built, spec_lexfree, spec_lexicon
from oracle import pyoracle as po
from text_b200 import synth


def run_streaming(A, G, lexicon, chunk, look_back, prune_every, tol, mode="max"):
    N, T = 30, 90
    em = synth.emissions(1, T, N, seed=91, sigma=2.0)[0]
    log_add = mode in ("logadd", "tokenlm_logadd")
    lm, lmw = ("zero",), 0.0
    if mode.startswith("tokenlm"):  # token-level 4-gram (full expansion on the device)
        import parity_cases

        path, words = parity_cases._arpa("p_small4.arpa", 300, 4, [0, 3000, 3000, 2000], 3)
        lm, lmw = ("arpa", path, words[:N]), 0.8
    if lexicon:
        sp = synth.lexicon(60 if mode != "max" else 200, N, 2, 4, seed=7, exclude=(0, N - 1))
        spec = spec_lexicon(N, 20, N, sp, 1e9 if mode == "max" else 30.0, word_score=0.3, log_add=log_add,
                            lm=lm, lm_weight=lmw, is_lm_token=mode.startswith("tokenlm"))
    else:
        spec = spec_lexfree(N, 12, N, 1e9 if mode == "max" else 30.0, log_add=log_add, lm=lm, lm_weight=lmw)
    ba, bg = Built(A, spec), Built(G, spec)
    for O, b in ((A, ba), (G, bg)):
        O.decode_begin(b.dec)
    step = 0
    for c in range(0, T, chunk):
        outs = []
        for O, b in ((A, ba), (G, bg)):
            O.decode_step(b.dec, em[c:c + chunk])
            best = O.best(b.dec, look_back, T + 2)
            nh, nf = O.n_hypothesis(b.dec), O.n_frames_in_buffer(b.dec)
            if step % prune_every == 0:
                O.prune(b.dec, look_back)
            outs.append((best, nh, nf, O.n_frames_in_buffer(b.dec)))
        step += 1
        (b0, *r0), (b1, *r1) = outs
        assert r0 == r1, (c, r0, r1)
        np.testing.assert_array_equal(b0["tokens"], b1["tokens"])
        np.testing.assert_array_equal(b0["words"], b1["words"])
        np.testing.assert_allclose(b0["scores"], b1["scores"], rtol=0, atol=tol)
    fin = []
    for O, b in ((A, ba), (G, bg)):
        O.decode_end(b.dec)
        fin.append(O.all_final(b.dec, 64, T + 2))
    assert fin[0]["n"] == fin[1]["n"] and fin[0]["n"] > 0
    np.testing.assert_array_equal(fin[0]["lens"], fin[1]["lens"])
    np.testing.assert_array_equal(fin[0]["tokens"], fin[1]["tokens"])
    np.testing.assert_array_equal(fin[0]["words"], fin[1]["words"])
    np.testing.assert_allclose(fin[0]["scores"], fin[1]["scores"], rtol=0, atol=tol)
    ba.close(), bg.close()


CASES = [(False, 15, 5, 1), (True, 15, 5, 1), (False, 7, 0, 2), (True, 30, 3, 1), (False, 90, 2, 1)]


@pytest.mark.parametrize("lexicon,chunk,look_back,prune_every", CASES)
def test_streaming_logic_harness(lexicon, chunk, look_back, prune_every):
    from flt_backend import FltBackend

    run_streaming(po.Oracle("ora"), FltBackend("model"), lexicon, chunk, look_back, prune_every, 1e-9)


@pytest.mark.parametrize("lexicon", [False, True])
@pytest.mark.parametrize("mode", ["logadd", "tokenlm", "tokenlm_logadd"])
def test_streaming_full_expansion_modes(lexicon, mode):
    from flt_backend import FltBackend

    run_streaming(po.Oracle("ora"), FltBackend("model"), lexicon, 15, 4, 2, 1e-9, mode)


@pytest.mark.parametrize("lexicon,mode", [(False, "logadd"), (True, "logadd"), (True, "max")])
def test_streaming_overflow_retry(monkeypatch, lexicon, mode):
    """A chunk whose frames overflow the candidate capacity is repeated with a larger one from the beam
    saved before it (FLT_TEST_CAP starts the capacity tiny); results equal the oracle's."""
    from flt_backend import FltBackend

    monkeypatch.setenv("FLT_TEST_CAP", "4")
    run_streaming(po.Oracle("ora"), FltBackend("model"), lexicon, 15, 4, 2, 1e-9, mode)


def test_batch_overflow_retry(monkeypatch):
    from flt_backend import FltBackend
    import parity_cases
    from cases import assert_same_nbest, has_ties

    monkeypatch.setenv("FLT_TEST_CAP", "4")
    A, M = po.Oracle("ora"), FltBackend("model")
    for name, spec, em, exact in parity_cases.widened_cases()[:2] + parity_cases.widened_cases()[8:10]:
        ba, bm = Built(A, spec), Built(M, spec)
        got = bm.O.decode_batch(bm.dec, em, spec["opt"].beamSize)
        for b, e in enumerate(em):
            ra = ba.decode(e)
            if has_ties(ra) or A.tie_events(ba.dec):
                continue
            assert_same_nbest(ra, got[b], 1e-9, what=f"{name} utt {b}")
        ba.close(), bm.close()


def test_unpruned_chunks_equal_offline_decode():
    """Without prune(), chunked decodeStep + decodeEnd gives exactly decode()'s n-best."""
    from flt_backend import FltBackend

    M = FltBackend("model")
    N, T = 40, 70
    em = synth.emissions(1, T, N, seed=5, sigma=2.0)[0]
    spec = spec_lexfree(N, 16, N, 1e9)
    b = Built(M, spec)
    off = b.decode(em)
    M.decode_begin(b.dec)
    for c in range(0, T, 13):
        M.decode_step(b.dec, em[c:c + 13])
    M.decode_end(b.dec)
    fin = M.all_final(b.dec, 16, T + 2)
    assert fin["n"] == off["n"]
    np.testing.assert_array_equal(fin["tokens"], off["tokens"])
    np.testing.assert_array_equal(fin["scores"], off["scores"])
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lexicon,chunk,look_back,prune_every", CASES)
def test_streaming_cuda(lexicon, chunk, look_back, prune_every):
    import torch

    assert torch.cuda.is_available()
    from flt_backend import FltBackend

    run_streaming(po.Oracle("ora"), FltBackend("cuda"), lexicon, chunk, look_back, prune_every, 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("lexicon", [False, True])
@pytest.mark.parametrize("mode", ["logadd", "tokenlm"])
def test_streaming_full_expansion_modes_cuda(lexicon, mode):
    from flt_backend import FltBackend

    run_streaming(po.Oracle("ora"), FltBackend("cuda"), lexicon, 15, 4, 2, 1e-4, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("lexicon,mode", [(False, "logadd"), (True, "max")])
def test_streaming_overflow_retry_cuda(monkeypatch, lexicon, mode):
    from flt_backend import FltBackend

    monkeypatch.setenv("FLT_TEST_CAP", "4")
    run_streaming(po.Oracle("ora"), FltBackend("cuda"), lexicon, 15, 4, 2, 1e-4, mode)

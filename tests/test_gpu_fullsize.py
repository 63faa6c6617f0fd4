"""Parity at BASELINE.json's FULL sizes (configs[1] and configs[2]: B=256, T=1000, N=10000), where the CPU
oracle cannot run in test time (0.1 utt/s at beamSizeToken = N). Size-independent properties of the
decoder's own output, all checked bit-exactly:

* score reconstruction: with ZeroLM, lmWeight = 0 and no sil / word bonus, a hypothesis' `score` and
  `emittingModelScore` are the sequential FP64 sum of the FP32 emissions along its own token path
  (LexiconFreeDecoder.cpp:64, LexiconDecoder.cpp:74: `score = prevHyp.score + emissions[t*N+n]`), and
  `lmScore` is 0 — recomputed here from the returned tokens and the input, independently of the kernels;
* the n-best is sorted by score, full (beam hypotheses) and made of distinct token paths;
* every path is a legal one: tokens[0] = tokens[T+1] = sil, and (lexicon) the words spelled along it are
  lexicon entries whose spelling is the collapsed token string between word ends;
* idempotence (decoding the same device buffer twice gives identical bits) and batch independence (a
  slice of the batch decoded on its own gives the slice of the batch result);
* the first frames of a few utterances against the oracle (prefix property: the n-best of a prefix is
  what the oracle computes for that prefix)."""
import numpy as np
import pytest

from cases import Built, assert_same_nbest, has_ties, spec_lexfree, spec_lexicon
from oracle import pyoracle as po
from text_b200 import synth

pytestmark = pytest.mark.gpu
B, T, N = 256, 1000, 10000


@pytest.fixture(scope="module")
def em():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(4321)
    e = torch.empty((B, T, N), dtype=torch.float32, device=dev)
    for b0 in range(0, B, 32):
        z = torch.randn((32, T, N), generator=gen, device=dev, dtype=torch.float32)
        e[b0:b0 + 32] = torch.log_softmax(z, dim=-1)
    del z
    torch.cuda.synchronize()
    yield e
    del e
    torch.cuda.empty_cache()


def decode(G, dec, e, b0, nb, nbest):
    G.api.set_nbest(dec, nbest)
    G.api.decode_batch_ptr(dec, e[b0:b0 + nb].data_ptr(), nb, T, N)
    return G.api.nbest(dec, nb, T, nbest)


def path_score(e, b, toks):
    """sequential FP64 sum of the FP32 emissions along frames 1..T of a returned token row"""
    import torch

    idx = torch.as_tensor(toks[1:T + 1].astype(np.int64), device=e.device)
    vals = e[b].gather(1, idx[:, None])[:, 0].double().cpu().numpy()
    return np.cumsum(vals)[-1]  # cumsum accumulates left to right, like the decoder


def check_common(e, r, K, sample, full=True):
    """`full`: the beam must come back full (lexicon-free). The lexicon decoder's decodeEnd keeps only the
    hypotheses that sit at the Trie root when any does (LexiconDecoder.cpp:233-246), so its count varies."""
    nb = r["tokens"].shape[1]
    cnt = r["counts"]
    assert (cnt == K).all() if full else ((cnt >= 1) & (cnt <= K)).all(), "number of final hypotheses"
    for b in range(len(cnt)):
        n = min(int(cnt[b]), nb)
        assert (np.diff(r["scores"][b, :n, 0]) <= 0).all(), "n-best not sorted by score"
        assert (r["tokens"][b, :n, 0] == 0).all() and (r["tokens"][b, :n, T + 1] == 0).all()
    for b in sample:
        n = min(int(cnt[b]), nb)
        rows = {r["tokens"][b, k].tobytes() + r["words"][b, k].tobytes() for k in range(n)}
        assert len(rows) == n, "duplicate hypotheses in the n-best"
        for k in sorted({0, min(1, n - 1), n - 1}):
            want = path_score(e, b, r["tokens"][b, k])
            assert r["scores"][b, k, 0] == want, f"utt {b} rank {k}: score is not the sum of its path"
            assert r["scores"][b, k, 1] == want, f"utt {b} rank {k}: emittingModelScore"


def masked(r):
    """copies with the rows beyond each utterance's hypothesis count cleared (they are not written)"""
    out = {k: r[k].copy() for k in ("tokens", "words", "scores", "counts")}
    for b, n in enumerate(out["counts"]):
        for k in ("tokens", "words", "scores"):
            out[k][b, int(n):] = 0
    return out


def check_repeatable(G, dec, e, r, nbest):
    r = masked(r)
    again = masked(decode(G, dec, e, 0, B, nbest))
    for key in ("tokens", "words", "scores", "counts"):
        assert np.array_equal(r[key], again[key]), f"{key} differ between two decodes of the same buffer"
    part = masked(decode(G, dec, e, 40, 24, nbest))
    for key in ("tokens", "words", "scores", "counts"):
        assert np.array_equal(part[key], r[key][40:64]), f"{key}: batch slice decoded alone differs"


def check_prefix_vs_oracle(G, spec, e, frames, utts):
    A = po.Oracle("ora")
    K = spec["opt"].beamSize
    sub = e[utts, :frames].contiguous().cpu().numpy()
    ba, bg = Built(A, spec), Built(G, spec)
    got = bg.O.decode_batch(bg.dec, sub, K)
    checked = 0
    for i in range(len(utts)):
        ra = ba.decode(sub[i])
        if has_ties(ra) or A.tie_events(ba.dec):
            continue
        assert_same_nbest(ra, got[i], 1e-4, what=f"prefix of utt {utts[i]}")
        assert np.array_equal(ra["scores"], got[i]["scores"])
        checked += 1
    ba.close(), bg.close()
    assert checked


def test_lexfree_cfg2_full_size(em):
    from flt_backend import FltBackend

    G = FltBackend("cuda")
    K = 50
    spec = spec_lexfree(N, K, N, 1e9, sil=0, blank=N - 1)
    b = Built(G, spec)
    r = decode(G, b.dec, em, 0, B, 8)
    check_common(em, r, K, sample=range(0, B, 16))
    assert (r["words"] == -1).all()
    assert (r["scores"][:, :, 2] == 0).all(), "lmScore must be 0 with ZeroLM"
    check_repeatable(G, b.dec, em, r, 8)
    b.close()
    check_prefix_vs_oracle(G, spec, em, 10, [0, 100, 255])


def test_lexicon_cfg3_full_size(em):
    from flt_backend import FltBackend

    G = FltBackend("cuda")
    K, W = 100, 200000
    sp = synth.lexicon(W, N, 2, 5, seed=7, exclude=(0, N - 1))
    spec = spec_lexicon(N, K, N, sp, 1e9, sil=0, blank=N - 1, unk=W)
    b = Built(G, spec)
    r = decode(G, b.dec, em, 0, B, 8)
    check_common(em, r, K, sample=range(0, B, 16), full=False)
    # every word end spells a lexicon entry: collapse the CTC alignment since the previous word end
    spell = {}
    for w, s in enumerate(sp):
        spell.setdefault(tuple(int(x) for x in s), set()).add(w)
    blank = N - 1
    for bb in range(0, B, 32):
        toks, words = r["tokens"][bb, 0], r["words"][bb, 0]
        cur, prev = [], None
        for t in range(1, T + 1):
            n = int(toks[t])
            if n == blank or n == 0:
                prev = n if n == blank else None
                if n == 0:
                    assert not cur or words[t] < 0
                continue
            if n != prev:
                cur.append(n)
            prev = n
            if words[t] >= 0:
                assert int(words[t]) in spell.get(tuple(cur), ()), f"utt {bb} frame {t}: word {words[t]} vs spelling {cur}"
                cur, prev = [], None
    check_repeatable(G, b.dec, em, r, 8)
    b.close()
    check_prefix_vs_oracle(G, spec, em, 8, [3, 77])

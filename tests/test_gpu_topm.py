"""Kernel-level test of the token-beam select (K1) through flt_topm_rows: for every row the list
must equal the reference's partial_sort result (decoder/LexiconFreeDecoder.cpp:39-51) under the
deterministic tie rule (value descending, then token ascending) — including adversarial rows that
defeat the chunk-maximum bound (sorted, constant, heavy ties, -inf) and shapes that take the generic
path (N % 4 != 0, misaligned base pointer, M > 256)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def expected(rows, M):
    N = rows.shape[1]
    idx = np.lexsort((np.arange(N)[None, :].repeat(rows.shape[0], 0), -rows.astype(np.float64)), axis=1)
    tok = idx[:, :M]
    return tok, np.take_along_axis(rows, tok, axis=1)


def make_rows(kind, R, N, rng):
    if kind == "gauss":
        return rng.standard_normal((R, N)).astype(np.float32)
    if kind == "logsoftmax":
        z = rng.standard_normal((R, N)).astype(np.float32)
        return (z - np.log(np.exp(z).sum(1, keepdims=True))).astype(np.float32)
    if kind == "ascending":
        return np.tile(np.arange(N, dtype=np.float32), (R, 1))
    if kind == "descending":
        return np.tile(-np.arange(N, dtype=np.float32), (R, 1))
    if kind == "constant":
        return np.full((R, N), -3.25, np.float32)
    if kind == "few_values":
        return rng.integers(0, 4, size=(R, N)).astype(np.float32)
    if kind == "neg_inf":
        x = rng.standard_normal((R, N)).astype(np.float32)
        x[:, ::3] = -np.inf
        return x
    raise ValueError(kind)


CASES = [
    # N, M, offset (floats)
    (10000, 53, 0), (10000, 1, 0), (10000, 256, 0), (10000, 300, 0), (10240, 53, 0), (10244, 53, 0),
    (9999, 53, 0), (10000, 53, 1), (5000, 103, 0), (29, 29, 0), (29, 5, 0), (64, 64, 0), (4096, 53, 0),
    (4100, 53, 0),
]


@pytest.mark.parametrize("kind", ["gauss", "logsoftmax", "ascending", "descending", "constant",
                                  "few_values", "neg_inf"])
@pytest.mark.parametrize("N,M,offset", CASES)
def test_topm_rows(kind, N, M, offset):
    import torch

    from text_b200 import capi

    api = capi.Api()
    rng = np.random.default_rng(N * 7 + M)
    R = 37
    rows = make_rows(kind, R, N, rng)
    flat = torch.empty(R * N + offset + 8, dtype=torch.float32, device="cuda")
    view = flat[offset:offset + R * N].view(R, N)
    view.copy_(torch.from_numpy(rows))
    tok = torch.full((R, M), -7, dtype=torch.int32, device="cuda")
    val = torch.zeros((R, M), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    api.topm_rows(view.data_ptr(), R, N, M, tok.data_ptr(), val.data_ptr(), None)
    torch.cuda.synchronize()
    etok, eval_ = expected(rows, M)
    np.testing.assert_array_equal(tok.cpu().numpy(), etok)
    np.testing.assert_array_equal(val.cpu().numpy(), eval_)


@pytest.mark.parametrize("kind,N,M,R", [("logsoftmax", 2048, 53, 5000), ("gauss", 1000, 205, 3000),
                                        ("few_values", 1024, 40, 2500), ("neg_inf", 4096, 100, 2000)])
def test_topm_many_rows(kind, N, M, R):
    """More rows than resident CTAs: every CTA of the streaming kernel goes through many rows with its stage,
    its mbarrier phase and the running guess of the select bound carried from row to row."""
    import torch

    from text_b200 import capi

    api = capi.Api()
    rng = np.random.default_rng(N + M + R)
    rows = make_rows(kind, R, N, rng)
    # the level of the rows drifts and jumps, so that guesses miss and the exact modes engage
    rows += (np.sin(np.arange(R) / 7.0) * 3.0 + (np.arange(R) % 97 == 0) * 20.0).astype(np.float32)[:, None]
    dev = torch.from_numpy(rows).cuda()
    tok = torch.full((R, M), -7, dtype=torch.int32, device="cuda")
    val = torch.zeros((R, M), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    api.topm_rows(dev.data_ptr(), R, N, M, tok.data_ptr(), val.data_ptr(), None)
    torch.cuda.synchronize()
    etok, eval_ = expected(rows, M)
    np.testing.assert_array_equal(tok.cpu().numpy(), etok)
    np.testing.assert_array_equal(val.cpu().numpy(), eval_)


@pytest.mark.parametrize("N,M,R,spread", [(10000, 105, 1500, 6.0), (10000, 205, 1500, 6.0), (10000, 205, 900, 0.5),
                                          (10000, 505, 900, 0.0), (10000, 505, 900, 6.0), (4096, 600, 700, 3.0),
                                          (1000, 53, 600, 6.0), (10000, 1000, 300, 0.0)])
def test_topm_biased_and_long_lists(N, M, R, spread):
    """The lexicon decoder's select: rank by e[n] + bias[n] (bias = lmWeight * smeared LM score of root child n,
    spread wider than the emissions when an n-gram LM is smeared into the Trie; -inf = not a word start), and
    the long lists of wide beams (beam 200 / 500: M = 205 / 505; M = 1000 takes the generic kernel). Values
    returned are the raw emissions of the selected tokens."""
    import torch

    from text_b200 import capi

    api = capi.Api()
    rng = np.random.default_rng(N + M + R)
    rows = make_rows("logsoftmax", R, N, rng)
    rows += (np.sin(np.arange(R) / 9.0) * 2.0 + (np.arange(R) % 131 == 0) * 15.0).astype(np.float32)[:, None]
    bias = None
    if spread > 0:
        bias = (-rng.random(N) * spread).astype(np.float32)
        bias[rng.random(N) < 0.03] = -np.inf
    dev = torch.from_numpy(rows).cuda()
    dbias = torch.from_numpy(bias).cuda() if bias is not None else None
    tok = torch.full((R, M), -7, dtype=torch.int32, device="cuda")
    val = torch.zeros((R, M), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    api.topm_rows(dev.data_ptr(), R, N, M, tok.data_ptr(), val.data_ptr(), None,
                  dbias.data_ptr() if dbias is not None else None)
    torch.cuda.synchronize()
    key = rows if bias is None else (rows + bias[None, :]).astype(np.float32)
    etok, _ = expected(key, M)
    np.testing.assert_array_equal(tok.cpu().numpy(), etok)
    np.testing.assert_array_equal(val.cpu().numpy(), np.take_along_axis(rows, etok, axis=1))


def test_topm_biased_few_eligible():
    """fewer word-starting tokens than the list is long: the list holds them all, then -1 / 0 padding"""
    import torch

    from text_b200 import capi

    api = capi.Api()
    N, M, R = 2048, 105, 400
    rng = np.random.default_rng(5)
    rows = make_rows("logsoftmax", R, N, rng)
    bias = np.full(N, -np.inf, np.float32)
    ok = rng.choice(N, size=40, replace=False)
    bias[ok] = -rng.random(40).astype(np.float32)
    dev, dbias = torch.from_numpy(rows).cuda(), torch.from_numpy(bias).cuda()
    tok = torch.full((R, M), -7, dtype=torch.int32, device="cuda")
    val = torch.zeros((R, M), dtype=torch.float32, device="cuda")
    api.topm_rows(dev.data_ptr(), R, N, M, tok.data_ptr(), val.data_ptr(), None, dbias.data_ptr())
    torch.cuda.synchronize()
    key = (rows + bias[None, :]).astype(np.float32)
    etok, _ = expected(key, 40)
    got = tok.cpu().numpy()
    np.testing.assert_array_equal(got[:, :40], etok)
    assert (got[:, 40:] == -1).all() and (val.cpu().numpy()[:, 40:] == 0).all()
    np.testing.assert_array_equal(val.cpu().numpy()[:, :40], np.take_along_axis(rows, etok, axis=1))

"""Seeded synthetic inputs for the decode hot path: emissions, unique-spelling
lexicons and well-formed back-off ARPA files. numpy only; used by tests/ and bench.py.
"""
import os

import numpy as np


def emissions(B, T, N, seed=1234, sigma=1.0):
    """emis[b,t,:] = log_softmax(sigma * z), z ~ N(0,1), fp32, row-major [B,T,N]."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((B, T, N), dtype=np.float32) * np.float32(sigma)
    m = z.max(axis=-1, keepdims=True)
    lse = m + np.log(np.exp(z - m).sum(axis=-1, keepdims=True, dtype=np.float32))
    return (z - lse).astype(np.float32)


def emissions_exact(B, T, N, seed=1, sigma=1.0, shift=-9.71):
    """Bit-reproducible emissions for the committed full-size golden vectors: integer hashing plus
    IEEE add / multiply / cast only (no exp / log, whose last bit may depend on the libm and CPU), so
    every machine produces the same fp32 bits. Each value is `shift + sigma * z` with z the
    standardised sum of four uniform 16-bit fields of a splitmix64 hash of (seed, b, t, n) — bell
    shaped like the N(0,1) logits of `emissions`, at the level of their log-softmax (lse ~ 9.7 at
    N = 10000). Not normalised; the decoders do not need that."""
    out = np.empty((B, T, N), dtype=np.float32)
    idx = np.arange(T * N, dtype=np.uint64)

    def mix(x):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))

    m = np.uint64(0xFFFF)
    for b in range(B):
        x = mix(idx + np.uint64((int(seed) * 1000003 + b) * 0x9E3779B97F4A7C15 % (1 << 64)))
        s = ((x & m) + ((x >> np.uint64(16)) & m) + ((x >> np.uint64(32)) & m) + (x >> np.uint64(48))).astype(np.int64)
        # a 32-bit dither below the 16-bit grid, so that equal values (ties at the token-beam cut, equal
        # path sums) are as rare as with continuous logits
        d = (mix(x + np.uint64(0x632BE59BD9B4E019)) >> np.uint64(32)).astype(np.float64) * (1.0 / 4294967296.0)
        z = ((s - 131070).astype(np.float64) + d) * (1.0 / 37837.227)
        out[b] = (z * float(sigma) + float(shift)).astype(np.float32).reshape(T, N)
    return out


def lexicon(W, N, min_len=2, max_len=5, seed=7, exclude=()):
    """W distinct spellings (so every terminal trie node has exactly one label: no homophone
    ties), tokens uniform in [0,N) minus `exclude` (sil / blank)."""
    rng = np.random.default_rng(seed)
    allowed = np.array([t for t in range(N) if t not in set(exclude)], dtype=np.int32)
    seen, out = set(), []
    while len(out) < W:
        need = W - len(out)
        lens = rng.integers(min_len, max_len + 1, size=need + 16)
        toks = allowed[rng.integers(0, len(allowed), size=(need + 16, max_len))]
        for L, row in zip(lens, toks):
            sp = tuple(int(x) for x in row[:L])
            if sp not in seen:
                seen.add(sp)
                out.append(np.array(sp, dtype=np.int32))
                if len(out) == W:
                    break
    return out


def word_names(W):
    return [f"w{i}" for i in range(W)]


def write_arpa(path, W, order=4, counts=None, seed=11, with_unk=True):
    """Well-formed synthetic back-off LM over words w0..w{W-1} (+ <unk>, <s>, </s>): every
    n-gram's (n-1)-word prefix is itself listed, log10 probs in [-6,-0.1], back-offs in [-1,0]
    (none on the highest order). Returns the list of n-gram counts per order."""
    rng = np.random.default_rng(seed)
    counts = list(counts or [])
    vocab = (["<unk>"] if with_unk else []) + ["<s>", "</s>"] + word_names(W)
    V = len(vocab)
    bos = vocab.index("<s>")
    eos = vocab.index("</s>")
    levels = []  # levels[k] = int array [n_k, k+1] of vocab ids
    uni = np.arange(V, dtype=np.int64)[:, None]
    levels.append(uni)
    for k in range(1, order):
        want = counts[k] if k < len(counts) else max(1, len(levels[-1]) // 2)
        prev = levels[-1]
        # contexts must not end in </s>; extend a random existing (k)-gram by one word
        ok = prev[prev[:, -1] != eos]
        pick = ok[rng.integers(0, len(ok), size=int(want * 1.15) + 8)]
        nxt = rng.integers(0, V, size=len(pick))
        nxt[nxt == bos] = eos
        cand = np.concatenate([pick, nxt[:, None]], axis=1)
        # distinct rows in lexicographic order (= np.unique(cand, axis=0), several times faster)
        order_ = np.lexsort(cand.T[::-1])
        cand = cand[order_]
        keep = np.ones(len(cand), dtype=bool)
        keep[1:] = (cand[1:] != cand[:-1]).any(axis=1)
        cand = cand[keep]
        rng.shuffle(cand)
        levels.append(cand[:want])
    varr = np.array(vocab, dtype=object)
    with open(path, "w") as f:
        f.write("\\data\\\n")
        for k, lv in enumerate(levels):
            f.write(f"ngram {k + 1}={len(lv)}\n")
        for k, lv in enumerate(levels):
            f.write(f"\n\\{k + 1}-grams:\n")
            probs = -(rng.random(len(lv)) * 5.9 + 0.1)
            bows = -rng.random(len(lv))
            last = k == order - 1
            # vectorised formatting (5 M lines in a Python loop took most of the benchmark's set-up time)
            ws = varr[lv[:, 0]]
            for col in range(1, lv.shape[1]):
                ws = ws + " " + varr[lv[:, col]]
            ps = np.char.mod("%.6f", probs).astype(object)
            bs = np.char.mod("%.6f", bows).astype(object)
            with_bow = ps + "\t" + ws + "\t" + bs + "\n"
            without = ps + "\t" + ws + "\n"
            ends_eos = lv[:, -1] == eos
            lines = without if last else np.where(ends_eos, without, with_bow)
            if k == 0:
                i_bos = int(np.nonzero(lv[:, 0] == bos)[0][0])
                lines = lines.copy()
                lines[i_bos] = "-99\t" + ws[i_bos] + "\t" + bs[i_bos] + "\n"
            f.write("".join(lines.tolist()))
        f.write("\n\\end\\\n")
    return [len(lv) for lv in levels]


def cache_dir():
    import tempfile

    d = os.environ.get("TEXT_B200_CACHE", os.path.join(tempfile.gettempdir(), "text_b200_cache"))
    os.makedirs(d, exist_ok=True)
    return d

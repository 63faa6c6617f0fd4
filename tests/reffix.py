"""Loader for the reference's own DecoderTest fixture (read in place from /root/reference, which
only exists in the build container; callers skip when it is absent). Mirrors the setup code of
flashlight/lib/text/test/decoder/DecoderTest.cpp:57-146 (loadWords / createWordDict / tkn2Idx with
one replabel), with word ids assigned in file order instead of unordered_map order.
"""
import os
import struct

import numpy as np

DATA = "/root/reference/flashlight/lib/text/test/decoder/data"


def present():
    return os.path.isdir(DATA)


def load():
    with open(os.path.join(DATA, "TN.bin"), "rb") as f:
        T, N = struct.unpack("ii", f.read(8))
    emissions = np.fromfile(os.path.join(DATA, "emission.bin"), np.float32).reshape(T, N)
    transitions = np.fromfile(os.path.join(DATA, "transition.bin"), np.float32)
    tokens = [l.strip() for l in open(os.path.join(DATA, "letters.lst")) if l.strip()]
    tokens.append("<1>")  # replabel emulation, DecoderTest.cpp:95-96
    tok2idx = {t: i for i, t in enumerate(tokens)}
    lexicon = {}  # word -> list of spellings (insertion ordered)
    for line in open(os.path.join(DATA, "words.lst")):
        f = line.split()
        if len(f) < 2:
            continue
        lexicon.setdefault(f[0], []).append(f[1:])
    lexicon.setdefault("<unk>", [])
    words = list(lexicon.keys())
    return dict(T=T, N=N, emissions=emissions, transitions=transitions, tokens=tokens,
                tok2idx=tok2idx, lexicon=lexicon, words=words,
                word2idx={w: i for i, w in enumerate(words)},
                arpa=os.path.join(DATA, "lm.arpa"))


def pack_replabels(idx, tok2idx, max_reps):
    """dictionary/Utils.cpp:90-121"""
    if not idx or max_reps <= 0:
        return list(idx)
    rep = {i: tok2idx[f"<{i}>"] for i in range(1, max_reps + 1)}
    out, prev, n = [], -1, 0
    for t in idx:
        if t == prev and n < max_reps:
            n += 1
        else:
            if n > 0:
                out.append(rep[n])
                n = 0
            out.append(t)
            prev = t
    if n > 0:
        out.append(rep[n])
    return out


def tkn2idx(spelling, tok2idx, max_reps):
    return pack_replabels([tok2idx[t] for t in spelling], tok2idx, max_reps)

#!/usr/bin/env python
"""Generate tests/golden/fullsize_nbest.npz: n-best lists at BASELINE.json's shapes (N = 10000; see
tests/fullsize_cases.py) from the UNMODIFIED reference (oracle/_ref/libflref.so, `bst = beam` rows, full
length) and from the oracle port (`bst = N` rows, a prefix: the reference itself allocates ~66 MB of
LMState per frame with token pruning off). Every `ref` row is also decoded by the port and must come out
bit-equal (unless the port's tie detector fires), which pins the port at these sizes too.

Stored per utterance: all final scores [n,3] fp64, the first NB token / word rows, a CRC32 of ALL n rows,
and the port's tie-event count (utterances with ties are excluded by the tests: their outcome is
libstdc++-internal in the reference). Inputs are not stored — seeds + a CRC of the
emissions and of the ARPA file are.

Run in the build container (needs oracle/_ref, i.e. /root/reference):  python tests/golden/make_fullsize.py
"""
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fullsize_cases as fc  # noqa: E402
from cases import Built, assert_same_nbest, has_ties  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

NB = 8  # token / word rows stored per utterance


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def file_crc(path):
    c = 0
    with open(path, "rb") as f:
        while True:
            blk = f.read(1 << 24)
            if not blk:
                break
            c = zlib.crc32(blk, c)
    return np.uint32(c)


def main():
    only = set(sys.argv[1:])
    po.build("ref")
    po.build("ora")
    R, A = po.Oracle("ref"), po.Oracle("ora")
    path = os.path.join(HERE, "fullsize_nbest.npz")
    out = dict(np.load(path)) if (only and os.path.exists(path)) else {}
    for row in fc.ROWS:
        name, kind, beam, bst, T, B, thr, oracle, seed = row
        if only and name not in only:
            continue
        t0 = time.time()
        spec, em = fc.build(row)
        out[f"{name}/crc"] = crc(em)
        if kind == "lexicon_lm":
            out[f"{name}/arpa_crc"] = file_crc(fc.arpa_4gram())
        K = spec["opt"].beamSize

        def one(u):
            ba = Built(A, spec)
            ra = ba.decode(em[u], K)
            ties = A.tie_events(ba.dec)
            ba.close()
            r = ra
            if oracle == "ref":
                br = Built(R, spec)
                r = br.decode(em[u], K)
                br.close()
                if ties == 0 and not has_ties(r):
                    assert_same_nbest(r, ra, 0.0, what=f"{name}/{u} reference vs port")
                    assert np.array_equal(r["scores"], ra["scores"])
            return u, r, ties

        with ThreadPoolExecutor(max_workers=min(B, os.cpu_count() or 1)) as ex:
            for u, r, ties in ex.map(one, range(B)):
                out[f"{name}/{u}/ties"] = np.int64(ties + (1 if has_ties(r) else 0))
                out[f"{name}/{u}/scores"] = r["scores"]
                out[f"{name}/{u}/tokens"] = r["tokens"][:NB].astype(np.int32)
                out[f"{name}/{u}/words"] = r["words"][:NB].astype(np.int32)
                out[f"{name}/{u}/rows_crc"] = np.array([crc(r["tokens"].astype(np.int32)),
                                                        crc(r["words"].astype(np.int32))], np.uint32)
                print(f"{name} utt {u}: n={r['n']} ties={int(out[f'{name}/{u}/ties'])} best={r['scores'][0, 0]:.6f}",
                      flush=True)
        print(f"{name}: {time.time() - t0:.1f} s", flush=True)
        np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

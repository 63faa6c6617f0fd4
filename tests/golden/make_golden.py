#!/usr/bin/env python
"""Generate tests/golden/ref_nbest.npz: n-best lists produced by the UNMODIFIED reference decoder
(oracle/_ref/libflref.so = /root/reference sources compiled in place, see oracle/Makefile) on the
seeded parity cases of tests/parity_cases.py, plus the reference's own DecoderTest emission fixture
(flashlight/lib/text/test/decoder/data/{TN,emission,transition}.bin; the 27 KB emission matrix is
embedded so the case can run where /root/reference is absent).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The committed .npz is what `tests/test_golden.py` checks the oracle restatement (CPU) and the CUDA
path (GPU) against; /root/reference is never read at test time.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity_cases  # noqa: E402
import reffix  # noqa: E402
from cases import Built, assert_same_nbest, has_ties, spec_lexfree  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def main():
    po.build("ref")
    R = po.Oracle("ref")
    po.build("ora")
    A = po.Oracle("ora")  # only for its tie detector (and a generation-time cross-check)
    out = {}
    names = []
    for name, spec, em in parity_cases.lexfree_cases() + parity_cases.lexicon_cases():
        K = spec["opt"].beamSize
        b = Built(R, spec)
        ba = Built(A, spec)
        out[f"{name}/crc"] = crc(em)
        for u, e in enumerate(em):
            r = b.decode(e, K)
            ra = ba.decode(e, K)
            ties = A.tie_events(ba.dec)
            out[f"{name}/{u}/ties"] = np.int64(ties)
            if ties == 0 and not has_ties(r):
                assert_same_nbest(r, ra, 0.0, what=f"{name}/{u} ref vs restatement")
            out[f"{name}/{u}/scores"] = r["scores"]
            out[f"{name}/{u}/tokens"] = r["tokens"].astype(np.int32)
            out[f"{name}/{u}/words"] = r["words"].astype(np.int32)
        b.close()
        ba.close()
        names.append(name)
    # the reference's own emission fixture, lexicon-free (SURVEY.md §8c vectors)
    fx = reffix.load()
    out["fixture/emissions"] = fx["emissions"]
    out["fixture/transitions"] = fx["transitions"]
    for tag, kw in (("ctc", {}), ("ctc_bst5_thr25", dict(bst=5, thr=25.0)),
                    ("asg", dict(criterion=po.ASG, transitions=fx["transitions"]))):
        spec = spec_lexfree(29, 10, kw.get("bst", 29), kw.get("thr", 1e9), sil=0,
                            blank=28 if "criterion" not in kw else -1,
                            criterion=kw.get("criterion", po.CTC), transitions=kw.get("transitions"))
        b = Built(R, spec)
        r = b.decode(fx["emissions"], 10)
        out[f"fixture/{tag}/scores"] = r["scores"]
        out[f"fixture/{tag}/tokens"] = r["tokens"].astype(np.int32)
        out[f"fixture/{tag}/words"] = r["words"].astype(np.int32)
        b.close()
    out["names"] = np.array(names)
    path = os.path.join(HERE, "ref_nbest.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(names)} cases")


if __name__ == "__main__":
    main()

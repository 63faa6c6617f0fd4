"""The C-ABI library loads (without a GPU) and exports every entry point include/flt_decoder.h
declares; the ctypes table in text_b200/capi.py covers the same set. No compute calls here."""
import ctypes
import os
import re

from text_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = open(os.path.join(ROOT, "include", "flt_decoder.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(flt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    assert sorted(capi.SYMBOLS) == declared()


def test_errors_are_codes_not_exceptions():
    api = capi.Api()
    t = api.trie_create(10, 0)
    try:
        api.trie_insert(t, [11], 0, 0.0)
        assert False, "expected FLT_ERR_OUT_OF_RANGE"
    except capi.FltError as e:
        assert e.code == capi.ERR_OUT_OF_RANGE and "Invalid letter index" in e.msg
    api.trie_destroy(t)

"""`flashlight.lib.text.decoder` as served by the B200 decode path: the same import path and names
as the reference package (bindings/python/flashlight/lib/text/decoder/__init__.py:12-32) for the
CTC/ASG decoders. Put `text_b200/compat` on sys.path to use it as a drop-in."""
from text_b200.pyext import load as _load

_m = _load()
CriterionType = _m.CriterionType
DecodeResult = _m.DecodeResult
LexiconDecoder = _m.LexiconDecoder
LexiconDecoderOptions = _m.LexiconDecoderOptions
LexiconFreeDecoder = _m.LexiconFreeDecoder
LexiconFreeDecoderOptions = _m.LexiconFreeDecoderOptions
LM = _m.LM
LMState = _m.LMState
SmearingMode = _m.SmearingMode
Trie = _m.Trie
TrieNode = _m.TrieNode
ZeroLM = _m.ZeroLM
KenLM = _m.KenLM  # ARPA files only (the reference moved it to decoder.kenlm; both paths work)

"""`flashlight.lib.text.dictionary` (bindings/python/flashlight/lib/text/dictionary.py): the host
setup path that feeds the decoders — Dictionary, lexicon loading, replabels."""
from text_b200.pyext import load as _load

_m = _load()
Dictionary = _m.Dictionary
create_word_dict = _m.create_word_dict
load_words = _m.load_words
pack_replabels = _m.pack_replabels
unpack_replabels = _m.unpack_replabels
tkn_to_idx = _m.tkn_to_idx
split_wrd = _m.split_wrd
build_trie = _m.build_trie

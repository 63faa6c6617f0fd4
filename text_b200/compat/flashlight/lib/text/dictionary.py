"""`flashlight.lib.text.dictionary.Dictionary` (subset needed to build a KenLM vocabulary map)."""
from text_b200.pyext import load as _load

Dictionary = _load().Dictionary

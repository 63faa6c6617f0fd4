"""`flashlight.lib.text.decoder.kenlm.KenLM` (bindings/python/flashlight/lib/text/decoder/kenlm.py):
ARPA back-off models, scored on the device."""
from text_b200.pyext import load as _load

KenLM = _load().KenLM

"""Loader of the pybind11 module `flashlight_lib_text_decoder` built in-tree by
`make -C text_b200/csrc pybind` (text_b200/lib/). The module links libflt_decoder.so (the CUDA
library); there is no CPU implementation behind it."""
import importlib.util
import os
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def path():
    return os.path.join(_HERE, "lib", "flashlight_lib_text_decoder" + sysconfig.get_config_var("EXT_SUFFIX"))


def load():
    global _mod
    if _mod is None:
        p = path()
        if not os.path.exists(p):
            raise ImportError(f"{p} not found: build it with `make -C text_b200/csrc pybind`")
        spec = importlib.util.spec_from_file_location("flashlight_lib_text_decoder", p)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
        sys.modules.setdefault("flashlight_lib_text_decoder", _mod)  # pickling looks classes up by module name
    return _mod

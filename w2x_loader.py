"""w2x_loader.py -- imports the hyphen-named package directory as the module `w2x_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "waifu2x-converter-cpp_b200")


def load():
    if "w2x_b200" in sys.modules:
        return sys.modules["w2x_b200"]
    spec = importlib.util.spec_from_file_location("w2x_b200", os.path.join(_PKG, "__init__.py"),
                                                  submodule_search_locations=[_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["w2x_b200"] = mod
    spec.loader.exec_module(mod)
    return mod

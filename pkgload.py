"""Loads the package directory ``prima.cpp_b200/`` (its name contains a dot, so a plain ``import`` cannot) as module
``prima_cpp_b200``."""
import importlib.util
import sys
from pathlib import Path

_NAME = "prima_cpp_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    pkg = Path(__file__).resolve().parent / "prima.cpp_b200"
    spec = importlib.util.spec_from_file_location(_NAME, pkg / "__init__.py", submodule_search_locations=[str(pkg)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod

"""Import shim: the package directory is named `optimpacknextgen.jl_b200` (it has
a dot, so `import` cannot name it); `import opkb200` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "optimpacknextgen.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "opkb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["opkb200"] = _mod
_spec.loader.exec_module(_mod)

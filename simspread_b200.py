"""Import shim: the product package lives in `simspread.jl_b200/` (a directory name Python cannot
import directly).  `import simspread_b200` loads that directory as the package `simspread_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simspread.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "simspread_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["simspread_b200"] = _mod
_spec.loader.exec_module(_mod)

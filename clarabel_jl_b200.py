"""Import shim: the package directory is named `clarabel.jl_b200/` (not a valid Python
identifier), so `import clarabel_jl_b200` loads it from there under this name."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clarabel.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "clarabel_jl_b200", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["clarabel_jl_b200"] = _mod
_spec.loader.exec_module(_mod)

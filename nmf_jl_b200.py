"""Import shim: makes the package directory `nmf.jl_b200/` (not a valid Python identifier) importable
as `nmf_jl_b200`.  `import nmf_jl_b200` returns the real package object."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nmf.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_root, "__init__.py"), submodule_search_locations=[_root]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

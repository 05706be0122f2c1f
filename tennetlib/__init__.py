"""Import shim: the product package lives in the directory `tennetlib.jl_b200/` (name fixed by the
repo layout); a dot cannot appear in a Python package name, so `import tennetlib.jl_b200` is wired up
here by loading that directory as the submodule `tennetlib.jl_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tennetlib.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "tennetlib.jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
jl_b200 = importlib.util.module_from_spec(_spec)
sys.modules["tennetlib.jl_b200"] = jl_b200
_spec.loader.exec_module(jl_b200)

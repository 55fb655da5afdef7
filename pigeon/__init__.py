"""Import shim: the package directory is `pigeon.jl_b200/` (named after the reference, StanfordASL/Pigeon.jl); this makes
`import pigeon.jl_b200` resolve to it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pigeon.jl_b200")
_spec = importlib.util.spec_from_file_location("pigeon.jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
jl_b200 = importlib.util.module_from_spec(_spec)
sys.modules["pigeon.jl_b200"] = jl_b200
_spec.loader.exec_module(jl_b200)

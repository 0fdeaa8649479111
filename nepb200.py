"""Import alias: `import nepb200` loads the package directory `nonlineareigenproblems.jl_b200/`
(whose name, fixed by the repo layout, is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nonlineareigenproblems.jl_b200")
_spec = importlib.util.spec_from_file_location("nepb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["nepb200"] = _mod
_spec.loader.exec_module(_mod)

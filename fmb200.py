"""Import alias: the package directory is `fmindex-collection_b200/` (not a valid Python identifier), so
`import fmb200` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fmindex-collection_b200")
_spec = importlib.util.spec_from_file_location(
    "fmb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fmb200"] = _mod
_spec.loader.exec_module(_mod)

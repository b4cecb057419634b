"""Import alias: the package lives in the directory `benlsip.jl_b200/` (a dotted name is not importable),
so `import benlsip_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "benlsip.jl_b200")
_spec = importlib.util.spec_from_file_location("benlsip_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["benlsip_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: the package directory is named `tfhe-research_b200` (hyphen, not importable as is)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tfhe-research_b200")
_spec = importlib.util.spec_from_file_location("tfhe_research_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tfhe_research_b200"] = _mod
_spec.loader.exec_module(_mod)

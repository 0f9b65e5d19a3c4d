"""Import shim: the package directory is named ``p-pad_b200`` (not a valid identifier), so
``import ppad_b200`` loads it from there and installs it in ``sys.modules`` under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "p-pad_b200")
_spec = importlib.util.spec_from_file_location("ppad_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ppad_b200"] = _mod
_spec.loader.exec_module(_mod)

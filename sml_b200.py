"""Import shim: exposes the package directory ``subspace-multimodal-learning_b200/`` as module ``sml_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "subspace-multimodal-learning_b200")
_spec = importlib.util.spec_from_file_location("sml_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sml_b200"] = _mod
_spec.loader.exec_module(_mod)

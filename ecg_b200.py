"""Import alias: the package directory is named `ecg-representation-learning_b200` (not a Python identifier),
so `import ecg_b200` loads it from there and registers it under this name."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ecg-representation-learning_b200')
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_DIR, '__init__.py'),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

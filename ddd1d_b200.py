"""Import shim: the package lives in the directory ``data-driven-discretization-1d_b200/``
(not a valid Python identifier), and is importable as ``ddd1d_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data-driven-discretization-1d_b200')
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_DIR, '__init__.py'), submodule_search_locations=[_DIR])
_module = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _module
_spec.loader.exec_module(_module)

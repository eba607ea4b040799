"""Import alias: makes the package directory ``sr-wavenet_b200/`` importable as
``sr_wavenet_b200`` (a hyphen is not valid in a Python module name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sr-wavenet_b200")
_spec = importlib.util.spec_from_file_location(
    "sr_wavenet_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sr_wavenet_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import alias: the product package directory is named ``tensorflow2-machine-vision_b200`` (not a valid
Python identifier), so ``import tfmv_b200`` loads it under this name, sub-packages included."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tensorflow2-machine-vision_b200")
_spec = importlib.util.spec_from_file_location(
    "tfmv_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tfmv_b200"] = _mod
_spec.loader.exec_module(_mod)
